"""
oracle/ -- CPU restatement of the reference's hot-path algorithms.  TEST INFRASTRUCTURE ONLY.

Nothing in the shipped package (event_flow_b200/) imports this directory.  Only tests/, the smoke check in
__graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py may call it, and only as the checker
or as the CPU baseline -- never as the product path.

What is restated (reference = tudelft/event_flow, file:line into /root/reference):
  oracle/spiking.py    models/spiking_util.py:13-109, models/spiking_submodules.py:24-875 (cell maths),
                       models/model.py:148-286 (FireNet chain), models/submodules.py:12-83 (1x1 tanh head)
  oracle/iwe.py        utils/iwe.py:4-153, loss/flow.py:26-301 (EventWarping) + analytic gradient
  oracle/encodings.py  dataloader/encodings.py:30-85, dataloader/base.py:159-222
  oracle/unet.py       models/unet.py:224-465, models/spiking_submodules.py:878-1013, models/submodules.py:140-312 (U-Net family)
  oracle/annzoo.py     models/submodules.py:314-374,421-686, models/unet.py:148-222,468-480 (ConvLSTM / ConvRecurrent / ConvLeaky*
                       cells and the models built from them)

Arithmetic is torch-CPU fp32 (convolutions live in the un-vendored third-party dependency PyTorch; the reference
pins torch==1.7.0, this container has 2.11.0+cu128).  The reference ships no tests and no golden vectors, so the
oracle is pinned against the reference ITSELF, imported from /root/reference in the build container by
oracle/pin_against_reference.py, which also writes the committed fixtures under tests/golden/.
Parity status: pinned against reference outputs generated here (see tests/golden/MANIFEST.json).
"""
