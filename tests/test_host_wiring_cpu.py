"""
Host-side wiring of every model class on the CPU (no GPU, no library compute calls): the launch wrappers in
event_flow_b200.ops are replaced by torch-CPU stand-ins built on the oracle, and each model is rolled out on the inputs of the
reference's golden fixtures.  This isolates the Python logic that mirrors the reference's modules -- layer order, state
threading, skip connections, residuals, crop / pad, multi-resolution flow upsampling, state_dict names -- from the CUDA kernels
(which the -m gpu tests cover): flows must match the reference's numbers stored in tests/golden/.
"""
import glob
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import spiking as osp
from tests.conftest import GOLDEN, load_golden


@pytest.fixture()
def cpu_ops(monkeypatch):
    from event_flow_b200 import ops

    def cell_step(neuron, x, state, w_ff, w_rec, chan, *, hard_reset, surrogate="arctanspike", width=10.0, stride=1, residual=None, x_kind=None,
                  detach=True):
        p = {"ff": w_ff, **chan}
        if w_rec is not None:
            p["rec"] = w_rec
        return osp.cell_step(neuron, x, state, p, hard_reset=hard_reset, surrogate=surrogate, width=width, stride=stride,
                             residual=0 if residual is None else residual, detach=detach)

    def conv_ann(x1, weight, bias, act, *, x2=None, x2_scale=None, residual=None, blend_h=None, blend_u=None, stride=1):
        x = x1 if x2 is None else torch.cat([x1, x2 if x2_scale is None else x2 * x2_scale], dim=1)
        out = F.conv2d(x, weight, bias, stride, 1)
        if residual is not None:
            out = out + residual
        out = {None: lambda t: t, "relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[act](out)
        return out if blend_h is None else blend_h * (1 - blend_u) + out * blend_u

    monkeypatch.setattr(ops, "cell_step", cell_step)
    monkeypatch.setattr(ops, "conv_ann", conv_ann)
    monkeypatch.setattr(ops, "pred_head", lambda x, w, b: torch.tanh(F.conv2d(x, w.reshape(w.shape[0], -1, 1, 1), b)))
    monkeypatch.setattr(ops, "upsample_bilinear2x", lambda x: F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False))
    monkeypatch.setattr(ops, "upsample_nearest", lambda x, fy, fx: x if fy == 1 and fx == 1 else F.interpolate(x, scale_factor=(float(fy), float(fx))))
    return ops


def _cfg(encoding, bins, base, spiking_neuron, acts):
    return dict(name="x", encoding=encoding, round_encoding=False, norm_input=False, num_bins=bins, base_num_channels=base, kernel_size=3,
                activations=acts, mask_output=True, spiking_neuron=spiking_neuron)


LIF_SN = dict(leak=[-4.0, 0.1], thresh=[0.8, 0.1], learn_leak=True, learn_thresh=True, hard_reset=True)
SPIKE = ["arctanspike", "arctanspike"]
RELU = ["relu", None]
CASES = {}
for _p in sorted(glob.glob(os.path.join(GOLDEN, "firenet_*.npz"))):
    _n = os.path.basename(_p)[:-4]
    _neuron, _enc = _n.split("_")[1:3]
    CASES[_n] = ({"lif": "LIFFireNet", "plif": "PLIFFireNet", "alif": "ALIFFireNet", "xlif": "XLIFFireNet"}[_neuron],
                 lambda g, e=_enc, nr=_neuron: _cfg(e, g["x_0"].shape[1], 32, LIF_SN if nr == "lif" else {}, SPIKE))
for _neuron, _cls in (("lif", "SpikingRecEVFlowNet"), ("plif", "PLIFRecEVFlowNet"), ("alif", "ALIFRecEVFlowNet"), ("xlif", "XLIFRecEVFlowNet")):
    CASES["unet_" + _neuron] = (_cls, lambda g, nr=_neuron: _cfg("cnt", 2, 4, None if nr == "lif" else {}, SPIKE))
CASES["ann_firenet"] = ("FireNet", lambda g: _cfg("voxel", 1, 32, None, RELU))
CASES["ann_fireflownet"] = ("FireFlowNet", lambda g: _cfg("cnt", 2, 32, None, RELU))
CASES["annunet_evflownet"] = ("EVFlowNet", lambda g: _cfg("cnt", 2, 4, None, RELU))
CASES["annunet_recevflownet"] = ("RecEVFlowNet", lambda g: _cfg("cnt", 2, 4, None, RELU))
for _n, _cls in (("rnnfirenet", "RNNFireNet"), ("leakyfirenet", "LeakyFireNet"), ("leakyfireflownet", "LeakyFireFlowNet"),
                 ("rnnrecevflownet", "RNNRecEVFlowNet"), ("leakyrecevflownet", "LeakyRecEVFlowNet"), ("e2vid", "E2VID")):
    CASES["annzoo_" + _n] = (_cls, lambda g, c=_cls: _cfg("cnt", 2, 8 if "Fire" in c else 4, {} if "Leaky" in c else None, RELU))


@pytest.mark.parametrize("name", sorted(CASES))
def test_model_wiring_reproduces_reference_flows_on_cpu(name, cpu_ops):
    import event_flow_b200.models.model as M

    cls, mk_cfg = CASES[name]
    g = load_golden(name)
    cfg = mk_cfg(g)
    torch.manual_seed(0)
    m = getattr(M, cls)(dict(cfg))
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd_")}
    assert list(m.state_dict().keys()) == list(sd.keys()), "state_dict names / order differ from the reference's"
    m.load_state_dict(sd)
    xs = [g[k] for k in sorted((k for k in g if k == "x" or k.startswith("x_")), key=lambda s: (len(s), s))]
    spiking = cfg["activations"][0] == "arctanspike"
    out = None
    with torch.no_grad():
        for t, x in enumerate(xs):
            out = m(x.clone(), x.clone(), log=name.startswith("firenet_"))
            if name.startswith("firenet_") and "activity" in g:  # log=True: fraction of non-zero entries per layer (model.py:267-284)
                assert list(out["activity"].keys()) == ["0:input", "1:head", "2:G1", "3:R1a", "4:R1b", "5:G2", "6:R2a", "7:R2b", "8:pred"]
                assert torch.allclose(torch.tensor(list(out["activity"].values())), g["activity"][t].float(), atol=1e-7)
            per_step = "flow_%d" % t in g and len(xs) > 1 and "flow_%d_0" % (len(xs) - 1) not in g and not name.startswith("annzoo") \
                and not name.startswith("annunet")
            if per_step:  # fixtures that store the flow of every step (FireNet families)
                _close(out["flow"][0], g["flow_%d" % t], spiking, f"{name} flow[{t}]")
    T = len(xs)
    if "flow_%d_0" % (T - 1) in g:  # spiking U-Nets: four scales of the last step
        for i in range(4):
            _close(out["flow"][i], g["flow_%d_%d" % (T - 1, i)], spiking, f"{name} scale {i}")
    elif name.startswith("annzoo") or name.startswith("annunet"):
        for i, f in enumerate(out["flow"]):
            _close(f, g["flow_%d" % i], spiking, f"{name} flow {i}")
    # the state API every model shares
    if cls != "EVFlowNet":  # stateless in the reference as well: only reset_states / detach_states exist
        assert m.states is not None
    m.detach_states()
    m.reset_states()


def _close(a, b, exact, what):
    if exact:  # spiking cells: the stand-in IS the oracle restatement, which is bit-equal to the reference
        assert torch.equal(a, b), what
    else:      # ANN cells: gate convolutions are batched differently from the reference (one conv for update+reset): fp32 noise
        assert (a - b).abs().max().item() <= 1e-5 * max(1.0, b.abs().max().item()), what


LOSSES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "loss_*.npz")))


@pytest.mark.parametrize("name", LOSSES)
def test_event_warping_bookkeeping_reproduces_reference_loss_on_cpu(name, monkeypatch):
    """
    EventWarping's host side (window accumulation, in-place timestamp offset, pass bookkeeping, overwrite_intermediate_flow,
    num_events / event_mask) with the kernel call replaced by the oracle's loss: value and flow gradients must be the reference's.
    """
    from event_flow_b200 import ops
    from event_flow_b200.loss.flow import EventWarping
    from oracle import iwe as oiwe

    def loss_stub(flows, events, pol_masks, masks, *, flow_scaling, weight, loss_scaling=True, smoothing_mask=True,
                  overwrite_intermediate=False):
        # the window arrives in pass form (per-pass tensors, nothing concatenated): restate it for the oracle
        passes = len(events)
        pass_of = torch.cat([torch.full((e.shape[1],), t) for t, e in enumerate(events)])
        maps = [torch.stack(per_scale, dim=1) for per_scale in flows]
        return oiwe.event_warping_loss(torch.cat(events, 1), torch.cat(pol_masks, 1), pass_of.long(), maps, torch.cat(masks, 1),
                                       tuple(maps[0].shape[-2:]), flow_scaling=flow_scaling, weight=weight, loss_scaling=loss_scaling,
                                       smoothing_mask=smoothing_mask, overwrite_intermediate=overwrite_intermediate, passes=passes)

    monkeypatch.setattr(ops, "event_warping_loss_passes", loss_stub)
    g = load_golden(name)
    scaling, smask, overwrite, weight, T, N = g["cfg"].tolist()
    T, N = int(T), int(N)
    flow = g["flow"]
    H, W = flow.shape[-2:]
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": weight, "overwrite_intermediate": bool(overwrite)},
           "model": {"mask_output": bool(smask)}}
    L = EventWarping(cfg, "cpu", loss_scaling=bool(scaling))
    flows = [flow[:, t].clone().requires_grad_(True) for t in range(T)]
    for t in range(T):
        e = g["events"][:, t * N:(t + 1) * N].clone()
        e[:, :, 0] -= t  # the fixture stores offset timestamps; the API offsets them itself, in place (loss/flow.py:90)
        L.event_flow_association([flows[t]], e, g["pol_mask"][:, t * N:(t + 1) * N], g["event_mask"][:, t:t + 1])
        assert torch.equal(e, g["events"][:, t * N:(t + 1) * N]), "the caller's event tensor must carry the pass offset afterwards"
    assert L.num_events == T * N
    if overwrite:
        L.overwrite_intermediate_flow([flows[-1]])
        assert L.event_mask.shape[1] == 1
    loss = L()
    loss.backward()
    assert abs(loss.item() - g["loss"].item()) <= 1e-6 * abs(g["loss"].item())
    for t, f in enumerate(flows):
        if f"grad_{t}" in g:
            ref = g[f"grad_{t}"]
            assert (f.grad - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-12
    L.reset()
    assert L.num_events == 0


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(GOLDEN.rstrip("/")).rsplit("/tests", 1)[0]
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "events/s" and line["higher_is_better"] is True and line["value"] > 0
    staged = os.path.isfile(os.path.join(root, "baseline", "_ref", "models", "model.py"))  # the unmodified reference, if staged (tools/stage_reference.py)
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"] and line["steps"] == 1 and line["warmup"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"] and line["metric"].startswith("events/s")


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("cls", ["EVFlowNet", "RecEVFlowNet", "SpikingRecEVFlowNet", "E2VID"])
def test_cropping_and_input_normalisation_match_the_live_reference(cls, cpu_ops):
    """init_cropping (pad to a size the pyramid divides, crop back) and norm_input against the unmodified reference, run live on CPU."""
    import importlib
    import sys

    import event_flow_b200.models.model as M

    spiking = "Spiking" in cls
    cfg = dict(name=cls, encoding="cnt", round_encoding=False, norm_input=True, num_bins=2, base_num_channels=4, kernel_size=3,
               activations=["arctanspike", "arctanspike"] if spiking else ["relu", None], mask_output=True, spiking_neuron=None)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    sys.path.insert(0, "/root/reference")
    try:
        ref_model = importlib.import_module("models.model")
        torch.manual_seed(5)
        ref = getattr(ref_model, cls)(dict(cfg))
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    torch.manual_seed(5)
    mine = getattr(M, cls)(dict(cfg))
    mine.load_state_dict(ref.state_dict())
    H, W = 27, 43  # neither is divisible by 2^3 / 2^4
    ref.init_cropping(W, H), mine.init_cropping(W, H)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for _ in range(2):
            cnt = torch.randint(0, 3, (1, 2, H, W), generator=g).float()
            a = mine(None, cnt.clone())["flow"]
            b = ref(None, cnt.clone())["flow"]
            assert len(a) == len(b)
            for fa, fb in zip(a, b):
                assert fa.shape == fb.shape == (1, 2, H, W)
                if spiking:
                    assert torch.equal(fa, fb)
                else:
                    assert (fa - fb).abs().max().item() <= 1e-5 * max(1.0, fb.abs().max().item())


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("opts", [dict(norm="BN"), dict(norm="IN"), dict(use_upsample_conv=False), dict(norm="BN", use_upsample_conv=False)])
def test_norm_and_transposed_conv_options_match_the_live_reference(opts, cpu_ops):
    """
    The optional layers of the ANN U-Nets (models/model.py:39-52: `norm`, `use_upsample_conv`; models/submodules.py:45-49,86-137):
    same state_dict names, same flows and recurrent states as the unmodified reference (train mode: batch statistics), run live on CPU.
    """
    import importlib
    import sys

    import event_flow_b200.models.model as M

    cfg = dict(name="RecEVFlowNet", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=4, kernel_size=3,
               activations=["relu", None], mask_output=True, spiking_neuron=None, **opts)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    sys.path.insert(0, "/root/reference")
    try:
        ref_model = importlib.import_module("models.model")
        torch.manual_seed(7)
        ref = ref_model.RecEVFlowNet(dict(cfg)).train()
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    torch.manual_seed(7)
    mine = M.RecEVFlowNet(dict(cfg)).train()
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    for a, b in zip(mine.state_dict().values(), ref.state_dict().values()):
        assert torch.equal(a, b), "same seed, same initial values"
    g = torch.Generator().manual_seed(2)
    H, W = 32, 48
    for _ in range(2):
        cnt = torch.randint(0, 3, (3, 2, H, W), generator=g).float()
        a, b = mine(None, cnt.clone())["flow"], ref(None, cnt.clone())["flow"]
        assert len(a) == len(b) == 4
        for fa, fb in zip(a, b):
            assert fa.shape == fb.shape and (fa - fb).abs().max().item() <= 2e-5 * max(1.0, fb.abs().max().item())
    loss_a, loss_b = sum(f.square().sum() for f in a), sum(f.square().sum() for f in b)
    loss_a.backward(), loss_b.backward()
    # (a conv bias in front of an instance norm has an exactly-zero true gradient: what autograd reports there is rounding noise, so
    # the scale of the comparison is the largest gradient of the model, not of the parameter)
    scale = max(pb.grad.abs().max().item() for pb in ref.parameters())
    for (n, pa), pb in zip(mine.named_parameters(), ref.parameters()):
        assert (pa.grad - pb.grad).abs().max().item() <= 1e-3 * max(pb.grad.abs().max().item(), 1e-2 * scale), n


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("cls,opts", [("LIFFireNet", dict(norm="weight")), ("LIFFireNet", dict(norm="group")), ("LIFFireNet", dict(detach=False)),
                                      ("ALIFFireNet", dict(detach=False, norm="weight")), ("PLIFFireNet", dict(detach=False)), ("XLIFFireNet", dict(detach=False)),
                                      ("SpikingRecEVFlowNet", dict(norm="weight", detach=False))])
def test_spiking_cell_options_match_the_live_reference(cls, opts, cpu_ops):
    """
    The cell options no shipped yml sets (spiking_submodules.py:86-94,110-112,501-514): weight / group normalisation and a
    differentiable reset (detach=False): same state_dict names, same flows over three steps, same BPTT gradients as the unmodified
    reference, run live on CPU (the kernel call replaced by the oracle's cell step).
    """
    import importlib
    import sys

    import event_flow_b200.models.model as M

    fire = "FireNet" in cls
    sn = dict(opts) if fire else dict(opts)
    # (group norm: the reference builds GroupNorm(min(1, input_size // 4), input_size), which needs at least 4 input channels)
    cfg = dict(name=cls, encoding="voxel" if fire else "cnt", round_encoding=False, norm_input=False, num_bins=4 if fire else 2,
               base_num_channels=8 if fire else 4, kernel_size=3, activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron=sn)
    if not fire and "norm" in opts:
        cfg["norm"] = opts["norm"]  # the U-Net takes the normalisation from the model config (model.py:426-439)
        cfg["spiking_neuron"] = {k: v for k, v in opts.items() if k != "norm"}
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    sys.path.insert(0, "/root/reference")
    try:
        ref_model = importlib.import_module("models.model")
        torch.manual_seed(11)
        if fire:
            getattr(ref_model, cls).kwargs = [{}] * 7
        ref = getattr(ref_model, cls)({**cfg, "spiking_neuron": dict(cfg["spiking_neuron"])}).train()
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    torch.manual_seed(11)
    mine = getattr(M, cls)({**cfg, "spiking_neuron": dict(cfg["spiking_neuron"])}).train()
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    with torch.no_grad():  # spikes through every layer on a small input
        for m in (mine, ref):
            for n, p in m.named_parameters():
                if "ff.weight" in n or "rec.weight" in n:
                    p.mul_(3.0)
    g = torch.Generator().manual_seed(3)
    H, W = 32, 32
    la = lb = 0.0
    for _ in range(3):
        x = torch.randint(0, 3, (2, cfg["num_bins"], H, W), generator=g).float()
        a = mine(x.clone(), x.clone())["flow"]
        b = ref(x.clone(), x.clone())["flow"]
        for fa, fb in zip(a, b):
            assert (fa - fb).abs().max().item() <= 1e-5 * max(1.0, fb.abs().max().item())
        la, lb = la + sum(f.square().sum() for f in a), lb + sum(f.square().sum() for f in b)
    la.backward(), lb.backward()
    scale = max(pb.grad.abs().max().item() for pb in ref.parameters() if pb.grad is not None)
    assert scale > 0
    for (n, pa), pb in zip(mine.named_parameters(), ref.parameters()):
        if pb.grad is None:
            assert pa.grad is None or pa.grad.abs().max() == 0, n
            continue
        assert (pa.grad - pb.grad).abs().max().item() <= 1e-4 * max(pb.grad.abs().max().item(), 1e-2 * scale), n


def test_cells_are_told_what_their_input_is():
    """
    The input kinds the models vouch for (exact-in-bf16 "spikes", or ("mixed", n): n fractional channels first) decide whether a cell may
    run on the tensor cores with exact products -- they must match the wiring of the forward passes: FireNet's head sees the encoding,
    every later spiking cell spikes; in the spiking U-Net the decoders after the first receive cat[prediction, x, skip] (unet.py:418-465).
    """
    import event_flow_b200.models.model as M

    fire = dict(name="x", encoding="voxel", round_encoding=False, norm_input=False, num_bins=5, base_num_channels=32, kernel_size=3,
                activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron={})
    m = M.LIFFireNet(dict(fire))
    kinds = {n: getattr(m, n).__dict__.get("_x_kind") for n in ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b")}
    assert kinds == {"head": "split", "G1": "spikes", "R1a": "spikes", "R1b": "spikes", "G2": "spikes", "R2a": "spikes", "R2b": "spikes"}
    unet = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=32, kernel_size=3,
                activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron=None)
    net = M.SpikingRecEVFlowNet(dict(unet)).net
    kind = lambda cell: cell.__dict__.get("_x_kind")  # noqa: E731
    assert kind(net.encoders[0].conv) is None  # the network input: event counts or a voxel grid, whatever the caller feeds
    assert all(kind(e.conv) == "spikes" for e in net.encoders[1:]) and all(kind(e.recurrent_block) == "spikes" for e in net.encoders)
    assert all(kind(r.conv1) == "spikes" and kind(r.conv2) == "spikes" for r in net.resblocks)
    assert kind(net.decoders[0].conv2d) == "spikes"
    assert all(kind(d.conv2d) == ("mixed", net.num_output_channels) for d in net.decoders[1:])
    # the channel bookkeeping behind ("mixed", n): prediction channels come first in the decoder's input
    for i, d in enumerate(net.decoders[1:], start=1):
        assert d.conv2d.input_size == 2 * net.encoder_output_sizes[::-1][i] + net.num_output_channels
    # the ANN / leaky twins share the wiring but have no such cells: nothing to mark, nothing breaks
    ann_cfg = dict(unet, activations=["relu", None])
    M.LeakyRecEVFlowNet(dict(ann_cfg)), M.RecEVFlowNet(dict(ann_cfg))
