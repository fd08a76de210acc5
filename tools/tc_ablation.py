"""Ablation timing of the tensor-core conv+LIF kernel (B=8, 128x128), back-to-back warm launches: what does each stream cost?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_flow_b200 import _lib as L  # noqa: E402
from event_flow_b200 import ops  # noqa: E402
from oracle import spiking as osp  # noqa: E402

DEV = "cuda"
B, H, W = 8, 128, 128
g = torch.Generator().manual_seed(1)
x_cl = ops.pack_cl((torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV))
z_cl = ops.pack_cl((torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV))
v = (torch.rand((B, 32, H, W), generator=g) * 1.2 - 0.1).to(DEV)
for rec in (False, True):
    params = osp.init_firenet_params("lif", 32, 32, seed=1, weight_gain=2.0)["G1" if rec else "R1a"]
    pd = {k: t.to(DEV).contiguous() for k, t in params.items()}
    ws = ops.split_weights(pd["ff"], pd.get("rec"))
    args = (x_cl, v, z_cl, pd["ff"], pd.get("rec"), pd["leak"].reshape(-1), pd["thresh"].reshape(-1))
    for mask, name in ((0, "production build"), (16384, "debug build, nothing off"), (1, "no v_out TMA store"), (2, "no v_in TMA load"), (3, "no v traffic"),
                       (8, "no z_out TMA store"), (1024, "no centre-z TMA load (ff)"), (1 + 2 + 8 + 1024, "only operand loads"), (4, "no MMAs"),
                       (4 + 1 + 2 + 8 + 1024, "no MMAs, only operand loads"), (32, "prologue + teardown only"), (128, "one tile per CTA")):
        L.lib().ef_debug_tc_skip(mask)
        for _ in range(3):
            ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
        e1.record()
        torch.cuda.synchronize()
        print(f"rec={rec} {name:36s} {e0.elapsed_time(e1) * 25:6.1f} us/launch (warm L2, back to back, incl. host allocation of outputs)")
    L.lib().ef_debug_tc_skip(0)
