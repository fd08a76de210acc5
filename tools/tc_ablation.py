"""Ablation timing of the tensor-core conv+LIF kernel (B=8, 128x128): which part of the per-tile loop costs what?"""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
import time


def empty_launch():
    ts = []
    a = torch.zeros(1024, device=DEV)
    for _ in range(12):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a.add_(1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f'tiny torch kernel after flush: {ts[len(ts)//2]:.1f} us (event-to-event floor)')


empty_launch()
for rec in (False, True):
    params = osp.init_firenet_params("lif", 32, 32, seed=1, weight_gain=2.0)["G1" if rec else "R1a"]
    pd = {k: t.to(DEV).contiguous() for k, t in params.items()}
    ws = ops.split_weights(pd["ff"], pd.get("rec"))
    args = (x_cl, v, z_cl, pd["ff"], pd.get("rec"), pd["leak"].reshape(-1), pd["thresh"].reshape(-1))
    for mask, name in ((0, "full"), (256, "tile-blocked membrane addressing"), (1, "no v_out stores"), (2, "no v_in loads"), (1024, "no z_in loads"), (2 + 1024, "no v_in, no z_in loads"), (1 + 8, "no v_out, no z_out stores"), (3, "no v traffic"), (4, "no MMAs"), (8, "no spike store/barriers"),
                       (16, "no tmem loads"), (4 + 16, "no MMA, no tmem ld"), (1 + 2 + 8, "no v traffic, no spike store"), (31, "everything off"),
                       (32, "prologue+teardown only"), (32 + 64, "prologue w/o weights"), (128, "1 tile per CTA"), (128 + 31, "1 tile/CTA, everything off"),
                       (64 + 31, "everything off, no weights"), (512 + 31, "everything off, no tile loads (barrier ring only)")):
        L.lib().ef_debug_tc_skip(mask)
        ts = []
        for _ in range(12):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print(f"rec={rec} {name:32s} {ts[len(ts) // 2]:6.1f} us")
    L.lib().ef_debug_tc_skip(0)
