"""Kernel time of the tensor-core conv+LIF step vs. batch size (tiles per CTA), CUDA events over back-to-back launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_flow_b200 import ops  # noqa: E402
from oracle import spiking as osp  # noqa: E402

DEV = "cuda"
H = W = 128
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for rec in (False, True):
    params = osp.init_firenet_params("lif", 32, 32, seed=1, weight_gain=2.0)["G1" if rec else "R1a"]
    pd = {k: t.to(DEV).contiguous() for k, t in params.items()}
    ws = ops.split_weights(pd["ff"], pd.get("rec"))
    for B in (1, 2, 4, 8, 16, 32):
        g = torch.Generator().manual_seed(1)
        x_cl = ops.pack_cl((torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV))
        z_cl = ops.pack_cl((torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV))
        v = (torch.rand((B, 32, H, W), generator=g) * 1.2 - 0.1).to(DEV)
        args = (x_cl, v, z_cl, pd["ff"], pd.get("rec"), pd["leak"].reshape(-1), pd["thresh"].reshape(-1))
        for _ in range(3):
            ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
        # warm L2 (back-to-back) and cold (flush between launches)
        res = []
        for cold in (False, True):
            ts = []
            for _ in range(10):
                if cold:
                    flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            res.append(ts[len(ts) // 2])
        tiles = B * 128
        print(f"rec={rec} B={B:2d} tiles={tiles:5d} ({tiles / 148:.1f}/CTA): warm {res[0]:6.1f} us  cold {res[1]:6.1f} us  | algorithmic fp32 bytes {B * 10.48576:.0f} MB -> {B * 10.48576e6 / (res[1] * 1e-6) / 1e9:6.0f} GB/s cold")
