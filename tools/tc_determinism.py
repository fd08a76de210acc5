"""Repeat-launch determinism check of the tensor-core conv+LIF kernel against the CUDA-core kernel (race detector)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_flow_b200 import ops  # noqa: E402
from oracle import spiking as osp  # noqa: E402

DEV = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
total_bad = 0
for (B, H, W) in ((8, 128, 128), (3, 64, 72), (16, 128, 128)):
    for rec in (False, True):
        for hard in (True, False):
            g = torch.Generator().manual_seed(B * 7 + H)
            params = osp.init_firenet_params("lif", 32, 32, seed=B * 7 + H, weight_gain=2.0)["G1" if rec else "R1a"]
            x = (torch.rand((B, 32, H, W), generator=g) < 0.3).float()
            st = torch.rand((2, B, 32, H, W), generator=g) * 1.2 - 0.1
            st[1] = (st[1] < 0.3).float()
            pd = {k: v.to(DEV).contiguous() for k, v in params.items()}
            x_cl, v_in, z_in = ops.pack_cl(x.to(DEV)), st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV))
            ws = ops.split_weights(pd["ff"], pd.get("rec"))
            args = (x_cl, v_in, z_in, pd["ff"], pd.get("rec"), pd["leak"].reshape(-1), pd["thresh"].reshape(-1))
            v_cc, z_cc = ops.lif_step_cl(*args, hard_reset=hard, w_split=None)
            bad = []
            for _ in range(N):
                v_tc, z_tc = ops.lif_step_cl(*args, hard_reset=hard, w_split=ws)
                torch.cuda.synchronize()
                bad.append(int(((v_tc - v_cc).abs() > 1e-3).sum()) + int((z_tc != z_cc).sum() > 50))
            total_bad += sum(bad)
            print(f"B={B} {H}x{W} rec={rec} hard={hard}: launches with errors {sum(b > 0 for b in bad)}/{N} {bad[:8]}")
print("TOTAL", total_bad)
