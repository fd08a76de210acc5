"""Launches the tensor-core conv+LIF kernel (ff and recurrent) and the head kernel a few times at bench size, for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_flow_b200 import ops  # noqa: E402
from oracle import spiking as osp  # noqa: E402

DEV = "cuda"
B, H, W = 8, 128, 128
g = torch.Generator().manual_seed(1)
x = (torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV)
v = (torch.rand((B, 32, H, W), generator=g) * 1.2 - 0.1).to(DEV)
z = (torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV)
x_cl, z_cl = ops.pack_cl(x), ops.pack_cl(z)
x5 = torch.randn((B, 5, H, W), generator=g).to(DEV)
for rec in (False, True):
    params = osp.init_firenet_params("lif", 32, 32, seed=1, weight_gain=2.0)["G1" if rec else "R1a"]
    pd = {k: t.to(DEV).contiguous() for k, t in params.items()}
    ws = ops.split_weights(pd["ff"], pd.get("rec"))
    for _ in range(4):
        ops.lif_step_cl(x_cl, v, z_cl, pd["ff"], pd.get("rec"), pd["leak"].reshape(-1), pd["thresh"].reshape(-1), hard_reset=True, w_split=ws)
hp = osp.init_firenet_params("lif", 5, 32, seed=1, weight_gain=2.0)["head"]
hd = {k: t.to(DEV).contiguous() for k, t in hp.items()}
for _ in range(4):
    ops.lif_step_cl(None, v, z_cl, hd["ff"], None, hd["leak"].reshape(-1), hd["thresh"].reshape(-1), hard_reset=True, x_f32=x5)
torch.cuda.synchronize()
print("ok")
