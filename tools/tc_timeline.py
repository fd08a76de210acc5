"""Per-CTA timeline (clock64) of the tensor-core conv+LIF kernel at bench size: where does a tile's time go?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_flow_b200 import _lib as L  # noqa: E402
from event_flow_b200 import ops  # noqa: E402
from oracle import spiking as osp  # noqa: E402

DEV = "cuda"
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 128, 128
rec = len(sys.argv) > 2 and sys.argv[2] == "rec"
if len(sys.argv) > 3:
    L.lib().ef_debug_tc_cpt(int(sys.argv[3]))
g = torch.Generator().manual_seed(1)
x_cl = ops.pack_cl((torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV))
z_cl = ops.pack_cl((torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV))
v = (torch.rand((B, 32, H, W), generator=g) * 1.2 - 0.1).to(DEV)
params = osp.init_firenet_params("lif", 32, 32, seed=1, weight_gain=2.0)["G1" if rec else "R1a"]
pd = {k: t.to(DEV).contiguous() for k, t in params.items()}
ws = ops.split_weights(pd["ff"], pd.get("rec"))
args = (x_cl, v, z_cl, pd["ff"], pd.get("rec"), pd["leak"].reshape(-1), pd["thresh"].reshape(-1))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for _ in range(3):
    ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
trace = torch.zeros((148, 32, 8), dtype=torch.int64, device=DEV)
for cold in (False,):
    trace.zero_()
    if cold:
        flush.fill_(1)
    else:
        ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
    L.lib().ef_debug_tc_trace(trace.data_ptr())
    ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
    torch.cuda.synchronize()
    L.lib().ef_debug_tc_trace(None)
    t = trace.cpu().double()
    names = ["loads issued", "TMA landed (MMA start)", "MMAs issued", "acc complete (epi)", "tmem read", "STS done", "store issued", "end"]
    print(f"--- {'cold L2' if cold else 'warm L2'}: B={B} rec={rec}; median over CTAs of cycles since CTA start (1965 MHz: 1000 cyc = 0.51 us)")
    n_t = min(32, (B * 128 + 147) // 148)
    for it in range(min(n_t, 8)):
        row = []
        for s_ in range(8):
            col = t[:, it, s_]
            col = col[col > 0]
            row.append(f"{col.median().item():7.0f}" if col.numel() else "      -")
        print(f"tile {it}: " + " ".join(f"{n.split()[0][:5]}={r}" for n, r in zip(names, row)))
