"""Timing probes of the MMA issue pattern of the tensor-core conv+LIF kernel (results are wrong under these switches)."""
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
params = osp.init_firenet_params("lif", 32, 32, seed=1, weight_gain=2.0)["R1a"]
pd = {k: t.to(DEV).contiguous() for k, t in params.items()}
ws = ops.split_weights(pd["ff"], None)
args = (x_cl, v, z_cl, pd["ff"], None, pd["leak"].reshape(-1), pd["thresh"].reshape(-1))
trace = torch.zeros((148, 32, 8), dtype=torch.int64, device=DEV)
for mask, name in ((0, "production"), (16384, "debug build, nothing changed"), (2048, "N=32"), (4096, "aligned taps (dx=0)"), (8192, "same k-slice twice"),
                   (4096 + 8192, "aligned + same k-slice"), (2048 + 4096, "N=32 aligned"), (4, "no MMAs")):
    L.lib().ef_debug_tc_skip(mask)
    for _ in range(3):
        ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
    trace.zero_()
    L.lib().ef_debug_tc_trace(trace.data_ptr())
    ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
    torch.cuda.synchronize()
    L.lib().ef_debug_tc_trace(None)
    t = trace.cpu().double()
    acc = [t[:, it, 3][t[:, it, 3] > 0].median().item() for it in range(7)]
    end = t[:, :, 7].max(dim=1).values.median().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:32s} {e0.elapsed_time(e1) * 50:6.1f} us/launch (warm, back to back)  acc-complete cycles per tile: "
          + " ".join(f"{a:6.0f}" for a in acc) + f"  end {end:6.0f}")
L.lib().ef_debug_tc_skip(0)
