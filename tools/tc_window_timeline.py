"""Per-CTA timeline (clock64) of the TIME-FUSED tensor-core conv+LIF launch (ef_lif_conv_fwd_window, T steps per tile): per item (tile, step)
the cycles of the pipeline events, medians over the CTAs.  usage: python tools/tc_window_timeline.py [T] [save_all_v] [cpt]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_flow_b200 import _lib as L  # noqa: E402
from event_flow_b200 import ops  # noqa: E402
from oracle import spiking as osp  # noqa: E402

DEV = "cuda"
B, H, W = 8, 128, 128
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
save_all = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if len(sys.argv) > 3:
    L.lib().ef_debug_tc_cpt(int(sys.argv[3]))  # channels per epilogue thread: 8 = 16 epilogue warps (default), 16 = 8 warps
g = torch.Generator().manual_seed(1)
x_cl = torch.stack([ops.pack_cl((torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV)) for _ in range(T)])
z_cl = ops.pack_cl((torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV))
v = (torch.rand((B, 32, H, W), generator=g) * 1.2 - 0.1).to(DEV)
params = osp.init_firenet_params("lif", 32, 32, seed=1, weight_gain=2.0)["R1a"]
pd = {k: t.to(DEV).contiguous() for k, t in params.items()}
ws = ops.split_weights(pd["ff"], None)
v_out = torch.empty((T if save_all else 1, B, 32, H, W), device=DEV)
z_out = torch.empty((T, B, H, W, 32), device=DEV, dtype=torch.bfloat16)
q = L.LifConvWindowParams()
q.B, q.T, q.H, q.W, q.hard_reset, q.save_all_v = B, T, H, W, 1, save_all
q.x_cl, q.v_in, q.z_in_cl = L.ptr(x_cl), L.ptr(v), L.ptr(z_cl)
q.leak, q.thresh, q.w_split = L.ptr(pd["leak"].reshape(-1).contiguous()), L.ptr(pd["thresh"].reshape(-1).contiguous()), L.ptr(ws)
q.v_out, q.z_out_cl = L.ptr(v_out), L.ptr(z_out)
for _ in range(3):
    L.call("ef_lif_conv_fwd_window", q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    L.call("ef_lif_conv_fwd_window", q)
e1.record()
torch.cuda.synchronize()
print(f"production kernel: {e0.elapsed_time(e1) * 100:.1f} us per launch (T={T}, save_all_v={save_all}, warm L2, back to back)")
trace = torch.zeros((148, 32, 8), dtype=torch.int64, device=DEV)
L.lib().ef_debug_tc_trace(trace.data_ptr())
L.call("ef_lif_conv_fwd_window", q)
torch.cuda.synchronize()
L.lib().ef_debug_tc_trace(None)
t = trace.cpu().double()
names = ["load", "mma0", "mma1", "acc", "tmem", "sts", "store", "end"]
med = torch.zeros(32, 8)
for it in range(32):
    for s in range(8):
        col = t[:, it, s]
        col = col[col > 0]
        med[it, s] = col.median() if col.numel() else float("nan")
print("item: " + " ".join(f"{n:>7s}" for n in names[:7]) + "   | d(acc) mma_dur epi(acc->store) tmem-acc sts-tmem store-sts")
for it in range(32):
    r = med[it]
    d_acc = r[3] - med[it - 1, 3] if it else float("nan")
    print(f"{it:4d}: " + " ".join(f"{v:7.0f}" for v in r[:7]) + f"   | {d_acc:6.0f} {r[2] - r[1]:7.0f} {r[6] - r[3]:7.0f} {r[4] - r[3]:7.0f} {r[5] - r[4]:7.0f} {r[6] - r[5]:7.0f}")
