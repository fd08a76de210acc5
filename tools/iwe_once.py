"""A few direct launches of the event-warping loss kernels on the two bench windows, for ncu:
   ncu --set full --clock-control none --import-source on -k regex:iwe_loss -o gpurun_out/prof_iwe python tools/iwe_once.py [large|cfg2]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from event_flow_b200 import _lib as L  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "large"
B, Hh, Ww, Tt, N = (32, 256, 256, 10, 50000) if which == "large" else (8, 128, 128, 10, 1000)
dev = torch.device("cuda")
g = torch.Generator(device="cpu").manual_seed(99)
ntot = Tt * N
ts = torch.rand((B, Tt, N), generator=g).sort(dim=2).values + torch.arange(Tt).view(1, Tt, 1)
ys = torch.randint(0, Hh, (B, Tt, N), generator=g).float()
xs = torch.randint(0, Ww, (B, Tt, N), generator=g).float()
ps = (torch.rand((B, Tt, N), generator=g) < 0.5).float() * 2 - 1
events = torch.stack([ts, ys, xs, ps], dim=3).reshape(B, ntot, 4).to(dev)
pol = torch.stack([(ps > 0).float(), (ps < 0).float()], dim=3).reshape(B, ntot, 2).to(dev)
flow = ((torch.rand((1, B, Tt, 2, Hh, Ww), generator=g) - 0.5) * 0.008).to(dev)
mask = torch.zeros((B, Tt, Hh * Ww))
mask.scatter_(2, (ys * Ww + xs).long(), 1.0)
mask = mask.view(B, Tt, Hh, Ww).to(dev)
p = L.IweLossParams()
p.S, p.B, p.T, p.T_maps, p.H, p.W = 1, B, Tt, Tt, Hh, Ww
p.n_total, p.n_per_pass = ntot, N
p.flow_scaling, p.weight = float(max(Hh, Ww)), 0.001
p.loss_scaling, p.smoothing_mask, p.overwrite_intermediate = 1, 1, 0
ws = torch.zeros(L.lib().ef_iwe_loss_workspace_elems(1, B, Hh, Ww), device=dev, dtype=torch.float32)
loss = torch.empty((), device=dev)
g_loss = torch.ones((), device=dev)
g_maps = torch.empty_like(flow)
p.events, p.pol_mask, p.flow_maps, p.event_mask = L.ptr(events), L.ptr(pol), L.ptr(flow), L.ptr(mask)
p.workspace, p.loss, p.g_loss, p.g_flow_maps = L.ptr(ws), L.ptr(loss), L.ptr(g_loss), L.ptr(g_maps)
for _ in range(3):
    L.call("ef_iwe_loss_fwd", p)
    L.call("ef_iwe_loss_bwd", p)
torch.cuda.synchronize()
print(which, "loss", loss.item())
