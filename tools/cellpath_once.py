"""One ANN FireNet evaluation step (batch 1, 128x128, launch by launch) and one ALIF FireNet training window of 2 steps (batch 8) between
cudaProfilerStart/Stop: the kernels of the cell path for an `ncu --set full --profile-from-start off` capture."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.bench_configs import ANN, FIRE  # noqa: E402

import event_flow_b200.models.model as M  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
ann = M.FireNet(dict(ANN)).to(dev).eval()
ann.__dict__["_graph_off"] = True
alif = M.ALIFFireNet(dict(FIRE))
with torch.no_grad():
    for n, p in alif.named_parameters():
        if n.endswith("ff.weight") or n.endswith("rec.weight"):
            p.mul_(2.5)
alif = alif.to(dev).train()
v1, c1 = torch.randn(1, 1, 128, 128, device=dev), torch.rand(1, 2, 128, 128, device=dev)
v8 = (torch.rand(8, 5, 128, 128, device=dev) < 0.1).float() * torch.randn(8, 5, 128, 128, device=dev)
c8 = torch.rand(8, 2, 128, 128, device=dev)


def run():
    with torch.no_grad():
        ann(v1, c1)
    alif.zero_grad(set_to_none=True)
    alif.reset_states()
    loss = 0.0
    for _ in range(2):
        loss = loss + alif(v8, c8)["flow"][0].square().sum()
    loss.backward()


for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
