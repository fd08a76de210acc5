"""Where a forward step of the PLIF / ALIF FireNet (cfg 5) goes: device time per kernel (torch.profiler) and wall time per step."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import event_flow_b200.models.model as M  # noqa: E402
from tools.bench_configs import FIRE, _events  # noqa: E402
from event_flow_b200.dataloader.encodings import encode_batch  # noqa: E402

dev = torch.device("cuda")
for cls in sys.argv[1:] or ["PLIFFireNet", "ALIFFireNet"]:
    torch.manual_seed(0)
    model = getattr(M, cls)(dict(FIRE))
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
    model = model.to(dev)
    vox = [encode_batch(_events(8, 1000, 128, 128, 7 + t).to(dev), (128, 128), 5)["event_voxel"] for t in range(6)]

    def steps():
        with torch.no_grad():
            for v in vox:
                model(v, None)

    steps()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps()
    torch.cuda.synchronize()
    print(f"{cls}: {(time.perf_counter() - t0) / len(vox) * 1e3:.3f} ms wall per step")
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        steps()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
