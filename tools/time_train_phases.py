"""Per-phase CUDA-event timing + torch.profiler kernel table of one train step of the bench workload (debug tool)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from event_flow_b200.dataloader.encodings import encode_batch  # noqa: E402
from event_flow_b200.loss.flow import EventWarping  # noqa: E402
from event_flow_b200.models.model import LIFFireNet  # noqa: E402
from event_flow_b200.parallel import DataParallelTrainer  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
model = LIFFireNet(bench.MODEL_CFG)
bench.scale_weights(model)
model = model.to(dev).train()
lossf = EventWarping(bench.LOSS_CFG, dev)
trainer = DataParallelTrainer(model)
win = []
for e in bench.make_events(0, 0):
    ed = e.to(dev)
    d = encode_batch(ed, (bench.H, bench.W), bench.BINS)
    win.append((d["event_voxel"], d["event_cnt"], ed, d["event_list_pol_mask"], d["event_mask"]))


def step(verbose=False):
    ts = [time.perf_counter()]
    model.reset_states()
    lossf.reset()
    for vox, cnt, ev, pm, mask in win:
        out = model(vox, cnt)
        lossf.event_flow_association(out["flow"], ev.clone(), pm, mask)
    torch.cuda.synchronize(); ts.append(time.perf_counter())
    loss = lossf()
    torch.cuda.synchronize(); ts.append(time.perf_counter())
    loss.backward()
    torch.cuda.synchronize(); ts.append(time.perf_counter())
    trainer.step()
    model.detach_states()
    torch.cuda.synchronize(); ts.append(time.perf_counter())
    if verbose:
        print("fwd %.2f ms | loss %.2f | bwd %.2f | opt %.2f | activity %s" % (*[(b - a) * 1e3 for a, b in zip(ts, ts[1:])],
              ["%.3f" % z.float().mean().item() for z in model._last_spikes]))


for _ in range(3):
    step(True)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
