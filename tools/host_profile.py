"""cProfile of the host side of a training-mode forward window (debug tool)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from event_flow_b200.dataloader.encodings import encode_batch  # noqa: E402
from event_flow_b200.loss.flow import EventWarping  # noqa: E402
from event_flow_b200.models.model import LIFFireNet  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
model = LIFFireNet(bench.MODEL_CFG)
bench.scale_weights(model)
model = model.to(dev).train()
lossf = EventWarping(bench.LOSS_CFG, dev)
win = []
for e in bench.make_events(0, 0):
    ed = e.to(dev)
    d = encode_batch(ed, (bench.H, bench.W), bench.BINS)
    win.append((d["event_voxel"], d["event_cnt"], ed, d["event_list_pol_mask"], d["event_mask"]))


def fwd(grad):
    model.reset_states()
    lossf.reset()
    with torch.set_grad_enabled(grad):
        for vox, cnt, ev, pm, mask in win:
            out = model(vox, cnt)
            lossf.event_flow_association(out["flow"], ev.clone(), pm, mask)
    return out


for grad in (False, True, False, True):
    for _ in range(3):
        fwd(grad)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fwd(grad)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"grad={grad}: host {1e3 * (t1 - t0) / 5:.2f} ms/window, +sync {1e3 * (t2 - t0) / 5:.2f}")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    fwd(True)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
