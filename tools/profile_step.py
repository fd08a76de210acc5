"""One warm forward window (+ loss) and one train step of the bench workload, for `ncu` launch lists.
   ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
       python tools/profile_step.py [fwd|train|both] [stepwise]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from event_flow_b200.dataloader.encodings import encode_batch  # noqa: E402
from event_flow_b200.loss.flow import EventWarping  # noqa: E402
from event_flow_b200.models.model import LIFFireNet  # noqa: E402
from event_flow_b200.parallel import DataParallelTrainer  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "both"
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


vox_all, cnt_all = torch.stack([w[0] for w in win]), torch.stack([w[1] for w in win])
stepwise = len(sys.argv) > 2 and sys.argv[2] == "stepwise"


def window(train):
    model.reset_states()
    lossf.reset()
    with torch.set_grad_enabled(train):
        if stepwise:
            outs = [model(vox, cnt) for vox, cnt, _, _, _ in win]
        else:
            outs = model.forward_window(vox_all, cnt_all)
        for out, (_, _, ev, pm, mask) in zip(outs, win):
            lossf.event_flow_association(out["flow"], ev.clone(), pm, mask)
        loss = lossf()
    if train:
        loss.backward()
        trainer.step()
        model.detach_states()
    return loss


for _ in range(2):
    window(False)
    if mode != "fwd":
        window(True)
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: only the measured window(s), on every thread (the backward runs on autograd's)
if mode in ("fwd", "both"):
    window(False)
if mode in ("train", "both"):
    window(True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
