"""Timing of the windowed (time-fused) forward against the step-by-step forward of the same LIFFireNet (cfg 2 shape), with and without the
membrane potentials of all steps saved (training / inference).  usage: python tools/window_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from event_flow_b200.dataloader.encodings import encode_batch  # noqa: E402
from event_flow_b200.models.model import LIFFireNet  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
model = LIFFireNet(bench.MODEL_CFG)
bench.scale_weights(model)
model = model.to(dev)
wins = []
for k in range(4):
    vox = torch.stack([encode_batch(e.to(dev), (bench.H, bench.W), bench.BINS)["event_voxel"] for e in bench.make_events(0, k)])
    wins.append(vox)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    for k in range(6):
        fn(wins[k % 4])
    torch.cuda.synchronize()
    tot = 0.0
    for k in range(reps):
        flush.fill_(k & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(wins[k % 4])
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def stepwise(vox):
    with torch.no_grad():
        for t in range(vox.shape[0]):
            model(vox[t], None)


def windowed(vox):
    with torch.no_grad():
        model.forward_window(vox, None)


def stepwise_grad(vox):
    for t in range(vox.shape[0]):
        model(vox[t], None)
    model.detach_states()


def windowed_grad(vox):
    model.forward_window(vox, None)
    model.detach_states()


for name, fn in (("stepwise no-grad", stepwise), ("windowed no-grad", windowed), ("stepwise grad", stepwise_grad), ("windowed grad", windowed_grad)):
    model.reset_states()
    print(f"{name:20s} {timed(fn):.4f} ms / window of {bench.T} steps", flush=True)
