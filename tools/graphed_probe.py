"""Per-step forward time of the cell-by-cell models under no_grad: CUDA-graph replay (graphed.py) vs launch by launch."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.bench_configs import ANN, FIRE, UNET  # noqa: E402


def run(cls, cfg, B, H, W, bins, off, steps=60):
    import event_flow_b200.models.model as M
    from event_flow_b200.graphed import _StepGraph

    torch.manual_seed(0)
    c = dict(cfg)
    c["num_bins"] = bins
    m = getattr(M, cls)(c).cuda().eval()
    if off:
        m.__dict__["_graph_off"] = True
    vox = torch.randn(B, bins, H, W, device="cuda")
    cnt = torch.rand(B, 2, H, W, device="cuda")
    with torch.no_grad():
        for _ in range(6):
            m(vox, cnt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            m(vox, cnt)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
    ng = sum(isinstance(g, _StepGraph) for g in m.__dict__.get("_step_graphs", {}).values())
    print(f"{cls:22s} B={B} {H}x{W} graphs={'off' if off else ng}: {e0.elapsed_time(e1) / steps:.3f} ms/step device, host enqueue {(t1 - t0) / steps * 1e3:.3f} ms/step"
          f" {m.__dict__.get('_graph_error', '')}", flush=True)


if __name__ == "__main__":
    for off in (True, False):
        run("FireNet", ANN, 1, 128, 128, 1, off)
        run("PLIFFireNet", FIRE, 8, 128, 128, 5, off)
        run("ALIFFireNet", FIRE, 8, 128, 128, 5, off)
        run("EVFlowNet", dict(ANN, encoding="cnt"), 1, 256, 256, 2, off)
        run("RecEVFlowNet", dict(ANN, encoding="cnt"), 1, 256, 256, 2, off)
        run("PLIFRecEVFlowNet", dict(UNET, spiking_neuron={}), 4, 256, 256, 2, off)
