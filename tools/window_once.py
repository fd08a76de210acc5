"""Two windows of the bench model through the windowed entry point with direct launches (no CUDA graph), for ncu:
   ncu --set full --clock-control none --import-source on -k regex:lif_conv_fwd_tc -s 25 -c 25 -o gpurun_out/prof_window python tools/window_once.py [grad]
Launch order of one window: head (fused window), G1 x T, R1a, R1b (fused window), G2 x T, R2a, R2b (fused window), prediction."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from event_flow_b200.dataloader.encodings import encode_batch  # noqa: E402
from event_flow_b200.models.model import LIFFireNet  # noqa: E402

grad = len(sys.argv) > 1 and sys.argv[1] == "grad"
dev = torch.device("cuda")
torch.manual_seed(0)
model = LIFFireNet(bench.MODEL_CFG)
bench.scale_weights(model)
model = model.to(dev)
model._use_graphs = False
for k in range(3):
    vox = torch.stack([encode_batch(e.to(dev), (bench.H, bench.W), bench.BINS)["event_voxel"] for e in bench.make_events(0, k)])
    with torch.set_grad_enabled(grad):
        out = model.forward_window(vox, None)
    if grad:
        model.detach_states()
torch.cuda.synchronize()
print("ok", float(out[-1]["flow"][0].abs().mean()))
