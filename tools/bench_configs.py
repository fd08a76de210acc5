"""
Timings of the other BASELINE.json configurations (they are parity-test cases, not bench lines; numbers for the record,
comparable with profiles/r01_reference_probe_b200.txt): cfg 4 (SpikingRecEVFlowNet 256x256, 50k events/window, batch 4 per
GPU) forward and train step, cfg 5 (PLIF / ALIF FireNet, 20-step sequence, batch 8) forward + loss + backward.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import event_flow_b200.models.model as M  # noqa: E402
from event_flow_b200.dataloader.encodings import encode_batch  # noqa: E402
from event_flow_b200.loss.flow import EventWarping  # noqa: E402
from bench import synthetic_events as _syn  # noqa: E402


def synthetic_events(B, N, H, W, seed):
    import bench
    bench.H, bench.W = H, W
    e = _syn(B, N, seed)
    return e[:, :, 0], e[:, :, 1], e[:, :, 2], e[:, :, 3]

dev = torch.device("cuda")


def windows(B, N, H, W, T, bins, seed):
    out = []
    for t in range(T):
        ts, ys, xs, ps = synthetic_events(B, N, H, W, seed + t)
        ev = torch.stack([ts, ys, xs, ps], dim=2).to(dev)
        d = encode_batch(ev, (H, W), bins)
        out.append((d["event_voxel"], d["event_cnt"], ev, d["event_list_pol_mask"], d["event_mask"]))
    return out


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(name, cls, cfg, B, N, H, W, T, bins, gain):
    torch.manual_seed(0)
    model = getattr(M, cls)(dict(cfg))
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(gain)
    model = model.to(dev).train()
    lossf = EventWarping({"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False},
                          "model": {"mask_output": True}}, dev)
    win = windows(B, N, H, W, T, bins, 7)

    def fwd():
        model.reset_states()
        with torch.no_grad():
            for vox, cnt, ev, pm, mask in win:
                model(vox, cnt)

    def train():
        model.zero_grad(set_to_none=True)
        model.reset_states()
        lossf.reset()
        for vox, cnt, ev, pm, mask in win:
            out = model(vox, cnt)
            lossf.event_flow_association(out["flow"], ev.clone(), pm, mask)
        loss = lossf()
        loss.backward()
        model.detach_states()

    ms_f = timed(fwd, reps=5)
    if os.environ.get("EF_FWD_ONLY"):
        print(json.dumps({"config": name, "model": cls, "batch": B, "resolution": [H, W], "timesteps": T, "fwd_ms_per_step": ms_f / T}), flush=True)
        return
    ms_t = timed(train)
    print(json.dumps({"config": name, "model": cls, "batch": B, "resolution": [H, W], "timesteps": T, "events_per_window": N,
                      "fwd_ms_per_step": ms_f / T, "fwd_loss_bwd_ms_per_window": ms_t, "events_per_s_train": B * N * T / ms_t * 1e3}), flush=True)


unet = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=32, kernel_size=3,
            activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron=None)
fire = dict(name="x", encoding="voxel", round_encoding=False, norm_input=False, num_bins=5, base_num_channels=32, kernel_size=3,
            activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron={})
run("cfg4", "SpikingRecEVFlowNet", unet, 4, 50000, 256, 256, 2, 2, 3.0)
run("cfg5-plif", "PLIFFireNet", fire, 8, 1000, 128, 128, 20, 5, 2.5)
run("cfg5-alif", "ALIFFireNet", fire, 8, 1000, 128, 128, 20, 5, 2.5)
