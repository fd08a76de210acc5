"""
Timings of the other BASELINE.json configurations (they are parity-test cases, not the bench line; bench.py reports them under
`other_configs`, comparable with profiles/r01_reference_probe_b200.txt):
  cfg 1  FireNet (ANN: ConvLayer_ cells + ConvGRU), 1 bin, batch 1, 128x128 (configs/eval_flow.yml plumbing): forward per step, training window;
  cfg 4  SpikingRecEVFlowNet 256x256, 50k events / window, batch 4 per GPU (32 over 8 GPUs): forward per step (general tcgen05 cell
         kernel, ef_lif_conv_fwd_g) with its tensor-pipe roofline, and the training window (fwd + EventWarping loss + BPTT);
  cfg 5  PLIF / ALIF FireNet, 20-step sequence, batch 8, 128x128: forward per step and the training window.
usage: python tools/bench_configs.py        (one JSON line per configuration)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

UNET = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=32, kernel_size=3,
            activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron=None)
FIRE = dict(name="x", encoding="voxel", round_encoding=False, norm_input=False, num_bins=5, base_num_channels=32, kernel_size=3,
            activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron={})
ANN = dict(name="x", encoding="voxel", round_encoding=False, norm_input=False, num_bins=1, base_num_channels=32, kernel_size=3,
           activations=["relu", None], mask_output=True, spiking_neuron=None)
CONFIGS = (
    ("cfg1", "FireNet", ANN, dict(B=1, N=1000, H=128, W=128, T=10, bins=1, gain=1.0)),  # the reference's own CPU-runnable case, on the GPU
    ("cfg4", "SpikingRecEVFlowNet", UNET, dict(B=4, N=50000, H=256, W=256, T=4, bins=2, gain=3.0)),
    ("cfg5-plif", "PLIFFireNet", FIRE, dict(B=8, N=1000, H=128, W=128, T=20, bins=5, gain=2.5)),
    ("cfg5-alif", "ALIFFireNet", FIRE, dict(B=8, N=1000, H=128, W=128, T=20, bins=5, gain=2.5)),
)


def _events(B, N, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    ts = torch.sort(torch.rand(B, N, generator=g))[0]
    ts = (ts - ts[:, :1]) / (ts[:, -1:] - ts[:, :1])
    ys = torch.randint(0, H, (B, N), generator=g).float()
    xs = torch.randint(0, W, (B, N), generator=g).float()
    ps = (torch.randint(0, 2, (B, N), generator=g) * 2 - 1).float()
    return torch.stack([ts, ys, xs, ps], dim=2)


def _timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(name, cls, cfg, B, N, H, W, T, bins, gain, dev, peaks=None, reps=3):
    import event_flow_b200.models.model as M
    from event_flow_b200 import _lib as L
    from event_flow_b200.dataloader.encodings import encode_batch
    from event_flow_b200.loss.flow import EventWarping

    torch.manual_seed(0)
    model = getattr(M, cls)(dict(cfg))
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(gain)
    model = model.to(dev).train()
    lossf = EventWarping({"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False},
                          "model": {"mask_output": True}}, dev)
    win = []
    for t in range(T):
        ev = _events(B, N, H, W, 7 + t).to(dev)
        d = encode_batch(ev, (H, W), bins)
        win.append((d["event_voxel"], d["event_cnt"], ev, d["event_list_pol_mask"], d["event_mask"]))

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

    fwd()  # (the no-grad step of the cell-by-cell models becomes a CUDA graph on its second visit: event_flow_b200/graphed.py)
    ms_f = _timed(fwd, reps)
    # tensor-core work of one forward window: multiply-accumulates of the general tcgen05 cell launches, counted from their arguments
    macs, orig = [0, 0], L.call

    def counting_call(fn, params, tag=None):
        if fn == "ef_lif_conv_fwd_g":
            k = sum(params.src_c[i] for i in range(params.n_src))
            macs[0] += params.B * params.H * params.W * params.C * k * (4 if params.s2d else 9)
            macs[1] += 1
        return orig(fn, params, tag)

    L.call = counting_call
    try:
        fwd()
    finally:
        L.call = orig
    torch.cuda.synchronize()
    ms_t = _timed(train, reps)
    out = {"config": name, "model": cls, "batch": B, "resolution": [H, W], "timesteps": T, "events_per_window": N,
           "fwd_ms_per_step": ms_f / T, "fwd_loss_bwd_ms_per_window": ms_t, "events_per_s_train": B * N * T / ms_t * 1e3}
    if macs[1]:
        # every fp32 weight is three exact bf16 terms: the tensor cores execute 3 x the convolution's multiply-accumulates
        conv_tf = 2 * macs[0] / (ms_f * 1e-3) / 1e12
        peak = (peaks or {}).get("bf16_tflops_sustained")
        out["tensor"] = {"tcgen05_cell_launches_per_window": macs[1], "conv_TFLOPs_fp32_equivalent": conv_tf, "bf16_TFLOPs_executed": 3 * conv_tf,
                         "split_factor": 3, "peak_bf16_tflops_sustained": peak, "frac_of_peak_executed": None if not peak else 3 * conv_tf / peak,
                         "note": "whole forward window (all kernels, host launch gaps included) against the tensor-core work of its tcgen05 cell launches"}
    del model, lossf, win
    torch.cuda.empty_cache()
    return out


def run_all(dev, peaks=None, reps=3):
    return [run(name, cls, cfg, dev=dev, peaks=peaks, reps=reps, **kw) for name, cls, cfg, kw in CONFIGS]


if __name__ == "__main__":
    pk = None
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    for name, cls, cfg, kw in CONFIGS:
        print(json.dumps(run(name, cls, cfg, dev=torch.device("cuda"), peaks=pk, **kw)), flush=True)
