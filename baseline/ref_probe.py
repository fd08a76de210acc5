"""
Survey-time baseline probe (measurement only, no product code).

Imports the UNMODIFIED reference modules (tudelft/event_flow) and times its two hot paths -- the recurrent
(spiking) conv model and the event-warping loss -- on this box's host CPU and, if present, on its GPU through
the reference's own stock-PyTorch path. Also measures how far the reference's own CUDA path (TF32 on / off)
drifts from its CPU path, which bounds what "parity" can mean for a spiking network.

Run:  python baseline/ref_probe.py            (CPU only)
      gpurun -- python baseline/ref_probe.py  (GPU box)
Writes gpurun_out/ref_probe.json.
"""

import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference" if os.path.isdir("/root/reference/models") else os.path.join(HERE, "_ref")
sys.path.insert(0, REF)

from dataloader.encodings import events_to_channels, events_to_voxel  # noqa: E402
from loss.flow import EventWarping  # noqa: E402
from models.model import ALIFFireNet, FireNet, LIFFireNet, PLIFFireNet, SpikingRecEVFlowNet  # noqa: E402
from models.spiking_submodules import ConvLIFRecurrent  # noqa: E402

OUT = {"ref_path": REF, "ref_at_root_reference": os.path.isdir("/root/reference/models")}
LIF = dict(leak=[-4.0, 0.1], thresh=[0.8, 0.1], learn_leak=True, learn_thresh=True, hard_reset=True)


def log(*a):
    print(*a, flush=True)


def gen_events(B, N, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    ts = torch.sort(torch.rand(B, N, generator=g))[0]
    ts = (ts - ts[:, :1]) / (ts[:, -1:] - ts[:, :1])
    ys = torch.randint(0, H, (B, N), generator=g).float()
    xs = torch.randint(0, W, (B, N), generator=g).float()
    ps = (torch.randint(0, 2, (B, N), generator=g) * 2 - 1).float()
    return ts, ys, xs, ps


def encode(ts, ys, xs, ps, H, W, bins):
    B = ts.shape[0]
    vox = torch.stack([events_to_voxel(xs[b], ys[b], ts[b], ps[b], bins, sensor_size=(H, W)) for b in range(B)])
    cnt = torch.stack([events_to_channels(xs[b], ys[b], ps[b], sensor_size=(H, W)) for b in range(B)])
    ev = torch.stack([ts, ys, xs, ps], 2)
    pm = torch.stack([(ps > 0).float(), (ps < 0).float()], 2)
    mask = (cnt.sum(1, keepdim=True) > 0).float()
    return vox, cnt, ev, pm, mask


def build(cls, bins, encoding, sn, acts, seed=0, scale=2.0):
    torch.manual_seed(seed)
    cfg = dict(name="x", encoding=encoding, round_encoding=False, norm_input=False, num_bins=bins, base_num_channels=32,
               kernel_size=3, activations=acts, mask_output=True, spiking_neuron=sn)
    if hasattr(cls, "kwargs"):
        cls.kwargs = [{}] * 7  # reference shares one class-level dict between all FireNet subclasses
    m = cls(cfg)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(scale)  # keeps spikes alive through the 7 layers with random init
    return m


def sync(dev):
    if dev.type == "cuda":
        torch.cuda.synchronize()


def best_of(fn, dev, warm, reps):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        sync(dev)
        t0 = time.perf_counter()
        r = fn()
        sync(dev)
        ts.append(time.perf_counter() - t0)
    return min(ts), r


def loss_cfg(H, W):
    return {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False},
            "model": {"mask_output": True}}


def bench_device(dev, tag):
    res = {}
    cpu = dev.type == "cpu"
    warm, reps = (1, 3) if cpu else (3, 10)
    H = W = 128

    # cfg1: FireNet ANN forward, 1k events, 1 bin, B=1
    m = build(FireNet, 1, "voxel", None, ["relu", None]).to(dev).eval()
    vox, cnt, *_ = encode(*gen_events(1, 1000, H, W, 0), H, W, 1)
    vox, cnt = vox.to(dev), cnt.to(dev)
    with torch.no_grad():
        t, _ = best_of(lambda: m(vox, cnt), dev, warm + 2, reps * 3)
    res["cfg1_firenet_fwd_ms"] = t * 1e3
    log(tag, "cfg1 FireNet fwd B=1: %.3f ms" % (t * 1e3))

    # cfg2/3: LIFFireNet 5 voxel bins, T=10, B=8 : fwd, loss, bwd
    for name, cls, sn, T in (("lif", LIFFireNet, LIF, 10), ("plif", PLIFFireNet, {}, 20), ("alif", ALIFFireNet, {}, 20)):
        B = 8
        m = build(cls, 5, "voxel", sn, ["arctanspike", "arctanspike"]).to(dev).train()
        L = EventWarping(loss_cfg(H, W), dev)
        data = [[x.to(dev) for x in encode(*gen_events(B, 1000, H, W, 100 + t), H, W, 5)] for t in range(T)]
        parts = {}

        def step():
            m.reset_states()
            L.reset()
            m.zero_grad()
            sync(dev)
            t0 = time.perf_counter()
            for vox, cnt, ev, pm, mk in data:
                out = m(vox, cnt)
                L.event_flow_association(out["flow"], ev.clone(), pm, mk)
            sync(dev)
            t1 = time.perf_counter()
            loss = L()
            sync(dev)
            t2 = time.perf_counter()
            loss.backward()
            sync(dev)
            t3 = time.perf_counter()
            parts.setdefault("fwd", []).append(t1 - t0)
            parts.setdefault("loss", []).append(t2 - t1)
            parts.setdefault("bwd", []).append(t3 - t2)
            return float(loss)

        for _ in range(warm):
            step()
        parts.clear()
        for _ in range(reps):
            lv = step()
        r = {k: min(v) * 1e3 for k, v in parts.items()}
        r["total"] = r["fwd"] + r["loss"] + r["bwd"]
        r["Mev_per_s_train"] = B * T * 1000 / r["total"] / 1e3
        r["loss_value"] = lv
        res["%sfirenet_B8_T%d_ms" % (name, T)] = r
        log(tag, "%sFireNet B=8 T=%d: fwd %.1f ms | loss %.2f ms | bwd %.1f ms | total %.1f ms (%.3f Mev/s) loss=%.5f"
            % (name.upper(), T, r["fwd"], r["loss"], r["bwd"], r["total"], r["Mev_per_s_train"], lv))
        if name == "lif" and not cpu:
            try:
                from torch.profiler import ProfilerActivity, profile
                with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
                    step()
                    sync(dev)
                ka = prof.key_averages()
                kern = [e for e in ka if getattr(e, "device_type", None) is not None and "cuda" in str(e.device_type).lower()]
                res["lif_train_step_cuda_kernel_launches"] = int(sum(e.count for e in kern))
                dt = lambda e: getattr(e, "self_device_time_total", getattr(e, "self_cuda_time_total", 0))
                res["lif_train_step_cuda_kernel_time_ms"] = sum(dt(e) for e in kern) / 1e3
                top = sorted(kern, key=lambda e: -dt(e))[:12]
                res["lif_train_step_top_kernels"] = [(e.key[:70], e.count, round(dt(e) / 1e3, 3)) for e in top]
                log(tag, "LIF train step: %d CUDA kernel launches, %.2f ms summed kernel time"
                    % (res["lif_train_step_cuda_kernel_launches"], res["lif_train_step_cuda_kernel_time_ms"]))
                for k in res["lif_train_step_top_kernels"]:
                    log("    ", k)
            except Exception as e:  # profiler is best-effort
                log(tag, "profiler failed:", repr(e))

    # IWE loss alone, fwd + bwd wrt flow: small (config) and large streams
    for (B, Hh, N, T) in ((8, 128, 1000, 10), (32, 256, 50000, 10)):
        if cpu and B == 32:
            B = 4  # keep the CPU run short; scale linearly
        L = EventWarping(loss_cfg(Hh, Hh), dev)
        evs = [[x.to(dev) for x in encode(*gen_events(B, N, Hh, Hh, 500 + t), Hh, Hh, 2)[2:]] for t in range(T)]
        flows = [torch.empty(B, 2, Hh, Hh, device=dev).uniform_(-0.05, 0.05).requires_grad_() for _ in range(T)]
        parts = {}

        def iwe():
            L.reset()
            for (ev, pm, mk), fl in zip(evs, flows):
                L.event_flow_association([fl], ev.clone(), pm, mk)
            sync(dev)
            t0 = time.perf_counter()
            loss = L()
            sync(dev)
            t1 = time.perf_counter()
            loss.backward()
            sync(dev)
            t2 = time.perf_counter()
            parts.setdefault("fwd", []).append(t1 - t0)
            parts.setdefault("bwd", []).append(t2 - t1)

        for _ in range(warm):
            iwe()
        parts.clear()
        for _ in range(reps):
            iwe()
        f, b = min(parts["fwd"]), min(parts["bwd"])
        nev = B * N * T
        key = "iwe_B%d_%d_N%d_T%d" % (B, Hh, N, T)
        res[key] = {"fwd_ms": f * 1e3, "bwd_ms": b * 1e3, "Mev_s_fwd": nev / f / 1e6, "Mev_s_fwdbwd": nev / (f + b) / 1e6}
        log(tag, key, "fwd %.3f ms bwd %.3f ms -> %.1f Mev/s fwd, %.1f Mev/s fwd+bwd" % (f * 1e3, b * 1e3, nev / f / 1e6, nev / (f + b) / 1e6))

    # cfg4: SpikingRecEVFlowNet 256x256, 50k events/window, B=4 per GPU: fwd, and train step with T=2 on GPU
    Hh = 256
    B = 4
    m = build(SpikingRecEVFlowNet, 2, "cnt", {}, ["arctanspike", "arctanspike"], scale=1.0).to(dev).train()
    vox, cnt, ev, pm, mk = [x.to(dev) for x in encode(*gen_events(B, 50000, Hh, Hh, 7), Hh, Hh, 2)]

    def fwd4():
        with torch.no_grad():
            return m(vox, cnt)

    t, _ = best_of(fwd4, dev, warm, reps)
    res["cfg4_snn_evflownet_fwd_B4_ms"] = t * 1e3
    log(tag, "cfg4 SpikingRecEVFlowNet 256x256 B=4 fwd: %.2f ms/step" % (t * 1e3))
    if not cpu:
        L = EventWarping(loss_cfg(Hh, Hh), dev)

        def train4():
            m.reset_states()
            L.reset()
            m.zero_grad()
            for _ in range(2):
                out = m(vox, cnt)
                L.event_flow_association(out["flow"], ev.clone(), pm, mk)
            loss = L()
            loss.backward()
            return float(loss)

        t, lv = best_of(train4, dev, 2, 5)
        res["cfg4_snn_evflownet_train_B4_T2_ms"] = t * 1e3
        log(tag, "cfg4 train step (T=2, 4 flow scales): %.2f ms loss=%.5f" % (t * 1e3, lv))
    return res


def drift_vs_cpu(dev):
    """How far is the reference's own CUDA path from its CPU path? (teacher-forced step and free rollout)"""
    res = {}
    B, C, H, W = 8, 32, 128, 128
    torch.manual_seed(0)
    cell = ConvLIFRecurrent(C, C, 3, leak=(-4.0, 0.1), thresh=(0.8, 0.1))
    with torch.no_grad():
        cell.ff.weight.mul_(2)
        cell.rec.weight.mul_(2)
        x = (torch.rand(B, C, H, W) < 0.3).float()
        st = torch.stack([torch.randn(B, C, H, W) * 0.5, (torch.rand(B, C, H, W) < 0.3).float()])
        _, ns_cpu = cell(x, st)
        gcell = ConvLIFRecurrent(C, C, 3).to(dev)
        gcell.load_state_dict(cell.state_dict())
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            _, ns = gcell(x.to(dev), st.to(dev))
            ns = ns.cpu()
            flips = int((ns[1] != ns_cpu[1]).sum())
            dv = float((ns[0] - ns_cpu[0]).abs().max())
            res["cell_step_tf32_%s" % tf32] = {"spike_flips": flips, "of": ns[1].numel(), "max_abs_dv": dv}
            log("drift: teacher-forced ConvLIFRecurrent step, cudnn.allow_tf32=%s: %d flips of %d, max|dv| %.3e" % (tf32, flips, ns[1].numel(), dv))
    # free-running rollout LIFFireNet, CUDA(tf32 off/on) vs CPU
    H = W = 128
    Bm = 2
    mc = build(LIFFireNet, 2, "cnt", LIF, ["arctanspike", "arctanspike"])
    data = [encode(*gen_events(Bm, 1000, H, W, 900 + t), H, W, 2) for t in range(10)]
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        mg = build(LIFFireNet, 2, "cnt", LIF, ["arctanspike", "arctanspike"]).to(dev)
        mg.load_state_dict(mc.state_dict())
        mc.reset_states()
        rows = []
        with torch.no_grad():
            for t, (vox, cnt, ev, pm, mk) in enumerate(data):
                fc = mc(vox, cnt)["flow"][0]
                fg = mg(vox.to(dev), cnt.to(dev))["flow"][0].cpu()
                mism = [float((a[1] != b[1].cpu()).float().mean()) for a, b in zip(mc._states, mg._states)]
                rows.append({"t": t, "flow_max_rel": float((fc - fg).abs().max() / (fc.abs().max() + 1e-12)),
                             "flow_mean_abs": float((fc - fg).abs().mean()), "spike_mismatch_last_layer": mism[-1]})
        res["rollout_tf32_%s" % tf32] = rows
        log("drift: LIFFireNet free rollout CUDA(tf32=%s) vs CPU:" % tf32,
            " ".join("t%d:%.1e/%.1e" % (r["t"], r["flow_max_rel"], r["spike_mismatch_last_layer"]) for r in rows), "(flow max-rel / last-layer spike mismatch)")
    torch.backends.cudnn.allow_tf32 = True
    return res


def main():
    OUT["torch"] = torch.__version__
    OUT["cpu_count"] = os.cpu_count()
    OUT["torch_threads"] = torch.get_num_threads()
    OUT["cuda"] = torch.cuda.is_available()
    log("ref:", REF, "| torch", torch.__version__, "| cpu_count", os.cpu_count(), "| torch threads", torch.get_num_threads())
    try:
        with open("/proc/cpuinfo") as f:
            names = [line.split(":")[1].strip() for line in f if line.startswith("model name")]
        OUT["cpu_model"] = names[0] if names else None
        log("cpu:", OUT["cpu_model"], "x", len(names))
    except Exception:
        pass
    if OUT["cuda"]:
        OUT["gpu"] = torch.cuda.get_device_name(0)
        OUT["gpu_count"] = torch.cuda.device_count()
        OUT["cudnn_allow_tf32_default"] = torch.backends.cudnn.allow_tf32
        OUT["matmul_allow_tf32_default"] = torch.backends.cuda.matmul.allow_tf32
        log("gpu:", OUT["gpu"], "x", OUT["gpu_count"], "| cudnn.allow_tf32 default", OUT["cudnn_allow_tf32_default"], "| cudnn", torch.backends.cudnn.version())
        for key, fn in (("cuda_bench", lambda: bench_device(torch.device("cuda:0"), "[cuda]")),
                        ("drift", lambda: drift_vs_cpu(torch.device("cuda:0")))):
            try:
                OUT[key] = fn()
            except Exception:  # keep going: the CPU baseline below is still wanted
                import traceback
                OUT[key + "_error"] = traceback.format_exc()
                log(OUT[key + "_error"])
    torch.set_num_threads(os.cpu_count())
    OUT["cpu_bench"] = bench_device(torch.device("cpu"), "[cpu %d thr]" % torch.get_num_threads())
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ref_probe.json", "w") as f:
        json.dump(OUT, f, indent=1)
    log("wrote gpurun_out/ref_probe.json")


if __name__ == "__main__":
    main()
