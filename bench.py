"""
bench.py -- headline benchmark of the event_flow hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], "cfg2"): LIFFireNet, 5-bin voxel input, 128x128, batch 8 per GPU, one training window =
10 timesteps of 1000 events per sample: 10 forward passes + the EventWarping loss.  One "step" = that window.
metric = events/s over the whole job (all ranks), inputs resident in HBM (`value`) and end-to-end from pinned host event
lists (`e2e`: H2D of the raw events each timestep, device-side encoding, model, loss, D2H of the loss).
Also reported: the full train step (forward + loss + BPTT + gradient all-reduce + clip + Adam) under "train".
The reference arm (--impl reference) times the CPU restatement of the reference's path (oracle/, "port") on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 128
B_PER_GPU = 8
T = 10
N_EV = 1000
BINS = 5
LIF = dict(leak=[-4.0, 0.1], thresh=[0.8, 0.1], learn_leak=True, learn_thresh=True, hard_reset=True)
MODEL_CFG = dict(name="LIFFireNet", encoding="voxel", round_encoding=False, norm_input=False, num_bins=BINS, base_num_channels=32,
                 kernel_size=3, activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron=LIF)
LOSS_CFG = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False, "clip_grad": 100.0},
            "model": {"mask_output": True}}
WORKLOAD = "cfg2: LIFFireNet fwd x10 + EventWarping loss, 128x128, 5 voxel bins, 1000 ev/window, batch 8 per GPU"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def make_events(rank, step):
    """Raw event lists of one window: T tensors [B,N,4] (ts,y,x,p); seed 1234 + 1000*rank + t."""
    from oracle.encodings import synthetic_events

    out = []
    for t in range(T):
        ts, ys, xs, ps = synthetic_events(B_PER_GPU, N_EV, H, W, 1234 + 1000 * rank + 100 * step + t)
        out.append(torch.stack([ts, ys, xs, ps], dim=2))
    return out


def scale_weights(model):
    """Reference init (seed 0) with conv weights x2.5 so that spikes reach the prediction layer on 1000 ev / 128^2."""
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        model.pred.conv2d.weight.mul_(20.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# event-warping loss in isolation (BASELINE metric: "IWE warp Mevents/s"; SURVEY 8d: also on a large stream)
# ---------------------------------------------------------------------------------------------------------------------
def iwe_bench(dev, peak_gbs, reps=10):
    """
    Kernel time of ef_iwe_loss_fwd (+ ef_iwe_loss_bwd) through ops.event_warping_loss on device-resident windows: the cfg-2
    window (B=8, 128x128, T=10, 1000 events per pass) and the large stream of SURVEY 8d (B=32, 256x256, T=10, 50 000 events
    per pass = 16 M events).  The calls are captured in a CUDA graph (kernels + memsets only) and replayed between CUDA events.
    Algorithmic bytes per sample: fwd 32*Ntot + 4*HW*(2T*S + T + 16*S), bwd 32*Ntot + 4*HW*(8*S + 2T*S).
    """
    from event_flow_b200 import _lib as L

    out = {}
    for name, (B, Hh, Ww, Tt, N) in (("cfg2_B8_128x128_T10_N1000", (8, 128, 128, 10, 1000)), ("large_B32_256x256_T10_N50000", (32, 256, 256, 10, 50000))):
        g = torch.Generator(device="cpu").manual_seed(99)
        ntot = Tt * N
        ts = torch.rand((B, Tt, N), generator=g).sort(dim=2).values + torch.arange(Tt).view(1, Tt, 1)  # pass offset already added (loss/flow.py:90)
        ys = torch.randint(0, Hh, (B, Tt, N), generator=g).float()
        xs = torch.randint(0, Ww, (B, Tt, N), generator=g).float()
        ps = (torch.rand((B, Tt, N), generator=g) < 0.5).float() * 2 - 1
        events = torch.stack([ts, ys, xs, ps], dim=3).reshape(B, ntot, 4).to(dev)
        pol = torch.stack([(ps > 0).float(), (ps < 0).float()], dim=3).reshape(B, ntot, 2).to(dev)
        flow = ((torch.rand((1, B, Tt, 2, Hh, Ww), generator=g) - 0.5) * 0.008).to(dev)
        mask = (torch.rand((B, Tt, Hh, Ww), generator=g) < 0.3).float().to(dev)
        del ts, ys, xs, ps
        p = L.IweLossParams()
        p.S, p.B, p.T, p.T_maps, p.H, p.W = 1, B, Tt, Tt, Hh, Ww
        p.n_total, p.n_per_pass = ntot, N
        p.flow_scaling, p.weight = float(max(Hh, Ww)), 0.001
        p.loss_scaling, p.smoothing_mask, p.overwrite_intermediate = 1, 1, 0
        ws = torch.empty(L.lib().ef_iwe_loss_workspace_elems(1, B, Hh, Ww), device=dev, dtype=torch.float32)
        loss = torch.empty((), device=dev)
        g_loss = torch.ones((), device=dev)
        g_maps = torch.empty_like(flow)
        p.events, p.pol_mask, p.flow_maps, p.event_mask = L.ptr(events), L.ptr(pol), L.ptr(flow), L.ptr(mask)
        p.workspace, p.loss, p.g_loss, p.g_flow_maps = L.ptr(ws), L.ptr(loss), L.ptr(g_loss), L.ptr(g_maps)

        def timed_graph(fn):
            fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, capture_error_mode="thread_local"):
                fn()
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        ms_f = timed_graph(lambda: L.call("ef_iwe_loss_fwd", p))
        ms_fb = timed_graph(lambda: (L.call("ef_iwe_loss_fwd", p), L.call("ef_iwe_loss_bwd", p)))
        hw = Hh * Ww
        bytes_f = B * (32 * ntot + 4 * hw * (2 * Tt + Tt + 16))
        bytes_b = B * (32 * ntot + 4 * hw * (8 + 2 * Tt))
        out[name] = {"events": B * ntot, "fwd_ms": ms_f, "fwd_bwd_ms": ms_fb, "fwd_Mev_s": B * ntot / ms_f / 1e3, "fwd_bwd_Mev_s": B * ntot / ms_fb / 1e3,
                     "fwd_GBs_algorithmic": bytes_f / ms_f / 1e6, "fwd_frac_of_hbm_peak": bytes_f / ms_f / 1e6 / peak_gbs,
                     "fwd_bwd_GBs_algorithmic": (bytes_f + bytes_b) / ms_fb / 1e6, "fwd_bwd_frac_of_hbm_peak": (bytes_f + bytes_b) / ms_fb / 1e6 / peak_gbs,
                     "loss": float(loss.item())}
        del events, pol, flow, mask, ws, g_maps
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from event_flow_b200 import _lib
    from event_flow_b200.dataloader.encodings import encode_batch
    from event_flow_b200.loss.flow import EventWarping
    from event_flow_b200.models.model import LIFFireNet
    from event_flow_b200.parallel import DataParallelTrainer

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert _lib.lib().ef_device_ok() == 1, "this benchmark needs a compute-capability 10.x GPU (B200)"

    torch.manual_seed(0)
    model = LIFFireNet(MODEL_CFG)
    scale_weights(model)
    model = model.to(dev).train()
    lossf = EventWarping(LOSS_CFG, dev)
    trainer = DataParallelTrainer(model, lr=2e-4, clip_grad=100.0)

    n_windows = 4  # distinct input windows cycled through (device-resident for `value`, pinned host for `e2e`)
    host = [[e.pin_memory() for e in make_events(rank, s)] for s in range(n_windows)]
    resident = []
    for win in host:
        enc = []
        for e in win:
            ed = e.to(dev)
            d = encode_batch(ed, (H, W), BINS)
            enc.append((d["event_voxel"], d["event_cnt"], ed, d["event_list_pol_mask"], d["event_mask"]))
        resident.append(enc)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def fwd_loss_resident(k):
        model.reset_states()
        lossf.reset()
        with torch.no_grad():
            for vox, cnt, ev, pm, mask in resident[k % n_windows]:
                out = model(vox, cnt)
                lossf.event_flow_association(out["flow"], ev.clone(), pm, mask)
            return lossf()

    def fwd_loss_e2e(k):
        model.reset_states()
        lossf.reset()
        with torch.no_grad():
            for e in host[k % n_windows]:
                ed = e.to(dev, non_blocking=True)
                d = encode_batch(ed, (H, W), BINS)
                out = model(d["event_voxel"], d["event_cnt"])
                lossf.event_flow_association(out["flow"], ed, d["event_list_pol_mask"], d["event_mask"])
            return lossf().item()  # D2H read of the result

    def train_step(k):
        model.reset_states()
        lossf.reset()
        for vox, cnt, ev, pm, mask in resident[k % n_windows]:
            out = model(vox, cnt)
            lossf.event_flow_association(out["flow"], ev.clone(), pm, mask)
        loss = lossf()
        loss.backward()
        trainer.step()
        model.detach_states()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for k in range(warmup):
            fn(k)
        barrier()
        evs = []
        n0 = _lib.lib().ef_launch_count() + _lib.GRAPH_KERNELS
        for k in range(steps):
            flush.fill_(k & 0xFF)  # L2 flush between timed iterations (outside the timed events)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(warmup + k)
            e1.record()
            evs.append((e0, e1))
        barrier()
        launches = _lib.lib().ef_launch_count() + _lib.GRAPH_KERNELS - n0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps, launches // steps

    events_per_step = B_PER_GPU * T * N_EV * world
    with ClockSampler(local) as clocks:
        ms_value, launches = timed(fwd_loss_resident, args.steps, args.warmup)
        ms_e2e, _ = timed(fwd_loss_e2e, args.steps, args.warmup)
        ms_train, launches_train = timed(train_step, max(2, args.steps // 2), max(3, args.warmup // 2))

    # roofline of the dominant kernel: the fused conv3x3+LIF step of the six 32->32 hidden layers.  The launches of a
    # whole window (T steps x 6 layers, real operands of this workload, 59 MB each: far more than L2 per replay) are
    # captured in one CUDA graph -- exactly how the model path issues them -- and the replay is timed with CUDA events on
    # the launching stream: average launch duration = replay time / launches (inter-kernel gaps included).
    from event_flow_b200 import fast

    xs = [enc[0] for enc in resident[0]] + [resident[1][0][0]]  # T + 1 inputs: step 0 only provides the previous state
    graph, n_hidden = fast.capture_window(model, xs, only_hidden=True)
    graph_all, n_all = fast.capture_window(model, xs, only_hidden=False)

    def replay_ms(gr, reps):
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    with ClockSampler(local) as clocks_k:
        ms_hidden = replay_ms(graph, max(10, args.steps))
        ms_all = replay_ms(graph_all, max(10, args.steps))
    del graph, graph_all
    pk, pk_kind = peaks()
    bytes_per_launch = 4 * H * W * (32 + 2 * 2 * 32) * B_PER_GPU  # 4*HW*(Cin + 2*S_r*C) per sample, fp32 reference semantics
    avg_ms = ms_hidden / n_hidden
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "fused conv3x3+LIF step, 32->32 ch (lif_conv_fwd_tc_kernel via ef_lif_conv_fwd)", "achieved": achieved,
                "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": 34274304, "peak_source": pk_kind + " (burst copy)",
                "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms, "launches_per_step": n_hidden,
                "traffic_source": "ncu --set full, profiles/r01d_ncu_tc_fwd_v5.txt: dram__bytes_read.sum 33 643 008 (= x 8.4 + z 8.4 + v 16.8 MB, exactly compulsory) + "
                                  "dram__bytes_write.sum 631 296 (the 25.2 MB of outputs are still in the 126 MB L2 when the kernel ends); fast-path formats move 58.7 MB per launch",
                "how": f"{n_hidden} launches (4 feed-forward + 2 recurrent cells x {T} steps) replayed as one CUDA graph, CUDA events around the replays",
                "share_of_step": ms_hidden / ms_all, "model_kernels_ms_per_window": ms_all, "clocks": clocks_k.summary()}

    iwe = iwe_bench(dev, pk["hbm_gbs"]) if rank == 0 else None
    if rank == 0:
        # the CPU baseline is taken at N=1 only: under torchrun the other ranks spin in the barrier and steal the host cores
        cpu = cpu_baseline(sample_steps=1) if world == 1 else None
        line = {
            "metric": "events/s (LIFFireNet fwd + EventWarping loss)", "value": events_per_step / (ms_value * 1e-3), "unit": "events/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B_PER_GPU * world, "timesteps": T, "events_per_window": N_EV, "resolution": [H, W],
                       "parallelism": f"dp{world}", "l2": "256 MB write between timed iterations (L2 flush)", "weights": "reference init seed 0, conv x2.5"},
            "e2e": {"value": events_per_step / (ms_e2e * 1e-3), "unit": "events/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": T * B_PER_GPU * N_EV * 4 * 4, "d2h_bytes_per_step": 4,
                    "path": "pinned host event lists -> H2D -> ef_encode_events -> LIFFireNet x10 -> EventWarping -> loss.item()"},
            "gpu_launches": int(launches),
            "train": {"ms_per_step": ms_train, "events_per_s": events_per_step / (ms_train * 1e-3), "gpu_launches": int(launches_train),
                      "what": "fwd x10 + loss + BPTT + grad all-reduce(SUM) + clip(100) + Adam, batch 8 per GPU"},
            "roofline": roofline, "iwe": iwe, "cpu_baseline": cpu, "clocks": clocks.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle restatement of the reference's path on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_step(params, windows):
    from oracle import iwe as oiwe
    from oracle import spiking as osp

    states = [None] * 7
    flows, evs, pms, masks = [], [], [], []
    with torch.no_grad():
        for t, d in enumerate(windows):
            flow, states, _ = osp.firenet_step("lif", params, states, d["event_voxel"])
            flows.append(flow)
            e = d["event_list"].clone()
            e[:, :, 0] += t
            evs.append(e), pms.append(d["event_list_pol_mask"]), masks.append(d["event_mask"])
        return oiwe.event_warping_loss(torch.cat(evs, 1), torch.cat(pms, 1), torch.arange(T).repeat_interleave(N_EV), [torch.stack(flows, 1)],
                                       torch.cat(masks, 1), (H, W), weight=0.001, passes=T)


def cpu_setup():
    from oracle import encodings as oenc
    from oracle import spiking as osp

    cores = os.cpu_count()
    torch.set_num_threads(cores)
    params = osp.init_firenet_params("lif", BINS, 32, seed=0, weight_gain=2.5)
    params["pred"]["weight"] = params["pred"]["weight"] * 20.0
    windows = []
    for e in make_events(0, 0):
        windows.append(oenc.encode_window(e[:, :, 0], e[:, :, 1], e[:, :, 2], e[:, :, 3], H, W, BINS))
    return cores, params, windows


def cpu_baseline(sample_steps=1):
    cores, params, windows = cpu_setup()
    cpu_step(params, windows)  # warm-up
    t0 = time.perf_counter()
    for _ in range(sample_steps):
        cpu_step(params, windows)
    dt = (time.perf_counter() - t0) / sample_steps
    return {"value": B_PER_GPU * T * N_EV / dt, "unit": "events/s", "cores": cores, "kind": "port", "ms_per_step": dt * 1e3,
            "sample": f"{sample_steps} full window(s) of the same workload (batch 8, 10 timesteps) after 1 warm-up, torch-CPU fp32 oracle, {cores} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores, params, windows = cpu_setup()
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_step(params, windows)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(params, windows)
    dt = (time.perf_counter() - t0) / steps
    v = B_PER_GPU * T * N_EV / dt
    sample = f"{steps} full windows (batch 8, 10 timesteps; one rank's share) after warm-up, torch-CPU fp32 oracle port of the reference, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "events/s (LIFFireNet fwd + EventWarping loss)", "value": v, "unit": "events/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "global_batch": B_PER_GPU, "parallelism": "cpu"},
        "cpu_baseline": {"value": v, "unit": "events/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
