"""
bench.py -- headline benchmark of the event_flow hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload: one rank's share of BASELINE.json configs[2] ("cfg3"): LIFFireNet (configs/train_SNN.yml), 5-bin voxel input, 128x128,
batch 8 per GPU (global batch 8 N), one truncated-BPTT window = 10 timesteps of 1000 events per sample.  One "step" = one
TRAINING window: 10 forward passes + the EventWarping loss + BPTT through the 10 steps + ONE all-reduce(SUM) of the flat
gradient over NCCL + gradient-norm clip + Adam (train_flow.py:98-171).  The neuron state carries over from window to window
(detach_states at the window boundary, no reset), like the reference's loop.
    value  = events/s of the whole job (all ranks), encoded inputs resident in HBM, timed on the device
    e2e    = the same from pinned host event lists: H2D of the raw events every timestep, device-side encoding, model, loss,
             backward, all-reduce, optimiser, D2H of the loss value
    fwd_loss = configs[1] ("cfg2"): the forward passes + loss only (no gradient), resident and end to end -- the inference figure
    roofline = the dominant kernel (fused conv3x3 + LIF step), iwe = the event-warping loss in isolation (incl. a 16 M-event stream)
The reference arm (--impl reference) runs the UNMODIFIED reference (staged in baseline/_ref by tools/stage_reference.py)
through its own public API on the host cores: the same training window, torch-CPU.
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
LR, CLIP = 2e-4, 100.0
W_GAIN, PRED_GAIN = 2.5, 20.0
LIF = dict(leak=[-4.0, 0.1], thresh=[0.8, 0.1], learn_leak=True, learn_thresh=True, hard_reset=True)
MODEL_CFG = dict(name="LIFFireNet", encoding="voxel", round_encoding=False, norm_input=False, num_bins=BINS, base_num_channels=32,
                 kernel_size=3, activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron=LIF)
LOSS_CFG = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False, "clip_grad": CLIP},
            "model": {"mask_output": True}}
WORKLOAD = ("cfg3 (one rank's share): LIFFireNet BPTT train step = fwd x10 + EventWarping loss + backward + grad all-reduce + clip + Adam, "
            "128x128, 5 voxel bins, 1000 ev/window, batch 8 per GPU")
WORKLOAD_FWD = "cfg2: LIFFireNet fwd x10 + EventWarping loss (no gradient), 128x128, 5 voxel bins, 1000 ev/window, batch 8 per GPU"
METRIC = "events/s (LIFFireNet BPTT train step)"
WEIGHTS = (f"reference init (seed 0); conv weights x{W_GAIN}, prediction weights x{PRED_GAIN} so that spikes reach the prediction layer at "
           "1000 events / 128^2 (default init dies at layer G2, SURVEY 8d; BASELINE.md suggests x2 -- dense kernels, cost independent of the values)")


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def synthetic_events(B, N, seed):
    """x,y uniform integer pixels (as fp32), p = +-1, ts sorted uniform normalised to [0,1] (dataloader/base.py:84-85)."""
    g = torch.Generator().manual_seed(seed)
    ts = torch.sort(torch.rand(B, N, generator=g))[0]
    ts = (ts - ts[:, :1]) / (ts[:, -1:] - ts[:, :1])
    ys = torch.randint(0, H, (B, N), generator=g).float()
    xs = torch.randint(0, W, (B, N), generator=g).float()
    ps = (torch.randint(0, 2, (B, N), generator=g) * 2 - 1).float()
    return torch.stack([ts, ys, xs, ps], dim=2)


def make_events(rank, step):
    """Raw event lists of one window: T tensors [B,N,4] (ts,y,x,p); seed 1234 + 1000*rank + 100*step + t."""
    return [synthetic_events(B_PER_GPU, N_EV, 1234 + 1000 * rank + 100 * step + t) for t in range(T)]


def scale_weights(model):
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(W_GAIN)
        model.pred.conv2d.weight.mul_(PRED_GAIN)


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
    Kernel time of ef_iwe_loss_fwd (+ ef_iwe_loss_bwd): ONE launch each, on device-resident windows: the cfg-2 window (B=8,
    128x128, T=10, 1000 events per pass) and the large stream of SURVEY 8d (B=32, 256x256, T=10, 50 000 events per pass =
    16 M events).  The calls are captured in a CUDA graph and replayed between CUDA events.
    Algorithmic bytes per sample: fwd 32*Ntot + 4*HW*(2T*S + T + 16*S), bwd 32*Ntot + 4*HW*(8*S + 2T*S).
    """
    from event_flow_b200 import _lib as L
    from event_flow_b200 import fast

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
        mask = torch.zeros((B, Tt, Hh * Ww))  # event mask of every pass: pixels that received an event (dataloader/base.py:159-172)
        mask.scatter_(2, (ys * Ww + xs).long(), 1.0)
        mask = mask.view(B, Tt, Hh, Ww).to(dev)
        del ts, ys, xs, ps
        p = L.IweLossParams()
        p.S, p.B, p.T, p.T_maps, p.H, p.W = 1, B, Tt, Tt, Hh, Ww
        p.n_total, p.n_per_pass = ntot, N
        p.flow_scaling, p.weight = float(max(Hh, Ww)), 0.001
        p.loss_scaling, p.smoothing_mask, p.overwrite_intermediate = 1, 1, 0
        ws = torch.zeros(L.lib().ef_iwe_loss_workspace_elems(1, B, Hh, Ww), device=dev, dtype=torch.float32)
        loss = torch.empty((), device=dev)
        g_loss = torch.ones((), device=dev)
        g_maps = torch.empty_like(flow)
        p.events, p.pol_mask, p.flow_maps, p.event_mask = L.ptr(events), L.ptr(pol), L.ptr(flow), L.ptr(mask)
        p.workspace, p.loss, p.g_loss, p.g_flow_maps = L.ptr(ws), L.ptr(loss), L.ptr(g_loss), L.ptr(g_maps)

        def timed_graph(fn):
            fn()
            torch.cuda.synchronize()
            gr = fast._capture(fn)  # (garbage collector paused inside the capture)
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
        out[name] = {"events": B * ntot, "launches_fwd": 1, "launches_bwd": 1, "fwd_ms": ms_f, "fwd_bwd_ms": ms_fb,
                     "fwd_Mev_s": B * ntot / ms_f / 1e3, "fwd_bwd_Mev_s": B * ntot / ms_fb / 1e3,
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

    from event_flow_b200 import _lib, fast
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

    dp_check = dp_gradient_check(model, lossf, rank, world, dev) if world > 1 else None
    trainer = DataParallelTrainer(model, lr=LR, clip_grad=CLIP)
    trainer_fused = trainer.fused is not None

    # every timed / warm-up window has its own inputs (event_flow_association offsets the timestamps IN PLACE on the caller's
    # tensor, loss/flow.py:90, so a window's event tensors are consumed by one use -- exactly like batches from a loader)
    n_per_leg = args.warmup + args.steps

    def host_window(k):
        return torch.stack(make_events(rank, k)).pin_memory()  # [T,B,N,4]: the raw events of one window, staged by the loader

    def resident_window(k):
        ed = torch.stack(make_events(rank, k)).to(dev)
        d = encode_batch(ed.view(T * B_PER_GPU, N_EV, 4), (H, W), BINS)  # one launch for the T*B images of the window
        return (d["event_voxel"].view(T, B_PER_GPU, BINS, H, W), d["event_cnt"].view(T, B_PER_GPU, 2, H, W), ed,
                d["event_list_pol_mask"].view(T, B_PER_GPU, N_EV, 2), d["event_mask"].view(T, B_PER_GPU, 1, H, W))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # The model runs through its WINDOWED entry point (model.forward_window: the T forward passes of a loss window, layer-major with the
    # time loop inside the kernels of the feed-forward cells) -- what event_flow_b200.train.train_windows(staged=True) calls; the
    # per-step API model(voxel, cnt) of the reference is timed beside it (`stepwise` keys).
    def associate(outs, ed, pm, mask):
        for t in range(T):
            lossf.event_flow_association(outs[t]["flow"], ed[t], pm[t], mask[t])

    def stage_e2e(e):
        ed = e.to(dev, non_blocking=True)  # H2D of the window's raw events
        d = encode_batch(ed.view(T * B_PER_GPU, N_EV, 4), (H, W), BINS)
        return (d["event_voxel"].view(T, B_PER_GPU, BINS, H, W), d["event_cnt"].view(T, B_PER_GPU, 2, H, W), ed,
                d["event_list_pol_mask"].view(T, B_PER_GPU, N_EV, 2), d["event_mask"].view(T, B_PER_GPU, 1, H, W))

    def fwd_loss_resident(win):
        vox, cnt, ed, pm, mask = win
        lossf.reset()
        with torch.no_grad():
            associate(model.forward_window(vox, cnt), ed, pm, mask)
            return lossf()

    def fwd_loss_e2e(e):
        lossf.reset()
        with torch.no_grad():
            vox, cnt, ed, pm, mask = stage_e2e(e)
            associate(model.forward_window(vox, cnt), ed, pm, mask)
            return lossf().item()  # D2H read of the result

    def fwd_loss_stepwise(win):
        vox, cnt, ed, pm, mask = win
        lossf.reset()
        with torch.no_grad():
            for t in range(T):
                out = model(vox[t], cnt[t])
                lossf.event_flow_association(out["flow"], ed[t], pm[t], mask[t])
            return lossf()

    def train_resident(win):
        vox, cnt, ed, pm, mask = win
        lossf.reset()
        associate(model.forward_window(vox, cnt), ed, pm, mask)
        loss = lossf()
        loss.backward()
        trainer.step()
        model.detach_states()
        return loss

    def train_stepwise(win):
        vox, cnt, ed, pm, mask = win
        lossf.reset()
        for t in range(T):
            out = model(vox[t], cnt[t])
            lossf.event_flow_association(out["flow"], ed[t], pm[t], mask[t])
        loss = lossf()
        loss.backward()
        trainer.step()
        model.detach_states()
        return loss

    def train_e2e(e):
        lossf.reset()
        vox, cnt, ed, pm, mask = stage_e2e(e)
        associate(model.forward_window(vox, cnt), ed, pm, mask)
        loss = lossf()
        loss.backward()
        trainer.step()
        model.detach_states()
        return loss.item()  # D2H read of the result

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, windows, steps, warmup):
        model.reset_states()
        for k in range(warmup):
            fn(windows[k])
        barrier()
        evs = []
        n0 = _lib.lib().ef_launch_count() + _lib.GRAPH_KERNELS
        for k in range(steps):
            flush.fill_(k & 0xFF)  # L2 flush between timed iterations (outside the timed events)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(windows[warmup + k])
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
        legs = []
        for i, (fn, mk) in enumerate(((train_resident, resident_window), (train_e2e, host_window), (fwd_loss_resident, resident_window),
                                      (fwd_loss_e2e, host_window), (train_stepwise, resident_window), (fwd_loss_stepwise, resident_window))):
            wins = [mk(i * n_per_leg + k) for k in range(n_per_leg)]
            legs.append(timed(fn, wins, args.steps, args.warmup))
            del wins
        (ms_train, launches_train), (ms_train_e2e, _), (ms_fwd, launches_fwd), (ms_fwd_e2e, _), (ms_train_sw, launches_train_sw), (ms_fwd_sw, launches_fwd_sw) = legs

    # roofline of the dominant kernel: the fused conv3x3+LIF cell (lif_conv_fwd_tc_kernel).  In a window it is launched in two ways:
    #   * recurrent cells G1, G2: one launch per step (the recurrent convolution needs the neighbours' spikes of the previous step),
    #     T x 2 launches per window -- the largest share of the forward window, hence the `roofline` block;
    #   * feed-forward cells R1a, R1b, R2a, R2b (and the head on split inputs): ONE launch per window, time loop inside, state in registers.
    # Each population is captured as one CUDA graph on the workload's real operands (exactly how the model path issues them; every replay
    # streams far more than L2) and timed with CUDA events on the launching stream: average launch duration = replay time / launches.
    model.reset_states()
    vox0, vox1 = resident_window(0)[0], resident_window(1)[0]
    FF, REC = ("R1a", "R1b", "R2a", "R2b"), ("G1", "G2")
    g_rec, n_rec, _ = fast.capture_window_fused(model, vox0, vox1, only=REC)
    g_ff, n_ff, _ = fast.capture_window_fused(model, vox0, vox1, only=FF, save_all_v=False)
    g_ff_train, _, _ = fast.capture_window_fused(model, vox0, vox1, only=FF, save_all_v=True)
    g_head, _, _ = fast.capture_window_fused(model, vox0, vox1, only=("head",), save_all_v=False)
    g_pred, _, _ = fast.capture_window_fused(model, vox0, vox1, only=("pred",))
    g_all, n_all, _ = fast.capture_window_fused(model, vox0, vox1)

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

    reps = max(10, args.steps)
    with ClockSampler(local) as clocks_k:
        ms_rec, ms_ff, ms_ff_train = replay_ms(g_rec, reps), replay_ms(g_ff, reps), replay_ms(g_ff_train, reps)
        ms_head, ms_pred, ms_all = replay_ms(g_head, reps), replay_ms(g_pred, reps), replay_ms(g_all, reps)
    del g_rec, g_ff, g_ff_train, g_head, g_pred, g_all
    pk, pk_kind = peaks()
    px = H * W * B_PER_GPU
    cell_step_bytes = 4 * (32 + 2 * 2 * 32) * px  # 4*HW*(Cin + 2*S_r*C) per sample and cell-step, fp32 reference semantics
    avg_ms = ms_rec / n_rec
    achieved = cell_step_bytes / (avg_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic()
    gbs = lambda nbytes, ms: nbytes / (ms * 1e-3) / 1e9  # noqa: E731
    # bytes a launch really moves (bf16 channels-last spikes, fp32 membrane): recurrent step = x 64 + z_in 64 + v_in 128 + v_out 128 + z_out 64
    # per pixel; fused window launch = per step x 64 + z_out 64 (+ v_out 128 when the backward needs every step), state once per window
    moved_rec = 448 * px
    moved_ff, moved_ff_train = (128 * T + 448 - 128) * px, (256 * T + 448 - 128 - 128) * px
    roofline = {"bound": "hbm", "kernel": "fused conv3x3+LIF step, 32->32 ch, recurrent cell (lif_conv_fwd_tc_kernel<REC> via ef_lif_conv_fwd)",
                "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic,
                "peak_source": pk_kind + " (burst copy)", "bytes_per_launch": cell_step_bytes, "avg_launch_ms": avg_ms, "launches_per_step": n_rec,
                "traffic_source": traffic_src, "moved_bytes_per_launch": moved_rec, "frac_on_moved_bytes": gbs(moved_rec, avg_ms) / pk["hbm_gbs"],
                "how": f"{n_rec} launches (2 recurrent cells x {T} steps) replayed as one CUDA graph, CUDA events around the replays",
                "share_of_fwd_window": ms_rec / ms_all, "share_of_step": ms_rec / ms_train, "model_kernels_ms_per_window": ms_all,
                "clocks": clocks_k.summary(),
                "fused_window_launch": {
                    "kernel": "the same kernel, feed-forward cell over a whole window in one launch (ef_lif_conv_fwd_window): state in registers over T",
                    "launches_per_step": n_ff, "cell_steps_per_launch": T, "algorithmic_bytes_per_launch": cell_step_bytes * T,
                    "avg_launch_ms": ms_ff / n_ff, "achieved_algorithmic_GBs": gbs(cell_step_bytes * T, ms_ff / n_ff),
                    "frac_algorithmic": gbs(cell_step_bytes * T, ms_ff / n_ff) / pk["hbm_gbs"],
                    "note": "above 1 = faster than the HBM roofline of the step-by-step algorithm: the state bytes SURVEY 8d counts never move",
                    "moved_bytes_per_launch": moved_ff, "frac_on_moved_bytes": gbs(moved_ff, ms_ff / n_ff) / pk["hbm_gbs"],
                    "training": {"avg_launch_ms": ms_ff_train / n_ff, "moved_bytes_per_launch": moved_ff_train,
                                 "frac_algorithmic": gbs(cell_step_bytes * T, ms_ff_train / n_ff) / pk["hbm_gbs"],
                                 "frac_on_moved_bytes": gbs(moved_ff_train, ms_ff_train / n_ff) / pk["hbm_gbs"]},
                    "share_of_fwd_window": ms_ff / ms_all},
                "head_window_launch_ms": ms_head, "pred_window_launch_ms": ms_pred}

    iwe = iwe_bench(dev, pk["hbm_gbs"]) if rank == 0 else None
    other = None
    if rank == 0 and world == 1:  # BASELINE configs 4 and 5 (parity-test cases; timed for the record, N = 1 only)
        from tools import bench_configs

        del model, trainer
        torch.cuda.empty_cache()
        try:
            other = bench_configs.run_all(dev, pk, reps=2)
        except Exception as exc:  # the headline must not depend on the extra configurations
            other = {"error": repr(exc)}
    if rank == 0:
        # the CPU baseline is taken at N=1 only: under torchrun the other ranks spin in the barrier and steal the host cores
        cpu = cpu_baseline() if world == 1 else None
        line = {
            "metric": METRIC, "value": events_per_step / (ms_train * 1e-3), "unit": "events/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_train, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B_PER_GPU * world, "timesteps": T, "events_per_window": N_EV, "resolution": [H, W],
                       "parallelism": f"dp{world}", "collective": ("none (1 rank)" if world == 1 else
                                      "1 x ef_dp_step per rank and step: one-shot all-reduce(SUM, fp32, 74 818 elements) over NVLink peer memory fused with "
                                      "clip + Adam" if trainer_fused else "1 x ncclAllReduce(SUM, fp32, 74 818 elements) per step"),
                       "l2": "256 MB write between timed iterations (L2 flush)", "weights": WEIGHTS},
            "e2e": {"value": events_per_step / (ms_train_e2e * 1e-3), "unit": "events/s", "ms_per_step": ms_train_e2e,
                    "h2d_bytes_per_step": T * B_PER_GPU * N_EV * 4 * 4, "d2h_bytes_per_step": 4,
                    "path": "pinned host event lists of the window -> H2D -> ef_encode_events -> LIFFireNet.forward_window (10 steps) -> "
                            "EventWarping -> backward -> all-reduce -> clip + Adam -> loss.item()"},
            "gpu_launches": int(launches_train),
            "fwd_loss": {"workload": WORKLOAD_FWD, "value": events_per_step / (ms_fwd * 1e-3), "unit": "events/s", "ms_per_step": ms_fwd,
                         "steps": args.steps, "warmup": args.warmup, "gpu_launches": int(launches_fwd),
                         "algorithmic_GB_per_window": 73859072 * B_PER_GPU * T / 1e9,
                         "frac_of_hbm_roofline": 73859072 * B_PER_GPU * T / (ms_fwd * 1e-3) / 1e9 / pk["hbm_gbs"],
                         "e2e": {"value": events_per_step / (ms_fwd_e2e * 1e-3), "unit": "events/s", "ms_per_step": ms_fwd_e2e,
                                 "h2d_bytes_per_step": T * B_PER_GPU * N_EV * 4 * 4, "d2h_bytes_per_step": 4}},
            "train": {"ms_per_step": ms_train, "events_per_s": events_per_step / (ms_train * 1e-3), "gpu_launches": int(launches_train)},
            "stepwise": {"what": "the same windows through the reference's per-step call model(voxel, cnt) (one CUDA graph of 8 kernels per step)",
                         "train_ms_per_step": ms_train_sw, "train_gpu_launches": int(launches_train_sw),
                         "fwd_loss_ms_per_step": ms_fwd_sw, "fwd_loss_gpu_launches": int(launches_fwd_sw)},
            "dp_check": dp_check,
            "roofline": roofline, "iwe": iwe, "other_configs": other, "cpu_baseline": cpu, "clocks": clocks.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full summary."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(path))
        return int(d["dram_bytes_read"] + d["dram_bytes_write"]), f"profiles/ncu_traffic.json ({d.get('source', 'ncu --set full')})"
    except Exception:
        return None, "no committed ncu capture found (profiles/ncu_traffic.json)"


def dp_gradient_check(model, lossf, rank, world, dev):
    """
    Hardware self-check of the data-parallel step (train_flow.py:141-171 semantics): the all-reduced SUM of the ranks'
    batch-8 gradients must equal the gradient of ONE process running the global batch 8*world.  Every rank runs its own shard,
    rank 0 also runs the concatenated batch on a copy of the model; returns the max-abs relative error of the flat gradient.
    """
    import copy

    import torch.distributed as dist

    from event_flow_b200.dataloader.encodings import encode_batch

    def window_grad(m, shards):
        m.reset_states()
        lossf.reset()
        m.zero_grad(set_to_none=True)
        for t in range(T):
            ed = torch.cat([make_events(r, 7)[t] for r in shards], dim=0).to(dev)
            d = encode_batch(ed, (H, W), BINS)
            out = m(d["event_voxel"], d["event_cnt"])
            lossf.event_flow_association(out["flow"], ed, d["event_list_pol_mask"], d["event_mask"])
        loss = lossf()
        loss.backward()
        m.reset_states()
        lossf.reset()
        return torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.requires_grad]), loss.detach()

    g_local, l_local = window_grad(model, [rank])
    dist.all_reduce(g_local, op=dist.ReduceOp.SUM)
    dist.all_reduce(l_local, op=dist.ReduceOp.SUM)
    out = None
    if rank == 0:
        twin = copy.deepcopy(model)  # (also exercises FireNet.__getstate__: run-time caches do not travel)
        g_all, l_all = window_grad(twin, list(range(world)))
        out = {"dp_grad_rel_err": ((g_local - g_all).abs().max() / g_all.abs().max()).item(),
               "dp_loss_rel_err": ((l_local - l_all).abs() / l_all.abs()).item(), "global_batch": B_PER_GPU * world,
               "what": "all-reduced SUM of per-rank batch-8 gradients vs one process on the concatenated batch"}
        del twin
        torch.cuda.empty_cache()
    model.zero_grad(set_to_none=True)
    dist.barrier()
    return out


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference on the host cores (fallback: the oracle port)
# ---------------------------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


class ReferenceCPU:
    """The reference's own modules (baseline/_ref, staged unmodified by tools/stage_reference.py), torch-CPU, all host threads."""

    kind = "reference"

    def __init__(self):
        sys.path.insert(0, REF_DIR)
        from dataloader.encodings import events_to_channels, events_to_voxel  # noqa: F401  (reference modules)
        from loss.flow import EventWarping
        from models.model import LIFFireNet

        self.cores = os.cpu_count()
        torch.set_num_threads(self.cores)
        torch.manual_seed(0)
        LIFFireNet.kwargs = [{}] * 7  # the reference shares one class-level kwargs list between all FireNets (model.py:159)
        self.model = LIFFireNet(dict(MODEL_CFG, spiking_neuron=dict(LIF))).train()
        scale_weights(self.model)
        self.lossf = EventWarping(LOSS_CFG, torch.device("cpu"))
        self.opt = torch.optim.Adam(self.model.parameters(), lr=LR)
        self.enc = (events_to_voxel, events_to_channels)
        self.model.reset_states()

    def window(self, k):
        to_voxel, to_channels = self.enc
        out = []
        for e in make_events(0, k):
            ts, ys, xs, ps = e[:, :, 0], e[:, :, 1], e[:, :, 2], e[:, :, 3]
            vox = torch.stack([to_voxel(xs[b], ys[b], ts[b], ps[b], BINS, sensor_size=(H, W)) for b in range(B_PER_GPU)])
            cnt = torch.stack([to_channels(xs[b], ys[b], ps[b], sensor_size=(H, W)) for b in range(B_PER_GPU)])
            pm = torch.stack([(ps > 0).float(), (ps < 0).float()], 2)
            mask = (cnt.sum(1, keepdim=True) > 0).float()
            out.append((vox, cnt, e.clone(), pm, mask))
        return out

    def train_step(self, win):  # train_flow.py:98-171
        self.lossf.reset()
        for vox, cnt, ev, pm, mask in win:
            x = self.model(vox, cnt)
            self.lossf.event_flow_association(x["flow"], ev, pm, mask)
        loss = self.lossf()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), CLIP)
        self.opt.step()
        self.opt.zero_grad()
        self.model.detach_states()
        return loss.item()

    def fwd_loss(self, win):
        self.lossf.reset()
        with torch.no_grad():
            for vox, cnt, ev, pm, mask in win:
                x = self.model(vox, cnt)
                self.lossf.event_flow_association(x["flow"], ev, pm, mask)
            return self.lossf().item()


class PortCPU:
    """Fallback when baseline/_ref is absent: the oracle's restatement of the same step (torch-CPU fp32, autograd)."""

    kind = "port"

    def __init__(self):
        from oracle import encodings as oenc
        from oracle import iwe as oiwe
        from oracle import spiking as osp

        self.oenc, self.oiwe, self.osp = oenc, oiwe, osp
        self.cores = os.cpu_count()
        torch.set_num_threads(self.cores)
        self.params = osp.init_firenet_params("lif", BINS, 32, seed=0, weight_gain=W_GAIN)
        self.params["pred"]["weight"] = self.params["pred"]["weight"] * PRED_GAIN
        self.leaves = [t.requires_grad_(True) for layer in self.params.values() for t in layer.values() if torch.is_tensor(t) and t.is_floating_point()]
        self.opt = torch.optim.Adam(self.leaves, lr=LR)
        self.states = [None] * 7

    def window(self, k):
        return [self.oenc.encode_window(e[:, :, 0], e[:, :, 1], e[:, :, 2], e[:, :, 3], H, W, BINS) for e in make_events(0, k)]

    def _loss(self, win):
        flows, evs, pms, masks = [], [], [], []
        for t, d in enumerate(win):
            flow, self.states, _ = self.osp.firenet_step("lif", self.params, self.states, d["event_voxel"])
            flows.append(flow)
            e = d["event_list"].clone()
            e[:, :, 0] += t
            evs.append(e), pms.append(d["event_list_pol_mask"]), masks.append(d["event_mask"])
        return self.oiwe.event_warping_loss(torch.cat(evs, 1), torch.cat(pms, 1), torch.arange(T).repeat_interleave(N_EV), [torch.stack(flows, 1)],
                                            torch.cat(masks, 1), (H, W), weight=0.001, passes=T)

    def train_step(self, win):
        loss = self._loss(win)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.leaves, CLIP)
        self.opt.step()
        self.opt.zero_grad()
        self.states = [None if s is None else s.detach() for s in self.states]
        return loss.item()

    def fwd_loss(self, win):
        with torch.no_grad():
            return self._loss(win).item()


def cpu_arm():
    if os.path.isfile(os.path.join(REF_DIR, "models", "model.py")):
        return ReferenceCPU()
    return PortCPU()


def cpu_time(fn, windows, steps, warmup):
    for k in range(warmup):
        fn(windows[k])
    t0 = time.perf_counter()
    for k in range(steps):
        fn(windows[warmup + k])
    return (time.perf_counter() - t0) / steps


def cpu_baseline():
    """Bounded sample for the `cpu_baseline` key of our own line: 1 warm-up + 2 timed training windows, 1 + 2 forward+loss windows."""
    arm = cpu_arm()
    wins = [arm.window(k) for k in range(6)]
    dt = cpu_time(arm.train_step, wins[:3], 2, 1)
    dt_f = cpu_time(arm.fwd_loss, wins[3:], 2, 1)
    what = "unmodified reference modules (baseline/_ref)" if arm.kind == "reference" else "oracle port of the reference"
    return {"value": B_PER_GPU * T * N_EV / dt, "unit": "events/s", "cores": arm.cores, "kind": arm.kind, "ms_per_step": dt * 1e3,
            "fwd_loss_value": B_PER_GPU * T * N_EV / dt_f, "fwd_loss_ms_per_step": dt_f * 1e3,
            "sample": f"2 full training windows (batch 8, 10 timesteps: fwd + loss + backward + clip + Adam) after 1 warm-up, {what}, torch-CPU fp32, "
                      f"{arm.cores} threads; fwd_loss_*: 2 forward+loss windows after 1 warm-up"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    arm = cpu_arm()
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    wins = [arm.window(k) for k in range(steps + warmup)]
    dt = cpu_time(arm.train_step, wins, steps, warmup)
    v = B_PER_GPU * T * N_EV / dt
    what = "unmodified reference modules (baseline/_ref)" if arm.kind == "reference" else "oracle port of the reference (baseline/_ref not staged)"
    sample = (f"{steps} full training windows (batch 8, 10 timesteps; one rank's share) after {warmup} warm-up, {what}, torch-CPU fp32, "
              f"{arm.cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "events/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "global_batch": B_PER_GPU, "parallelism": "cpu", "weights": WEIGHTS},
        "cpu_baseline": {"value": v, "unit": "events/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
