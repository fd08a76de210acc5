"""
Per-kernel timing of the forward path at bench size (B=8, 128x128, 32 ch), for kernel work between bench runs.

Each kernel is launched back to back over a ring of buffer sets larger than the 126 MB L2 (so every launch streams its
operands from HBM, like inside a model step), bracketed by ONE pair of CUDA events: the figure is the average launch
duration including the inter-launch gap, i.e. what the kernel costs inside a step.  Prints algorithmic GB/s (SURVEY 8d:
4*HW*(Cin + 2*S_r*C) bytes per sample) against MEASURED_PEAKS.json.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from event_flow_b200 import _lib as L  # noqa: E402
from event_flow_b200 import ops  # noqa: E402
from oracle import spiking as osp  # noqa: E402

DEV = "cuda"
B, H, W = 8, 128, 128
NSET = 6  # 6 x ~59 MB of operands per launch > L2
REPS = 10
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
g = torch.Generator().manual_seed(1)


def rnd_sets(cin_f32=None):
    sets = []
    for _ in range(NSET):
        x = (torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV)
        v = (torch.rand((B, 32, H, W), generator=g) * 1.2 - 0.1).to(DEV)
        z = (torch.rand((B, 32, H, W), generator=g) < 0.3).float().to(DEV)
        d = dict(x_cl=ops.pack_cl(x), v=v, z_cl=ops.pack_cl(z), v_out=torch.empty_like(v), z_out=torch.empty((B, H, W, 32), device=DEV, dtype=torch.bfloat16))
        if cin_f32:
            d["x5"] = torch.randn((B, cin_f32, H, W), generator=g).to(DEV)
        sets.append(d)
    return sets


def timed(fn, label, alg_bytes):
    """REPS x NSET launches captured in ONE CUDA graph (no host launch cost in the figure, like a model step), best of 3 replays."""
    for i in range(NSET):
        fn(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for r in range(REPS):
            for i in range(NSET):
                fn(i)
    graph.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (REPS * NSET))
    gbs = alg_bytes / best / 1e3
    print(f"{label:34s} {best:8.2f} us/launch   {gbs:8.0f} GB/s algorithmic = {100 * gbs / PEAK:5.1f}% of measured {PEAK:.0f} GB/s", flush=True)
    return best


def fill(p, cell, cin, x_f32, x_cl, d, ws):
    p.B, p.Cin, p.C, p.H, p.W = B, cin, 32, H, W
    p.ksize, p.stride, p.neuron, p.hard_reset = 3, 1, L.EF_LIF, 1
    p.surrogate, p.act_width = L.SURROGATE_CODES["arctanspike"], 10.0
    p.x, p.x_cl = L.ptr(x_f32), L.ptr(x_cl)
    p.v_in, p.z_in_cl = L.ptr(d["v"]), L.ptr(d["z_cl"])
    p.w_ff, p.w_rec = L.ptr(cell["ff"]), L.ptr(cell.get("rec"))
    p.leak, p.thresh = L.ptr(cell["leak"]), L.ptr(cell["thresh"])
    p.v_out, p.z_out_cl = L.ptr(d["v_out"]), L.ptr(d["z_out"])
    p.w_split = L.ptr(ws)


def ablate(sets):
    """Per-stream cost inside the tensor-core kernel (debug build; results are wrong under these switches)."""
    alg = 4 * H * W * (32 + 2 * 2 * 32) * B
    for rec in (False, True):
        cell = {k: t.to(DEV).contiguous().reshape(-1) if k in ("leak", "thresh") else t.to(DEV).contiguous()
                for k, t in osp.init_firenet_params("lif", 32, 32, seed=1, weight_gain=2.0)["G1" if rec else "R1a"].items()}
        ws = ops.split_weights(cell["ff"], cell.get("rec"))
        ps = []
        for d in sets:
            p = L.LifConvParams()
            fill(p, cell, 32, None, d["x_cl"], d, ws)
            ps.append(p)
        for mask, name in ((0, "production build"), (16384, "debug build, nothing off"), (1, "no v_out TMA store"), (2, "no v_in TMA load"), (3, "no v traffic"),
                           (8, "no z_out TMA store"), (1024, "no centre-z TMA load (ff)"), (1 + 2 + 8 + 1024, "only operand loads"), (4, "no MMAs"),
                           (4 + 1 + 2 + 8 + 1024, "no MMAs, only operand loads"), (32, "prologue + teardown only"), (128, "one tile per CTA")):
            L.lib().ef_debug_tc_skip(mask)
            timed(lambda i: L.call("ef_lif_conv_fwd", ps[i]), f"rec={int(rec)} {name}", alg)
        L.lib().ef_debug_tc_skip(0)


def main():
    if len(sys.argv) > 1:
        L.lib().ef_debug_tc_cpt(int(sys.argv[1]))
        print("tensor-core kernel: channels per epilogue thread =", sys.argv[1])
    sets = rnd_sets(cin_f32=5)
    res = {}
    if len(sys.argv) > 2 and sys.argv[2] == "ablate":
        return ablate(sets)
    alg = 4 * H * W * (32 + 2 * 2 * 32) * B
    for rec in (False, True):
        cell = {k: t.to(DEV).contiguous().reshape(-1) if k in ("leak", "thresh") else t.to(DEV).contiguous()
                for k, t in osp.init_firenet_params("lif", 32, 32, seed=1, weight_gain=2.0)["G1" if rec else "R1a"].items()}
        ws = ops.split_weights(cell["ff"], cell.get("rec"))
        ps = []
        for d in sets:
            p = L.LifConvParams()
            fill(p, cell, 32, None, d["x_cl"], d, ws)
            ps.append(p)
        res["tc_rec" if rec else "tc_ff"] = timed(lambda i: L.call("ef_lif_conv_fwd", ps[i]), f"conv+LIF 32->32 {'rec' if rec else 'ff '} (tcgen05)", alg)
    cell = {k: t.to(DEV).contiguous().reshape(-1) if k in ("leak", "thresh") else t.to(DEV).contiguous()
            for k, t in osp.init_firenet_params("lif", 5, 32, seed=1, weight_gain=2.0)["head"].items()}
    ps = []
    for d in sets:
        p = L.LifConvParams()
        fill(p, cell, 5, d["x5"], None, d, None)
        ps.append(p)
    res["head"] = timed(lambda i: L.call("ef_lif_conv_fwd", ps[i]), "conv+LIF head 5->32 (fp32 cores)", 4 * H * W * (5 + 2 * 2 * 32) * B)
    w = torch.randn((2, 32), generator=g).to(DEV) * 0.1
    b = torch.zeros(2, device=DEV)
    flow = torch.empty((B, 2, H, W), device=DEV)
    ps = []
    for d in sets:
        pp = L.PredParams()
        pp.B, pp.Cin, pp.Cout, pp.H, pp.W = B, 32, 2, H, W
        pp.x_cl, pp.w, pp.b, pp.y = L.ptr(d["x_cl"]), L.ptr(w), L.ptr(b), L.ptr(flow)
        ps.append(pp)
    res["pred"] = timed(lambda i: L.call("ef_pred_fwd", ps[i]), "pred 1x1+tanh 32->2", 4 * H * W * (32 + 2) * B)
    step = res["head"] + 2 * res["tc_rec"] + 4 * res["tc_ff"] + res["pred"]
    print(f"model step (sum of 8 kernels): {step:.1f} us -> window of 10: {step / 100:.3f} ms")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "kbench.json"), "w"))


if __name__ == "__main__":
    main()
