"""Measures the accumulation error of the tcgen05 and CUDA-core conv+LIF kernels and the CPU fp32 path against an fp64 oracle,
and times both kernels (CUDA events).  Run on the GPU box:  python tools/tc_accuracy_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_flow_b200 import ops  # noqa: E402
from oracle import spiking as osp  # noqa: E402

DEV = "cuda"
for rec in (False, True):
    for gain in (1.0, 2.0):
        B, H, W = 8, 128, 128
        g = torch.Generator().manual_seed(1)
        params = osp.init_firenet_params("lif", 32, 32, seed=1, weight_gain=gain)["G1" if rec else "R1a"]
        x = (torch.rand((B, 32, H, W), generator=g) < 0.3).float()
        st = torch.rand((2, B, 32, H, W), generator=g) * 1.2 - 0.1
        st[1] = (st[1] < 0.3).float()
        pd = {k: v.to(DEV).contiguous() for k, v in params.items()}
        x_cl, v_in, z_in = ops.pack_cl(x.to(DEV)), st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV))
        ws = ops.split_weights(pd["ff"], pd.get("rec"))
        leak, thresh = pd["leak"].reshape(-1), pd["thresh"].reshape(-1)
        args = (x_cl, v_in, z_in, pd["ff"], pd.get("rec"), leak, thresh)
        v_tc, z_tc = ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
        v_cc, z_cc = ops.lif_step_cl(*args, hard_reset=True, w_split=None)
        _, ns32 = osp.cell_step("lif", x, st, params, hard_reset=True)
        p64 = {k: v.double() for k, v in params.items()}
        _, ns64 = osp.cell_step("lif", x.double(), st.double(), p64, hard_reset=True)
        e = lambda a: (a.double().cpu() - ns64[0]).abs().max().item()
        thr = params["thresh"].clamp_min(0.01).double()
        near = (ns64[0] - thr).abs() < 1e-5
        fl = lambda z: ((ops.unpack_cl(z).cpu().double() != ns64[1]) & ~near).sum().item()
        times = {}
        for name, wsx in (("tc", ws), ("cc", None)):
            for _ in range(3):
                ops.lif_step_cl(*args, hard_reset=True, w_split=wsx)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                ops.lif_step_cl(*args, hard_reset=True, w_split=wsx)
            t1.record()
            torch.cuda.synchronize()
            times[name] = t0.elapsed_time(t1) / 20 * 1e3
        print(f"rec={rec} gain={gain}: max|v-v64| tc={e(v_tc):.2e} cuda-core={e(v_cc):.2e} cpu-fp32={e(ns32[0]):.2e} max|v|={ns64[0].abs().max():.1f} "
              f"flips-outside-band tc={fl(z_tc)} cc={fl(z_cc)} of {z_in.numel()} | us/launch tc={times['tc']:.1f} cc={times['cc']:.1f} (incl. host launch path)")
