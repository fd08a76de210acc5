"""Kernel-time breakdown of one PLIF / ALIF FireNet (cfg 5) or spiking EV-FlowNet (cfg 4) training window with torch.profiler: device time per kernel vs wall time."""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.bench_configs import FIRE, _events  # noqa: E402


def main(cls="ALIFFireNet", B=8, N=1000, H=128, W=128, T=20, bins=5, gain=2.5):
    cfg = FIRE
    if cls.endswith("EVFlowNet"):  # cfg 4: the U-Net family at 256x256, 50k events per window, batch 4
        from tools.bench_configs import UNET

        cfg, B, N, H, W, T, bins, gain = UNET, 4, 50000, 256, 256, 4, 2, 3.0
    import event_flow_b200.models.model as M
    from event_flow_b200.dataloader.encodings import encode_batch
    from event_flow_b200.loss.flow import EventWarping

    dev = torch.device("cuda")
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

    def train():
        model.zero_grad(set_to_none=True)
        model.reset_states()
        lossf.reset()
        for vox, cnt, ev, pm, mask in win:
            out = model(vox, cnt)
            lossf.event_flow_association(out["flow"], ev.clone(), pm, mask)
        loss = lossf()
        t1 = time.perf_counter()
        loss.backward()
        model.detach_states()
        return t1

    for _ in range(2):
        train()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t1 = train()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"{cls}: host forward {1e3 * (t1 - t0):.1f} ms, host backward {1e3 * (t2 - t1):.1f} ms, wall {1e3 * (t3 - t0):.1f} ms")
    if os.environ.get("EF_CPROFILE"):
        import cProfile
        import pstats

        pr = cProfile.Profile()
        pr.enable()
        train()
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats("tottime").print_stats(38)
        return
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        train()
        torch.cuda.synchronize()
    rows = [(e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
    rows.sort(key=lambda r: -r[2])
    print(f"device time total {sum(r[2] for r in rows):.2f} ms")
    for k, c, t in rows[:16]:
        print(f"  {t:8.3f} ms  n={c:5d}  {k[:100]}")


if __name__ == "__main__":
    main(*sys.argv[1:2])
