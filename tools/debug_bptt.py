"""Debug aid: per-parameter relative BPTT gradient error of the LIFFireNet fast path vs the teacher-forced CPU oracle, for several
backward variants (tensor-core / CUDA-core backward, CUDA graphs on / off, window index)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import encodings as oenc, iwe as oiwe
from tests.util import firenet_cfg, oracle_params_of, oracle_bptt_teacher_forced, model_grads_by_layer, rel_err
import event_flow_b200.models.model as M
from event_flow_b200.loss.flow import EventWarping

DEV = "cuda"
B, H, W, T, N, bins = [int(v) for v in (sys.argv[1:7] if len(sys.argv) > 6 else (8, 128, 128, 10, 1000, 5))]


def run(tag, windows, use_loss=True, **flags):
    torch.manual_seed(0)
    m = M.LIFFireNet(firenet_cfg(bins, "voxel"))
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        m.pred.conv2d.weight.mul_(20.0)
    params = oracle_params_of(m)
    m = m.to(DEV).train()
    for k, v in flags.items():
        m.__dict__[k] = v
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False}, "model": {"mask_output": True}}
    lossf = EventWarping(cfg, DEV)
    g = torch.Generator().manual_seed(1)
    gws = [torch.rand((B, 2, H, W), generator=g) - 0.5 for _ in range(T)]
    for win in range(windows):
        states0 = None if win == 0 else [s.cpu() for s in m.states]
        lossf.reset()
        m.zero_grad(set_to_none=True)
        data, spikes, loss = [], [], 0
        for t in range(T):
            d = oenc.encode_window(*oenc.synthetic_events(B, N, H, W, 7000 + 100 * win + t), H, W, bins)
            data.append(d)
            out = m(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV))
            if use_loss:
                lossf.event_flow_association(out["flow"], d["event_list"].clone().to(DEV), d["event_list_pol_mask"].to(DEV), d["event_mask"].to(DEV))
            else:
                loss = loss + (out["flow"][0] * gws[t].to(DEV)).sum()
            if win == windows - 1:
                spikes.append([s[1].cpu() for s in m.states])
        if use_loss:
            loss = lossf()
        loss.backward()
        if win < windows - 1:
            m.detach_states()

    def oracle_loss(flows):
        if not use_loss:
            return sum((f * gws[t]).sum() for t, f in enumerate(flows))
        evs = []
        for t, d in enumerate(data):
            e = d["event_list"].clone()
            e[:, :, 0] += t
            evs.append(e)
        return oiwe.event_warping_loss(torch.cat(evs, 1), torch.cat([d["event_list_pol_mask"] for d in data], 1), torch.arange(T).repeat_interleave(N),
                                       [torch.stack(flows, 1)], torch.cat([d["event_mask"] for d in data], 1), (H, W), weight=0.001, passes=T)

    loss_o, ref, _ = oracle_bptt_teacher_forced("lif", params, [d["event_voxel"] for d in data], spikes, oracle_loss, states0=states0)
    mine = model_grads_by_layer(m)
    row = []
    for l, lp in ref.items():
        for k, gr in lp.items():
            if gr is not None and gr.abs().max() > 0:
                row.append(f"{l}.{k}={rel_err(mine[l][k].detach().cpu().double().reshape(gr.shape), gr.double()):.1e}")
    print(f"[{tag}] loss {loss.item():.6f} vs {loss_o.item():.6f}\n   " + " ".join(row), flush=True)


run("window backward, random-weight loss, window 1", 1, use_loss=False)
run("window backward, random-weight loss, window 3", 3, use_loss=False)
run("step-by-step tc backward, random-weight loss, window 3", 3, use_loss=False, _window_backward=False)
run("step-by-step cuda-core backward, random-weight loss, window 1", 1, use_loss=False, _tc_backward=False)
run("window backward, event-warping loss, window 3 (flows not forced: see profiles/r02_loss_gradient_noise_floor.txt)", 3)
