"""Debug: windowed vs stepwise BPTT -- compare the saved activations of the arena and the gradients, and the run-to-run noise floor."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_flow_b200 import fast  # noqa: E402
from event_flow_b200.loss.flow import EventWarping  # noqa: E402
from tests.test_gpu_window import DEV, _model, _windows  # noqa: E402

B, H, W, T = 2, 32, 48, 4
cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False}, "model": {"mask_output": True}}
wins = _windows(B, 600, H, W, 2 * T, 5, 110)


def run(windowed):
    model = _model()
    lossf = EventWarping(cfg, DEV)
    res = []
    for w in range(2):
        part = wins[w * T:(w + 1) * T]
        model.zero_grad(set_to_none=True)
        if windowed:
            outs = model.forward_window(torch.stack([d["event_voxel"] for d in part]).to(DEV), torch.stack([d["event_cnt"] for d in part]).to(DEV))
        else:
            outs = [model(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV)) for d in part]
        for d, out in zip(part, outs):
            lossf.event_flow_association(out["flow"], d["event_list"].clone().to(DEV), d["event_list_pol_mask"].to(DEV), d["event_mask"].to(DEV))
        loss = lossf()
        bank = model._arena.banks[model._arena.parity]
        acts = {"v": [v[:T].clone() for v in bank.v], "z": [z[:T + 1].clone() for z in bank.zs], "flow": bank.flow[:T].clone(), "x": bank.x_cl[:T].clone()}
        loss.backward()
        lossf.reset()
        model.detach_states()
        res.append((loss.item(), {n: p.grad.clone() for n, p in model.named_parameters()}, acts))
    return res


def cmp(ra, rb, what):
    print("==", what)
    for w in range(2):
        la, ga, aa = ra[w]
        lb, gb, ab = rb[w]
        print(f" window {w}: loss {la!r} {lb!r}")
        for i in range(7):
            dv = (aa["v"][i] - ab["v"][i]).abs().max().item()
            dz = (aa["z"][i][1:].float() - ab["z"][i][1:].float()).abs().max().item()
            if dv or dz:
                print(f"   layer {i}: dv {dv:.3e} dz {dz:.3e}")
        print("   flow", (aa["flow"] - ab["flow"]).abs().max().item(), "x", (aa["x"].float() - ab["x"].float()).abs().max().item())
        for n in gb:
            d = (ga[n] - gb[n]).abs().max().item()
            if d:
                print(f"   grad {n}: {d:.3e} rel {d / (gb[n].abs().max().item() + 1e-30):.3e}")


s1, s2, w1, w2 = run(False), run(False), run(True), run(True)
cmp(s1, s2, "stepwise vs stepwise")
cmp(w1, w2, "windowed vs windowed")
cmp(w1, s1, "windowed vs stepwise")
