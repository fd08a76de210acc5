"""
Pins the oracle against the reference ITSELF and writes the committed golden fixtures.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the only place /root/reference exists):
    python oracle/pin_against_reference.py            # check + (re)write tests/golden/*.npz and MANIFEST.json
    python oracle/pin_against_reference.py --check    # check only

For every case the UNMODIFIED reference modules (imported from /root/reference) and the oracle restatement are run on
the same seeded inputs; outputs AND autograd gradients must agree (bit-equal for forward values on this torch build,
<=1e-6 relative for gradients).  The reference's outputs are what is stored in the fixtures, so tests on the GPU box
(where /root/reference does not exist) still compare against the reference's own numbers.
"""

import argparse
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import encodings as oenc  # noqa: E402
from oracle import iwe as oiwe  # noqa: E402
from oracle import spiking as osp  # noqa: E402
from oracle import unet as ounet  # noqa: E402
from oracle import annzoo as ozoo  # noqa: E402


def import_reference():
    assert os.path.isdir(os.path.join(REF, "models")), "reference not mounted; run this in the build container"
    sys.path.insert(0, REF)
    import dataloader.encodings as renc
    import loss.flow as rflow
    import models.model as rmodel
    import models.spiking_submodules as rcells
    import utils.iwe as riwe

    return rcells, rmodel, rflow, riwe, renc


def close(a, b, rtol, what):
    a, b = a.detach(), b.detach()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    ok = err <= rtol * max(ref, 1e-30) if rtol > 0 else err == 0
    assert ok, f"{what}: max|d|={err:.3e} vs max|ref|={ref:.3e} (rtol {rtol})"
    return err


CELL_PARAM_NAMES = {
    "lif": ("leak", "thresh"),
    "plif": ("leak_v", "leak_pt", "add_pt", "thresh"),
    "alif": ("leak_v", "leak_t", "t0", "t1"),
    "xlif": ("leak_v", "leak_pt", "t0", "t1"),
}


def params_of(cell, neuron):
    p = {"ff": cell.ff.weight}
    if hasattr(cell, "rec"):
        p["rec"] = cell.rec.weight
    for n in CELL_PARAM_NAMES[neuron]:
        p[n] = getattr(cell, n)
    return p


def spiky_input(shape, g, density=0.3, signed=False):
    x = (torch.rand(shape, generator=g) < density).float()
    if signed:
        x = x * (torch.randint(0, 2, shape, generator=g) * 2 - 1).float() * torch.randint(1, 3, shape, generator=g).float()
    return x


GOLDEN_CELLS = (
    "cell_lif_ff_hard_cin32", "cell_lif_ff_soft_cin32", "cell_lif_rec_hard_cin32", "cell_lif_rec_soft_cin32",
    "cell_lif_ff_hard_cin5", "cell_plif_rec_hard_cin32", "cell_plif_ff_hard_cin5", "cell_alif_rec_soft_cin32",
    "cell_xlif_rec_soft_cin32", "cell_alif_ff_hard_cin32",
)


def pin_cells(rcells, golden, shape=(2, 12, 16)):
    """All 8 cells x {hard,soft} x {state None, state given}, forward + autograd (teacher-forced T1/T2)."""
    classes = {
        ("lif", False): rcells.ConvLIF,
        ("lif", True): rcells.ConvLIFRecurrent,
        ("plif", False): rcells.ConvPLIF,
        ("plif", True): rcells.ConvPLIFRecurrent,
        ("alif", False): rcells.ConvALIF,
        ("alif", True): rcells.ConvALIFRecurrent,
        ("xlif", False): rcells.ConvXLIF,
        ("xlif", True): rcells.ConvXLIFRecurrent,
    }
    (B, H, W), C = shape, 32
    n = 0
    for (neuron, rec), cls in classes.items():
        for hard in (True, False):
            for cin in (32, 5):
                for surrogate in ("arctanspike",) if cin == 5 else osp.SURROGATES:
                    torch.manual_seed(11 + n)
                    kw = dict(hard_reset=hard, activation=surrogate, learn_thresh=True)
                    if neuron in ("lif", "plif"):
                        kw["thresh"] = (0.8, 0.1)
                    else:
                        kw["t0"], kw["t1"] = (0.05, 0.01), (1.8, 0.1)
                    cell = cls(cin, C, 3, **kw)
                    with torch.no_grad():
                        cell.ff.weight.mul_(2.0)
                    g = torch.Generator().manual_seed(100 + n)
                    x = spiky_input((B, cin, H, W), g, signed=(cin == 5)).requires_grad_(True)
                    n_state = 2 if neuron == "lif" else 3
                    st = torch.rand((n_state, B, C, H, W), generator=g)
                    st[1] = (st[1] < 0.3).float()
                    st = st.requires_grad_(True)
                    width = 0.5 if surrogate == "mgspike" else (1.0 if surrogate == "trianglespike" else 10.0)
                    with torch.no_grad():
                        cell.act_width.fill_(width)
                    for state in (st, None):
                        out_r, s_r = cell(x, state)
                        gout = torch.rand(out_r.shape, generator=g)
                        gst = torch.rand(s_r.shape, generator=g)
                        gst[1] = 0  # z_out reaches the loss through `out` and through next-step state; keep separate
                        inputs_r = [x] + ([state] if state is not None else []) + [
                            p for p in cell.parameters() if p.requires_grad
                        ]
                        grads_r = torch.autograd.grad((out_r * gout).sum() + (s_r * gst).sum(), inputs_r, allow_unused=True)

                        p = params_of(cell, neuron)
                        out_o, s_o = osp.cell_step(neuron, x, state, p, hard_reset=hard, surrogate=surrogate, width=width)
                        grads_o = torch.autograd.grad((out_o * gout).sum() + (s_o * gst).sum(), inputs_r, allow_unused=True)
                        tag = f"cell {neuron} rec={rec} hard={hard} cin={cin} {surrogate} state={'given' if state is not None else 'None'}"
                        close(out_o, out_r, 0, tag + " out")
                        close(s_o, s_r, 0, tag + " state")
                        for go, gr, t in zip(grads_o, grads_r, inputs_r):
                            if gr is None:
                                assert go is None, tag
                            else:
                                close(go, gr, 1e-6, tag + f" grad{tuple(t.shape)}")
                        if state is not None and surrogate == "arctanspike":
                            key = f"cell_{neuron}_{'rec' if rec else 'ff'}_{'hard' if hard else 'soft'}_cin{cin}"
                            if golden is None or key not in GOLDEN_CELLS:
                                continue
                            names = ["x", "state"] + [nm for nm, q in cell.named_parameters() if q.requires_grad]
                            d = {"x": x, "state": state, "out": out_r, "new_state": s_r, "g_out": gout, "g_state": gst}
                            d.update({"p_" + k: v for k, v in p.items()})
                            d.update({"grad_" + nm.replace(".weight", ""): gr for nm, gr in zip(names, grads_r) if gr is not None})
                            d["width"] = torch.tensor(width)
                            golden[key] = d
                    n += 1
    print(f"cells: {n} configurations x 2 state modes pinned")


def pin_firenet(rmodel, golden, shape=(2, 16, 24, 4)):
    """LIF / PLIF / ALIF / XLIF FireNet, T-step rollout + BPTT gradient of a random linear functional of the flows."""
    classes = {"lif": rmodel.LIFFireNet, "plif": rmodel.PLIFFireNet, "alif": rmodel.ALIFFireNet, "xlif": rmodel.XLIFFireNet}
    B, H, W, T = shape
    for neuron, cls in classes.items():
        for bins, enc in ((5, "voxel"), (2, "cnt")):
            cls.kwargs = [{}] * 7  # the reference shares one class-level list of dicts (model.py:159)
            torch.manual_seed(5)
            sn = dict(leak=[-4.0, 0.1], thresh=[0.8, 0.1], learn_leak=True, learn_thresh=True, hard_reset=True) if neuron == "lif" else {}
            cfg = dict(name="x", encoding=enc, round_encoding=False, norm_input=False, num_bins=bins, base_num_channels=32,
                       kernel_size=3, activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron=sn)
            m = cls(cfg)
            with torch.no_grad():
                for nm, q in m.named_parameters():
                    if nm.endswith("ff.weight") or nm.endswith("rec.weight"):
                        q.mul_(2.5)
                m.pred.conv2d.weight.mul_(20.0)
            params = {nm: params_of(getattr(m, nm), neuron) for nm in osp.FIRENET_LAYERS}
            params["pred"] = {"weight": m.pred.conv2d.weight, "bias": m.pred.conv2d.bias}
            g = torch.Generator().manual_seed(77)
            xs = []
            for t in range(T):
                ts, ys, xx, ps = oenc.synthetic_events(B, 400, H, W, 1000 + t)
                d = oenc.encode_window(ts, ys, xx, ps, H, W, bins)
                xs.append((d["event_voxel"], d["event_cnt"]))
            gw = [torch.rand((B, 2, H, W), generator=g) - 0.5 for _ in range(T)]
            # reference
            m.reset_states()
            flows_r, acts_r = [], []
            for t in range(T):
                o = m(xs[t][0].clone(), xs[t][1].clone(), log=True)
                flows_r.append(o["flow"][0])
                acts_r.append(list(o["activity"].values()))
            plist = [q for q in m.parameters() if q.requires_grad]
            loss_r = sum((f * w).sum() for f, w in zip(flows_r, gw))
            grads_r = torch.autograd.grad(loss_r, plist, allow_unused=True)
            states_r = m.states
            # oracle
            states = [None] * 7
            flows_o = []
            for t in range(T):
                x = xs[t][0] if enc == "voxel" else xs[t][1]
                f, states, _ = osp.firenet_step(neuron, params, states, x)
                flows_o.append(f)
            loss_o = sum((f * w).sum() for f, w in zip(flows_o, gw))
            grads_o = torch.autograd.grad(loss_o, plist, allow_unused=True)
            tag = f"firenet {neuron} {enc}"
            for t in range(T):
                close(flows_o[t], flows_r[t], 0, tag + f" flow[{t}]")
            for i in range(7):
                close(states[i], states_r[i], 0, tag + f" state[{i}]")
            for go, gr, (nm, _) in zip(grads_o, grads_r, [(n_, q) for n_, q in m.named_parameters() if q.requires_grad]):
                if gr is not None:
                    close(go, gr, 1e-5, tag + " grad " + nm)
            print(tag, "activity last step:", ["%.3f" % a for a in acts_r[-1]])
            if golden is not None and (enc == "voxel" or neuron == "lif"):
                d = {"x_%d" % t: (xs[t][0] if enc == "voxel" else xs[t][1]) for t in range(T)}
                d.update({"gw_%d" % t: gw[t] for t in range(T)})
                d.update({"flow_%d" % t: flows_r[t] for t in range(T)})
                if neuron in ("lif", "alif") and enc == "voxel":
                    d.update({"state_%d" % i: states_r[i] for i in range(7)})
                else:
                    d.update({"state_%d" % i: states_r[i] for i in (1, 6)})
                d.update({"activity": torch.tensor(acts_r)})
                for nm, q in m.state_dict().items():
                    d["sd_" + nm] = q
                for (nm, q), gr in zip([(n_, q) for n_, q in m.named_parameters() if q.requires_grad], grads_r):
                    if gr is not None:
                        d["grad_" + nm] = gr
                golden[f"firenet_{neuron}_{enc}"] = d


def make_window(B, H, W, T, N, seed, flow_mode):
    g = torch.Generator().manual_seed(seed)
    ev, pm, masks = [], [], []
    for t in range(T):
        ts, ys, xs, ps = oenc.synthetic_events(B, N, H, W, seed * 100 + t)
        d = oenc.encode_window(ts, ys, xs, ps, H, W, 2)
        ev.append(d["event_list"])
        pm.append(d["event_list_pol_mask"])
        masks.append(d["event_mask"])
    if flow_mode == "zero":
        flows = [torch.zeros(B, 2, H, W) for _ in range(T)]
    elif flow_mode == "halfint":  # lands many events on exact integer / half-integer coordinates and out of bounds
        flows = [torch.randint(-4, 5, (B, 2, H, W), generator=g).float() / (2.0 * max(H, W)) for _ in range(T)]
    else:
        flows = [(torch.rand((B, 2, H, W), generator=g) - 0.5) * 0.2 for _ in range(T)]
    return ev, pm, masks, flows


def pin_loss(rflow, golden):
    B, H, W, T, N = 2, 16, 20, 3, 150
    for flow_mode in ("random", "zero", "halfint"):
        for scaling, smask, overwrite, weight in ((True, True, False, 0.001), (False, False, False, 0.5), (True, True, True, 0.1)):
            ev, pm, masks, flows = make_window(B, H, W, T, N, 3, flow_mode)
            flows = [f.requires_grad_(True) for f in flows]
            cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": weight, "overwrite_intermediate": overwrite},
                   "model": {"mask_output": smask}}
            L = rflow.EventWarping(cfg, torch.device("cpu"), loss_scaling=scaling)
            for t in range(T):
                L.event_flow_association([flows[t]], ev[t].clone(), pm[t], masks[t])
            if overwrite:
                L.overwrite_intermediate_flow([flows[-1]])
            loss_r = L()
            grads_r = torch.autograd.grad(loss_r, flows, allow_unused=True)
            # oracle, map form
            events = torch.cat([e.clone() for e in ev], dim=1)
            for t in range(T):
                events[:, t * N:(t + 1) * N, 0] += t
            pol = torch.cat(pm, dim=1)
            pass_of_event = torch.arange(T).repeat_interleave(N)
            if overwrite:
                fm = [flows[-1].unsqueeze(1)]
                em = torch.cat(masks, dim=1).sum(1, keepdim=True).clamp(max=1)
            else:
                fm = [torch.stack(flows, dim=1)]
                em = torch.cat(masks, dim=1)
            loss_o = oiwe.event_warping_loss(events, pol, pass_of_event, fm, em, (H, W), weight=weight, loss_scaling=scaling,
                                             smoothing_mask=smask, overwrite_intermediate=overwrite, passes=T)
            grads_o = torch.autograd.grad(loss_o, flows, allow_unused=True)
            tag = f"loss {flow_mode} scaling={scaling} mask={smask} overwrite={overwrite}"
            close(loss_o, loss_r, 1e-6, tag)
            for t in range(T):
                if grads_r[t] is None:
                    assert grads_o[t] is None
                else:
                    close(grads_o[t], grads_r[t], 1e-5, tag + f" grad[{t}]")
            d = {"events": events, "pol_mask": pol, "pass_of_event": pass_of_event.float(), "event_mask": torch.cat(masks, dim=1),
                 "flow": torch.stack([f.detach() for f in flows], dim=1), "loss": loss_r.detach(),
                 "cfg": torch.tensor([float(scaling), float(smask), float(overwrite), weight, T, N])}
            for t in range(T):
                if grads_r[t] is not None:
                    d["grad_%d" % t] = grads_r[t]
            golden[f"loss_{flow_mode}_{int(scaling)}{int(smask)}{int(overwrite)}"] = d
            print(tag, "loss=%.6f" % loss_r.item())


def pin_iwe_and_encodings(riwe, renc, golden):
    B, H, W, N = 2, 16, 20, 300
    ts, ys, xs, ps = oenc.synthetic_events(B, N, H, W, 9)
    g = torch.Generator().manual_seed(9)
    flow = (torch.rand((B, 2, H, W), generator=g) - 0.5) * 0.3
    ev = torch.stack([ts, ys, xs, ps], dim=2)
    pm = torch.stack([oenc.polarity_mask(ps[b]) for b in range(B)])
    d = {"events": ev, "pol_mask": pm, "flow": flow}
    for rnd in (True, False):
        r = riwe.compute_pol_iwe(flow, ev, (H, W), pm[:, :, 0:1], pm[:, :, 1:2], flow_scaling=max(H, W), round_idx=rnd)
        o = oiwe.pol_iwe(flow, ev, (H, W), pm[:, :, 0:1], pm[:, :, 1:2], flow_scaling=max(H, W), round_idx=rnd)
        close(o, r, 0, f"compute_pol_iwe round={rnd}")
        d["iwe_round" if rnd else "iwe_bilinear"] = r
    ev_flow = oiwe.gather_event_flow(flow, ev, (H, W))
    for tref in (0, 1):
        ir, wr = riwe.get_interpolation(ev, ev_flow, tref, (H, W), max(H, W))
        io, wo = oiwe.warp_and_split(ev, ev_flow, tref, (H, W), max(H, W))
        close(io, ir, 0, "get_interpolation idx")
        close(wo, wr, 0, "get_interpolation weights")
        d[f"idx_tref{tref}"], d[f"w_tref{tref}"] = ir, wr
    golden["iwe_image"] = d
    # encodings
    e = {}
    for bins in (2, 5):
        for b in range(B):
            vr = renc.events_to_voxel(xs[b], ys[b], ts[b], ps[b], bins, sensor_size=(H, W))
            close(oenc.events_to_voxel(xs[b], ys[b], ts[b], ps[b], bins, (H, W)), vr, 0, "voxel")
            e[f"voxel{bins}_{b}"] = vr
    for b in range(B):
        cr = renc.events_to_channels(xs[b], ys[b], ps[b], sensor_size=(H, W))
        close(oenc.events_to_channels(xs[b], ys[b], ps[b], (H, W)), cr, 0, "cnt")
        mr = renc.events_to_image(xs[b], ys[b], ps[b].abs(), sensor_size=(H, W), accumulate=False)
        close(oenc.event_mask(xs[b], ys[b], ps[b], (H, W))[0], mr, 0, "mask")
        e[f"cnt_{b}"], e[f"mask_{b}"] = cr, mr
    e.update({"ts": ts, "ys": ys, "xs": xs, "ps": ps})
    golden["encodings"] = e
    print("iwe image + encodings pinned")


def ann_params_of(m, recurrent):
    params = {}
    for name in osp.FIRENET_LAYERS:
        cell = getattr(m, name)
        if recurrent and name in osp.FIRENET_RECURRENT:
            params[name] = {"update_w": cell.update_gate.weight, "update_b": cell.update_gate.bias, "reset_w": cell.reset_gate.weight,
                            "reset_b": cell.reset_gate.bias, "out_w": cell.out_gate.weight, "out_b": cell.out_gate.bias}
        else:
            params[name] = {"w": cell.conv2d.weight, "b": cell.conv2d.bias}
    params["pred"] = {"weight": m.pred.conv2d.weight, "bias": m.pred.conv2d.bias}
    return params


def pin_ann_firenet(rmodel, golden):
    """ANN FireNet (ConvGRU) and FireFlowNet, BASELINE cfg 1 plumbing: 1 voxel bin, batch 1; plus a 2-bin cnt variant."""
    for cls, recurrent, bins, enc in ((rmodel.FireNet, True, 1, "voxel"), (rmodel.FireFlowNet, False, 2, "cnt")):
        cls.kwargs = [{}] * 7
        torch.manual_seed(7)
        cfg = dict(name="x", encoding=enc, round_encoding=False, norm_input=False, num_bins=bins, base_num_channels=32, kernel_size=3,
                   activations=["relu", None], mask_output=True, spiking_neuron=None)
        m = cls(cfg).eval()
        params = ann_params_of(m, recurrent)
        B, H, W, T = 1, 32, 40, 3
        m.reset_states()
        states = [None] * 7
        d_all = {}
        with torch.no_grad():
            for t in range(T):
                ts, ys, xs, ps = oenc.synthetic_events(B, 1000, H, W, 4000 + t)
                d = oenc.encode_window(ts, ys, xs, ps, H, W, bins)
                x = d["event_voxel"] if enc == "voxel" else d["event_cnt"]
                out = m(d["event_voxel"].clone(), d["event_cnt"].clone())
                flow_o, states, _ = osp.firenet_ann_step(params, states, x, recurrent=recurrent)
                close(flow_o, out["flow"][0], 0, f"ann {cls.__name__} flow[{t}]")
                d_all[f"x_{t}"], d_all[f"flow_{t}"] = x, out["flow"][0]
            for i, s_ in enumerate(m.states):
                if recurrent and i in (1, 4):
                    close(states[i], s_, 0, f"ann {cls.__name__} state[{i}]")
                    d_all[f"state_{i}"] = s_
        # BPTT gradients of a random linear functional of the flows (reference autograd; the ANN cells are smooth, so the
        # GPU path is compared at 1e-3 relative without any trajectory caveat)
        g = torch.Generator().manual_seed(31)
        gw = [torch.rand((B, 2, H, W), generator=g) - 0.5 for _ in range(T)]
        named = [(n_, q) for n_, q in m.named_parameters() if q.requires_grad]
        m.reset_states()
        loss = 0.0
        for t in range(T):
            xt = d_all[f"x_{t}"]
            loss = loss + (m(xt.clone(), xt.clone())["flow"][0] * gw[t]).sum()
        grads = torch.autograd.grad(loss, [q for _, q in named], allow_unused=True)
        for t in range(T):
            d_all[f"gw_{t}"] = gw[t]
        for (nm, _), gr in zip(named, grads):
            if gr is not None:
                d_all["grad_" + nm] = gr
        for nm, q in m.state_dict().items():
            d_all["sd_" + nm] = q
        golden[f"ann_{cls.__name__.lower()}"] = d_all
        print(f"ann {cls.__name__} pinned; |flow| max {out['flow'][0].abs().max().item():.4f}")


def pin_metrics(rflow, golden):
    """FWL / RSAT / AEE (loss/flow.py:468-628) on a 3-pass validation window, with and without overwrite_intermediate."""
    B, H, W, T, N = 2, 16, 20, 3, 200
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"overwrite_intermediate": False}}
    d_all = {}
    for overwrite in (False, True):
        g = torch.Generator().manual_seed(21)  # identical inputs for both variants
        cfg["loss"]["overwrite_intermediate"] = overwrite
        metrics = {n: getattr(rflow, n)(cfg, torch.device("cpu"), flow_scaling=max(H, W)) for n in ("FWL", "RSAT", "AEE")}
        flows, inputs = [], []
        for t in range(T):
            ts, ys, xs, ps = oenc.synthetic_events(B, N, H, W, 700 + t)
            d = oenc.encode_window(ts, ys, xs, ps, H, W, 2)
            gt = (torch.rand((B, 2, H, W), generator=g) - 0.5) * 12
            gt[:, :, :3, :] = 0  # some pixels without valid ground truth
            inp = {"event_list": d["event_list"], "event_list_pol_mask": d["event_list_pol_mask"], "event_mask": d["event_mask"],
                   "gtflow": gt, "dt_input": torch.tensor(0.01), "dt_gt": torch.tensor(0.03)}  # eval runs with batch size 1: scalar dt
            f = (torch.rand((B, 2, H, W), generator=g) - 0.5) * 0.3
            flows.append(f), inputs.append(inp)
            for m in metrics.values():
                m.event_flow_association([f], inp)
        if overwrite:
            for m in metrics.values():
                m.overwrite_intermediate_flow([flows[-1]])
        fwl_r, rsat_r = metrics["FWL"](), metrics["RSAT"]()
        aee_r, pct_r = metrics["AEE"]()
        # oracle
        events = torch.cat([inp["event_list"].clone() for inp in inputs], dim=1)
        for t in range(T):
            events[:, t * N:(t + 1) * N, 0] += t
        pol = torch.cat([inp["event_list_pol_mask"] for inp in inputs], dim=1)
        ev_flow = torch.cat([oiwe.gather_event_flow(flows[-1] if overwrite else flows[t], inputs[t]["event_list"], (H, W)) for t in range(T)], dim=1)
        fwl_o, rsat_o = oiwe.fwl_rsat(events, pol, ev_flow, T, (H, W), max(H, W))
        last_mask = inputs[-1]["event_mask"][:, 0]
        if overwrite:  # overwrite_intermediate_flow collapses the masks of the window into their union (loss/flow.py:412-414)
            last_mask = torch.cat([inp["event_mask"] for inp in inputs], 1).sum(1).clamp(max=1)
        aee_o, pct_o = oiwe.aee(flows[-1], inputs[-1]["gtflow"], last_mask, inputs[-1]["dt_gt"], inputs[-1]["dt_input"], max(H, W))
        close(pct_o.reshape(-1), pct_r.reshape(-1), 1e-6, f"metrics overwrite={overwrite} percent_AEE")
        tag = f"metrics overwrite={overwrite}"
        close(fwl_o, fwl_r, 1e-6, tag + " FWL"), close(rsat_o, rsat_r, 1e-6, tag + " RSAT")
        close(aee_o, aee_r.view(-1) if aee_r.dim() else aee_r, 1e-6, tag + " AEE")
        k = "ow" if overwrite else "seq"
        d_all.update({f"{k}_fwl": fwl_r, f"{k}_rsat": rsat_r, f"{k}_aee": aee_r.reshape(-1), f"{k}_pct": pct_r.reshape(-1)})
        if not overwrite:
            d_all.update({"events": events, "pol_mask": pol, "flows": torch.stack(flows, 1), "gtflow": inputs[-1]["gtflow"],
                          "event_masks": torch.cat([inp["event_mask"] for inp in inputs], 1), "dt_input": inputs[-1]["dt_input"],
                          "dt_gt": inputs[-1]["dt_gt"]})
        print(tag, "FWL", fwl_r.tolist(), "RSAT", rsat_r.tolist(), "AEE", aee_r.reshape(-1).tolist())
    golden["metrics"] = d_all


def pin_unet(rmodel, golden, shape=(1, 32, 48, 3)):
    """Spiking recurrent EV-FlowNet, LIF / PLIF / ALIF / XLIF cells, base_num_channels 4: T-step rollout."""
    classes = {"lif": rmodel.SpikingRecEVFlowNet, "plif": rmodel.PLIFRecEVFlowNet, "alif": rmodel.ALIFRecEVFlowNet, "xlif": rmodel.XLIFRecEVFlowNet}
    B, H, W, T = shape
    for neuron, cls in classes.items():
        torch.manual_seed(11)
        cfg = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=4, kernel_size=3,
                   activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron=None if neuron == "lif" else {})
        m = cls(cfg)
        with torch.no_grad():
            for nm, q in m.named_parameters():
                if nm.endswith("ff.weight") or nm.endswith("rec.weight"):
                    q.mul_(3.0)
                if "preds" in nm and nm.endswith("weight"):
                    q.mul_(20.0)
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        P = ounet.unet_params(sd, neuron)
        xs = []
        for t in range(T):
            ts, ys, xx, ps = oenc.synthetic_events(B, 1500, H, W, 2000 + t)
            xs.append(oenc.encode_window(ts, ys, xx, ps, H, W, 2)["event_cnt"])
        m.reset_states()
        states = [None] * 10
        acts = []
        with torch.no_grad():
            for t in range(T):
                o = m(None, xs[t].clone())
                trace = []
                preds, flows, states = ounet.unet_step(neuron, P, states, xs[t], trace=trace)
                for i in range(4):
                    close(flows[i], o["flow"][i], 0, f"unet {neuron} flow[{t}][{i}]")
                acts = [tr[3].ne(0).float().mean().item() for tr in trace]
        states_r = m.states
        for i in range(10):
            close(states[i], states_r[i], 0, f"unet {neuron} state[{i}]")
        print(f"unet {neuron}: activity per cell, last step:", ["%.3f" % a for a in acts])
        # BPTT gradient of a random linear functional of all flow scales of all steps: reference autograd vs oracle autograd
        g = torch.Generator().manual_seed(123)
        gw = [[torch.rand((B, 2, H, W), generator=g) - 0.5 for _ in range(4)] for _ in range(T)]
        named = [(n_, q) for n_, q in m.named_parameters() if q.requires_grad]
        plist = [q for _, q in named]
        m.reset_states()
        loss_r = 0.0
        for t in range(T):
            o_g = m(None, xs[t].clone())
            loss_r = loss_r + sum((f * w).sum() for f, w in zip(o_g["flow"], gw[t]))
        grads_r = torch.autograd.grad(loss_r, plist, allow_unused=True)
        Pg = ounet.unet_params(dict(m.named_parameters()) | {k: v for k, v in m.named_buffers()}, neuron)
        st = [None] * 10
        loss_o = 0.0
        for t in range(T):
            _, fl, st = ounet.unet_step(neuron, Pg, st, xs[t])
            loss_o = loss_o + sum((f * w).sum() for f, w in zip(fl, gw[t]))
        grads_o = torch.autograd.grad(loss_o, plist, allow_unused=True)
        for (nm, _), go, gr in zip(named, grads_o, grads_r):
            if gr is not None:
                close(go, gr, 1e-5, f"unet {neuron} grad {nm}")
        if golden is not None:
            d = {"x_%d" % t: xs[t] for t in range(T)}
            d.update({"flow_%d_%d" % (T - 1, i): o["flow"][i] for i in range(4)})
            d.update({"state_%d" % i: states_r[i] for i in (0, 3, 5, 9)})
            for nm, q in sd.items():
                d["sd_" + nm] = q
            for t in range(T):
                for i in range(4):
                    d["gw_%d_%d" % (t, i)] = gw[t][i]
            for (nm, _), gr in zip(named, grads_r):
                if gr is not None and (neuron == "lif" or gr.numel() <= 5000):
                    d["grad_" + nm] = gr
            d["loss"] = loss_r.detach()
            golden[f"unet_{neuron}"] = d


def pin_ann_unet(rmodel, golden, shape=(1, 32, 48)):
    """EV-FlowNet (ANN twin of the spiking U-Net), base_num_channels 4: one forward pass, all four flow scales."""
    B, H, W = shape
    torch.manual_seed(21)
    cfg = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=4, kernel_size=3,
               activations=["relu", None], mask_output=True, spiking_neuron=None)
    m = rmodel.EVFlowNet(cfg)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    ts, ys, xx, ps = oenc.synthetic_events(B, 1500, H, W, 3000)
    x = oenc.encode_window(ts, ys, xx, ps, H, W, 2)["event_cnt"]
    with torch.no_grad():
        o = m(None, x.clone())
        preds, flows = ounet.ann_unet_forward(sd, x)
    for i in range(4):
        close(flows[i], o["flow"][i], 0, f"ann unet flow[{i}]")
    print("ann EVFlowNet pinned; |flow| max %.4f" % max(f.abs().max().item() for f in flows))
    g = torch.Generator().manual_seed(41)
    gw = [torch.rand((B, 2, H, W), generator=g) - 0.5 for _ in range(4)]
    named = [(n_, q) for n_, q in m.named_parameters() if q.requires_grad]
    loss = sum((f * w).sum() for f, w in zip(m(None, x.clone())["flow"], gw))
    grads = torch.autograd.grad(loss, [q for _, q in named], allow_unused=True)
    if golden is not None:
        d = {"x": x}
        d.update({"flow_%d" % i: o["flow"][i] for i in range(4)})
        d.update({"gw_%d" % i: gw[i] for i in range(4)})
        for (nm, _), gr in zip(named, grads):
            if gr is not None:
                d["grad_" + nm] = gr
        for nm, q in sd.items():
            d["sd_" + nm] = q
        golden["annunet_evflownet"] = d


def pin_ann_rec_unet(rmodel, golden, shape=(1, 32, 48, 3)):
    """RecEVFlowNet (ConvGRU encoders), base_num_channels 4: T-step rollout, final flows, states and BPTT gradients."""
    B, H, W, T = shape
    torch.manual_seed(23)
    cfg = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=4, kernel_size=3,
               activations=["relu", None], mask_output=True, spiking_neuron=None)
    m = rmodel.RecEVFlowNet(cfg)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    xs = []
    for t in range(T):
        ts, ys, xx, ps = oenc.synthetic_events(B, 1500, H, W, 3100 + t)
        xs.append(oenc.encode_window(ts, ys, xx, ps, H, W, 2)["event_cnt"])
    m.reset_states()
    states = [None] * 4
    with torch.no_grad():
        for t in range(T):
            o = m(None, xs[t].clone())
            preds, flows = ounet.ann_unet_forward(sd, xs[t], prefix="multires_unetrec.", states=states)
            for i in range(4):
                close(flows[i], o["flow"][i], 0, f"ann rec unet flow[{t}][{i}]")
    states_r = m.states
    for i in range(4):
        close(states[i], states_r[i], 0, f"ann rec unet state[{i}]")
    g = torch.Generator().manual_seed(43)
    gw = [[torch.rand((B, 2, H, W), generator=g) - 0.5 for _ in range(4)] for _ in range(T)]
    named = [(n_, q) for n_, q in m.named_parameters() if q.requires_grad]
    m.reset_states()
    loss = 0.0
    for t in range(T):
        loss = loss + sum((f * w).sum() for f, w in zip(m(None, xs[t].clone())["flow"], gw[t]))
    grads = torch.autograd.grad(loss, [q for _, q in named], allow_unused=True)
    print("ann RecEVFlowNet pinned; |flow| max %.4f" % max(f.abs().max().item() for f in flows))
    if golden is not None:
        d = {"x_%d" % t: xs[t] for t in range(T)}
        d.update({"flow_%d" % i: o["flow"][i] for i in range(4)})
        d.update({"state_%d" % i: states_r[i] for i in range(4)})
        for t in range(T):
            for i in range(4):
                d["gw_%d_%d" % (t, i)] = gw[t][i]
        for (nm, _), gr in zip(named, grads):
            if gr is not None:
                d["grad_" + nm] = gr
        for nm, q in sd.items():
            d["sd_" + nm] = q
        golden["annunet_recevflownet"] = d


def pin_ann_zoo(rmodel, golden, shape=(1, 32, 48, 3)):
    """The rest of the import list: RNN / Leaky FireNets and EV-FlowNets, E2VID: T-step rollouts + BPTT gradients."""
    B, H, W, T = shape
    cases = (("RNNFireNet", "rnn"), ("LeakyFireNet", "leaky"), ("LeakyFireFlowNet", "leakyflow"), ("RNNRecEVFlowNet", "rnnunet"),
             ("LeakyRecEVFlowNet", "leakyunet"), ("E2VID", "e2vid"))
    xs = []
    for t in range(T):
        ts, ys, xx, ps = oenc.synthetic_events(B, 1500, H, W, 3300 + t)
        xs.append(oenc.encode_window(ts, ys, xx, ps, H, W, 2)["event_cnt"])
    for name, kind in cases:
        cls = getattr(rmodel, name)
        if hasattr(cls, "kwargs"):
            cls.kwargs = [{}] * 7
        torch.manual_seed(29)
        fire = "Fire" in name
        cfg = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=8 if fire else 4, kernel_size=3,
                   activations=["relu", None], mask_output=True, spiking_neuron={} if "Leaky" in name else None)
        m = cls(cfg)
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        n_states = {"rnn": 7, "leaky": 7, "leakyflow": 7, "rnnunet": 4, "leakyunet": 10, "e2vid": 3}[kind]

        def oracle_rollout(sdict):
            st = [None] * n_states
            outs = []
            for t in range(T):
                if kind in ("rnn", "leaky", "leakyflow"):
                    f, st = ozoo.firenet_zoo_step(kind, sdict, st, xs[t])
                    outs.append([f])
                elif kind == "rnnunet":
                    outs.append(ozoo.rnn_unet_forward(sdict, xs[t], st))
                elif kind == "leakyunet":
                    outs.append(ozoo.leaky_unet_step(sdict, st, xs[t]))
                else:
                    outs.append([ozoo.e2vid_step(sdict, st, xs[t])])
            return outs

        m.reset_states()
        with torch.no_grad():
            ref_out = [m(None, xs[t].clone())["flow"] for t in range(T)]
            ora_out = oracle_rollout(sd)
        for t in range(T):
            for i, (a, b) in enumerate(zip(ora_out[t], ref_out[t])):
                close(a, b, 0, f"{name} flow[{t}][{i}]")
        g = torch.Generator().manual_seed(47)
        gw = [[torch.rand(f.shape, generator=g) - 0.5 for f in ref_out[t]] for t in range(T)]
        named = [(n_, q) for n_, q in m.named_parameters() if q.requires_grad]
        m.reset_states()
        loss = 0.0
        for t in range(T):
            loss = loss + sum((f * w).sum() for f, w in zip(m(None, xs[t].clone())["flow"], gw[t]))
        grads = torch.autograd.grad(loss, [q for _, q in named], allow_unused=True)
        live = dict(m.named_parameters()) | dict(m.named_buffers())
        loss_o = sum((f * w).sum() for t, fl in enumerate(oracle_rollout(live)) for f, w in zip(fl, gw[t]))
        grads_o = torch.autograd.grad(loss_o, [q for _, q in named], allow_unused=True)
        for (nm, _), go, gr in zip(named, grads_o, grads):
            if gr is not None:
                close(go, gr, 1e-5, f"{name} grad {nm}")
        print(f"{name} pinned; |flow| max {max(f.abs().max().item() for f in ref_out[-1]):.4f}")
        if golden is not None:
            d = {"x_%d" % t: xs[t] for t in range(T)}
            for i, f in enumerate(ref_out[-1]):
                d["flow_%d" % i] = f
            for t in range(T):
                for i, w in enumerate(gw[t]):
                    d["gw_%d_%d" % (t, i)] = w
            for (nm, _), gr in zip(named, grads):
                if gr is not None and gr.numel() <= 20000:
                    d["grad_" + nm] = gr
            for nm, q in sd.items():
                d["sd_" + nm] = q
            golden["annzoo_" + name.lower()] = d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(1)  # fixed summation order inside mkldnn so that "bit-equal" is meaningful
    rcells, rmodel, rflow, riwe, renc = import_reference()
    golden = {}
    pin_cells(rcells, None, (2, 12, 16))  # wide check, nothing stored
    pin_cells(rcells, golden, (1, 8, 12))  # small shapes for the committed fixtures
    pin_firenet(rmodel, None, (2, 16, 24, 4))
    pin_firenet(rmodel, golden, (1, 16, 16, 3))
    pin_loss(rflow, golden)
    pin_iwe_and_encodings(riwe, renc, golden)
    pin_metrics(rflow, golden)
    pin_ann_firenet(rmodel, golden)
    pin_unet(rmodel, golden)
    pin_ann_unet(rmodel, golden)
    pin_ann_rec_unet(rmodel, golden)
    pin_ann_zoo(rmodel, golden)
    if args.check:
        print("oracle == reference on all cases (check only)")
        return
    os.makedirs(GOLD, exist_ok=True)
    manifest = {"generator": "oracle/pin_against_reference.py", "torch": torch.__version__, "reference": "tudelft/event_flow @ e81f963",
                "files": {}}
    for key, d in golden.items():
        path = os.path.join(GOLD, key + ".npz")
        np.savez_compressed(path, **{k: v.detach().numpy() for k, v in d.items()})
        manifest["files"][key + ".npz"] = {"bytes": os.path.getsize(path), "sha1": hashlib.sha1(open(path, "rb").read()).hexdigest()}
    json.dump(manifest, open(os.path.join(GOLD, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    tot = sum(v["bytes"] for v in manifest["files"].values())
    print(f"wrote {len(golden)} fixtures, {tot / 1e6:.2f} MB, to {GOLD}")


if __name__ == "__main__":
    main()
