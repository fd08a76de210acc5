"""CPU: the oracle restatement reproduces the reference outputs stored in tests/golden/ (parity pinning)."""
import glob
import json
import os

import pytest
import torch

from oracle import encodings as oenc
from oracle import iwe as oiwe
from oracle import spiking as osp
from tests.conftest import GOLDEN, load_golden

torch.set_num_threads(1)

CELLS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "cell_*.npz")))
FIRENETS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "firenet_*.npz")))
LOSSES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "loss_*.npz")))


def test_manifest_lists_every_fixture():
    man = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))
    on_disk = {os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*.npz"))}
    assert on_disk == set(man["files"]) and len(on_disk) >= 20


def cell_params(g):
    return {k[2:]: v for k, v in g.items() if k.startswith("p_")}


@pytest.mark.parametrize("name", CELLS)
def test_cell_step_matches_reference(name):
    g = load_golden(name)
    _, neuron, _, reset, _ = name.split("_")
    x = g["x"].requires_grad_(True)
    st = g["state"].requires_grad_(True)
    p = {k: v.requires_grad_(True) for k, v in cell_params(g).items()}
    out, ns = osp.cell_step(neuron, x, st, p, hard_reset=(reset == "hard"), width=float(g["width"]))
    assert torch.equal(out, g["out"]) and torch.equal(ns, g["new_state"])
    ((out * g["g_out"]).sum() + (ns * g["g_state"]).sum()).backward()
    torch.testing.assert_close(x.grad, g["grad_x"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(st.grad, g["grad_state"], rtol=1e-5, atol=1e-7)
    for k, v in p.items():
        if "grad_" + k in g:
            torch.testing.assert_close(v.grad, g["grad_" + k], rtol=1e-4, atol=1e-6)


def firenet_params(g, neuron):
    params = {}
    for layer in osp.FIRENET_LAYERS:
        p = {"ff": g[f"sd_{layer}.ff.weight"]}
        if f"sd_{layer}.rec.weight" in g:
            p["rec"] = g[f"sd_{layer}.rec.weight"]
        for k in ("leak", "thresh", "leak_v", "leak_pt", "leak_t", "add_pt", "t0", "t1"):
            if f"sd_{layer}.{k}" in g:
                p[k] = g[f"sd_{layer}.{k}"]
        params[layer] = p
    params["pred"] = {"weight": g["sd_pred.conv2d.weight"], "bias": g["sd_pred.conv2d.bias"]}
    return params


@pytest.mark.parametrize("name", FIRENETS)
def test_firenet_rollout_matches_reference(name):
    g = load_golden(name)
    neuron = name.split("_")[1]
    params = firenet_params(g, neuron)
    leaves = []
    for lp in params.values():
        for k in lp:
            lp[k] = lp[k].clone().requires_grad_(True)
            leaves.append(lp[k])
    T = sum(1 for k in g if k.startswith("x_"))
    states, loss = [None] * 7, 0
    for t in range(T):
        flow, states, _ = osp.firenet_step(neuron, params, states, g[f"x_{t}"])
        assert torch.equal(flow, g[f"flow_{t}"])
        loss = loss + (flow * g[f"gw_{t}"]).sum()
    for i in range(7):
        if f"state_{i}" in g:
            assert torch.equal(states[i], g[f"state_{i}"])
    loss.backward()
    for layer, lp in params.items():
        for k, v in lp.items():
            key = f"grad_{layer}.{k}" + (".weight" if k in ("ff", "rec") else "") if layer != "pred" else f"grad_pred.conv2d.{k}"
            if key in g:
                torch.testing.assert_close(v.grad, g[key], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", LOSSES)
def test_event_warping_loss_matches_reference(name):
    g = load_golden(name)
    scaling, smask, overwrite, weight, T, N = g["cfg"].tolist()
    T = int(T)
    flow = g["flow"].requires_grad_(True)  # [B,T,2,H,W]
    H, W = flow.shape[-2:]
    if overwrite:
        fm = [flow[:, -1:]]
        em = g["event_mask"].sum(1, keepdim=True).clamp(max=1)
    else:
        fm, em = [flow], g["event_mask"]
    loss = oiwe.event_warping_loss(g["events"], g["pol_mask"], g["pass_of_event"].long(), fm, em, (H, W), weight=weight,
                                   loss_scaling=bool(scaling), smoothing_mask=bool(smask), overwrite_intermediate=bool(overwrite), passes=T)
    torch.testing.assert_close(loss.detach(), g["loss"], rtol=1e-6, atol=0)
    loss.backward()
    for t in range(T):
        if f"grad_{t}" in g:
            torch.testing.assert_close(flow.grad[:, t], g[f"grad_{t}"], rtol=1e-5, atol=1e-8)


def test_iwe_image_and_interpolation_match_reference():
    g = load_golden("iwe_image")
    H, W = g["flow"].shape[-2:]
    pm = g["pol_mask"]
    for rnd, key in ((True, "iwe_round"), (False, "iwe_bilinear")):
        out = oiwe.pol_iwe(g["flow"], g["events"], (H, W), pm[:, :, 0:1], pm[:, :, 1:2], flow_scaling=max(H, W), round_idx=rnd)
        assert torch.equal(out, g[key])
    ev_flow = oiwe.gather_event_flow(g["flow"], g["events"], (H, W))
    for tref in (0, 1):
        idx, w = oiwe.warp_and_split(g["events"], ev_flow, tref, (H, W), max(H, W))
        assert torch.equal(idx, g[f"idx_tref{tref}"]) and torch.equal(w, g[f"w_tref{tref}"])


def test_encodings_match_reference():
    g = load_golden("encodings")
    B = g["ts"].shape[0]
    H, W = g["cnt_0"].shape[-2:]
    for b in range(B):
        a = (g["xs"][b], g["ys"][b], g["ps"][b])
        assert torch.equal(oenc.events_to_channels(*a, (H, W)), g[f"cnt_{b}"])
        assert torch.equal(oenc.event_mask(*a, (H, W))[0], g[f"mask_{b}"])
        for bins in (2, 5):
            assert torch.equal(oenc.events_to_voxel(g["xs"][b], g["ys"][b], g["ts"][b], g["ps"][b], bins, (H, W)), g[f"voxel{bins}_{b}"])


def test_surrogate_curves_closed_form():
    # models/spiking_util.py:112-141 (__main__ curves)
    x = torch.linspace(-5, 5, 1001)
    assert torch.allclose(osp.surrogate_grad(x, 10.0, "arctanspike"), 1 / (1 + 10 * x * x))
    assert torch.allclose(osp.surrogate_grad(x, 10.0, "superspike"), 1 / (1 + 10 * x.abs()) ** 2)
    assert torch.allclose(osp.surrogate_grad(x, 1.0, "trianglespike"), torch.relu(1 - x.abs()))


def test_validation_metrics_match_reference():
    g = load_golden("metrics")
    B, T = g["flows"].shape[:2]
    H, W = g["flows"].shape[-2:]
    N = g["events"].shape[1] // T
    for key, overwrite in (("seq", False), ("ow", True)):
        ev_flow = torch.cat([oiwe.gather_event_flow(g["flows"][:, -1 if overwrite else t], g["events"][:, t * N:(t + 1) * N], (H, W))
                             for t in range(T)], dim=1)
        fwl, rsat = oiwe.fwl_rsat(g["events"], g["pol_mask"], ev_flow, T, (H, W), max(H, W))
        torch.testing.assert_close(fwl, g[f"{key}_fwl"], rtol=1e-6, atol=0)
        torch.testing.assert_close(rsat, g[f"{key}_rsat"], rtol=1e-6, atol=0)
        mask = g["event_masks"].sum(1).clamp(max=1) if overwrite else g["event_masks"][:, -1]
        aee, pct = oiwe.aee(g["flows"][:, -1], g["gtflow"], mask, g["dt_gt"], g["dt_input"], max(H, W))
        torch.testing.assert_close(aee, g[f"{key}_aee"], rtol=1e-6, atol=0)
        torch.testing.assert_close(pct.reshape(-1), g[f"{key}_pct"], rtol=1e-6, atol=0)


def ann_params_from_golden(g, recurrent):
    params = {}
    for name in osp.FIRENET_LAYERS:
        if recurrent and name in osp.FIRENET_RECURRENT:
            params[name] = {f"{k}_{s}": g[f"sd_{name}.{k}_gate.{'weight' if s == 'w' else 'bias'}"] for k in ("update", "reset", "out") for s in ("w", "b")}
        else:
            params[name] = {"w": g[f"sd_{name}.conv2d.weight"], "b": g[f"sd_{name}.conv2d.bias"]}
    params["pred"] = {"weight": g["sd_pred.conv2d.weight"], "bias": g["sd_pred.conv2d.bias"]}
    return params


@pytest.mark.parametrize("name,recurrent", [("ann_firenet", True), ("ann_fireflownet", False)])
def test_ann_firenet_matches_reference(name, recurrent):
    g = load_golden(name)
    params = ann_params_from_golden(g, recurrent)
    states = [None] * 7
    T = sum(1 for k in g if k.startswith("x_"))
    for t in range(T):
        flow, states, _ = osp.firenet_ann_step(params, states, g[f"x_{t}"], recurrent=recurrent)
        assert torch.equal(flow, g[f"flow_{t}"])


UNETS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "unet_*.npz")))


@pytest.mark.parametrize("name", UNETS)
def test_spiking_unet_rollout_matches_reference(name):
    """the oracle's U-Net restatement reproduces the reference's SpikingRecEVFlowNet family bit for bit."""
    from oracle import unet as ounet

    g = load_golden(name)
    neuron = name.split("_")[1]
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd_")}
    P = ounet.unet_params(sd, neuron)
    T = len([k for k in g if k.startswith("x_")])
    states = [None] * 10
    with torch.no_grad():
        for t in range(T):
            preds, flows, states = ounet.unet_step(neuron, P, states, g["x_%d" % t])
    for i in range(4):
        assert torch.equal(flows[i], g["flow_%d_%d" % (T - 1, i)]), f"flow scale {i}"
    for i in (0, 3, 5, 9):
        assert torch.equal(states[i], g["state_%d" % i]), f"state {i}"


def test_ann_evflownet_matches_reference():
    from oracle import unet as ounet

    g = load_golden("annunet_evflownet")
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd_")}
    with torch.no_grad():
        _, flows = ounet.ann_unet_forward(sd, g["x"])
    for i in range(4):
        assert torch.equal(flows[i], g["flow_%d" % i])


def test_ann_recevflownet_matches_reference():
    from oracle import unet as ounet

    g = load_golden("annunet_recevflownet")
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd_")}
    T = len([k for k in g if k.startswith("x_")])
    states = [None] * 4
    with torch.no_grad():
        for t in range(T):
            _, flows = ounet.ann_unet_forward(sd, g["x_%d" % t], prefix="multires_unetrec.", states=states)
    for i in range(4):
        assert torch.equal(flows[i], g["flow_%d" % i]) and torch.equal(states[i], g["state_%d" % i])


ZOO = {"rnnfirenet": "rnn", "leakyfirenet": "leaky", "leakyfireflownet": "leakyflow", "rnnrecevflownet": "rnnunet",
       "leakyrecevflownet": "leakyunet", "e2vid": "e2vid"}


@pytest.mark.parametrize("name", sorted(ZOO))
def test_ann_zoo_rollout_matches_reference(name):
    """the oracle restatement of the remaining ANN cell zoo reproduces the reference's rollouts bit for bit."""
    from oracle import annzoo as ozoo

    g = load_golden("annzoo_" + name)
    kind = ZOO[name]
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd_")}
    T = len([k for k in g if k.startswith("x_")])
    st = [None] * {"rnn": 7, "leaky": 7, "leakyflow": 7, "rnnunet": 4, "leakyunet": 10, "e2vid": 3}[kind]
    with torch.no_grad():
        for t in range(T):
            x = g["x_%d" % t]
            if kind in ("rnn", "leaky", "leakyflow"):
                f, st = ozoo.firenet_zoo_step(kind, sd, st, x)
                flows = [f]
            elif kind == "rnnunet":
                flows = ozoo.rnn_unet_forward(sd, x, st)
            elif kind == "leakyunet":
                flows = ozoo.leaky_unet_step(sd, st, x)
            else:
                flows = [ozoo.e2vid_step(sd, st, x)]
    for i, f in enumerate(flows):
        assert torch.equal(f, g["flow_%d" % i])
