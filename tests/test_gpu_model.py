"""
GPU parity of the LIF/PLIF/ALIF/XLIF FireNet models (T3/T4): per-step teacher-forced comparison with the CPU oracle,
golden rollouts written from the reference, BPTT gradients, state API, drop-in training loop.
"""
import glob
import os

import pytest
import torch

from oracle import encodings as oenc
from oracle import spiking as osp
from tests.conftest import GOLDEN, load_golden
from tests.util import assert_rel, compare_grads_by_layer, firenet_cfg, model_grads_by_layer, oracle_bptt_teacher_forced

pytestmark = pytest.mark.gpu
DEV = "cuda"
FIRENETS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "firenet_*.npz")))


def model_from_golden(g, neuron, enc):
    import event_flow_b200.models.model as M

    cls = {"lif": M.LIFFireNet, "plif": M.PLIFFireNet, "alif": M.ALIFFireNet, "xlif": M.XLIFFireNet}[neuron]
    bins = g["x_0"].shape[1]
    m = cls(firenet_cfg(bins, enc, neuron))
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd_")}
    m.load_state_dict(sd)  # checkpoint compatibility: the reference's state_dict loads unchanged
    return m.to(DEV)


@pytest.mark.parametrize("name", FIRENETS)
def test_firenet_rollout_and_bptt_match_reference_golden(name):
    g = load_golden(name)
    _, neuron, enc = name.split("_")
    m = model_from_golden(g, neuron, enc)
    T = sum(1 for k in g if k.startswith("x_"))
    loss = 0
    flows, spikes = [], []
    for t in range(T):
        x = g[f"x_{t}"].to(DEV)
        out = m(x, x, log=True)
        flows.append(out["flow"][0])
        loss = loss + (out["flow"][0] * g[f"gw_{t}"].to(DEV)).sum()
        assert abs(out["activity"]["7:R2b"] - g["activity"][t][7].item()) < 2e-3
        spikes.append([s[1].cpu() for s in m.states])
    states = m.states
    exact = all(torch.equal(states[i][1].cpu(), g[f"state_{i}"][1]) for i in range(7) if f"state_{i}" in g)
    loss.backward()
    mine = model_grads_by_layer(m)
    if exact:  # the free-running rollout reproduced the reference's spikes: compare with the reference's own numbers
        for t in range(T):
            torch.testing.assert_close(flows[t].detach().cpu(), g[f"flow_{t}"], rtol=1e-4, atol=1e-6)
        for n, p in m.named_parameters():
            if "grad_" + n in g and g["grad_" + n].abs().max() > 0:
                assert_rel(p.grad, g["grad_" + n], 1e-3, n)
    # ALWAYS (no skip, no condition): BPTT against the oracle's autograd on the trajectory this rollout actually took -- the oracle
    # is teacher-forced with the emitted spikes, so a borderline spike that flipped cannot switch the check off
    from tests.util import oracle_params_of

    params, kw = oracle_params_of(m), {}
    _, ref, flows_o = oracle_bptt_teacher_forced(neuron, params, [g[f"x_{t}"] for t in range(T)], spikes,
                                                 lambda fl: sum((f * g[f"gw_{t}"]).sum() for t, f in enumerate(fl)), **kw)
    for t in range(T):
        torch.testing.assert_close(flows[t].detach().cpu(), flows_o[t], rtol=1e-4, atol=1e-6)
    compare_grads_by_layer(mine, ref, 1e-3)


@pytest.mark.parametrize("neuron", osp.NEURONS)
def test_firenet_teacher_forced_layers_match_oracle(neuron):
    """Free-running CUDA rollout; every layer of every step is re-computed by the oracle from the inputs the CUDA path fed it."""
    import event_flow_b200.models.model as M
    from tests.util import capture_layers, oracle_params_of

    cls = {"lif": M.LIFFireNet, "plif": M.PLIFFireNet, "alif": M.ALIFFireNet, "xlif": M.XLIFFireNet}[neuron]
    B, H, W, T, bins = 2, 48, 64, 5, 5
    torch.manual_seed(1)
    m = cls(firenet_cfg(bins, "voxel", neuron))
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        m.pred.conv2d.weight.mul_(20.0)
    params = oracle_params_of(m)
    m = m.to(DEV)
    captured, handles = capture_layers(m)
    flips_in = total = 0
    for t in range(T):
        d = oenc.encode_window(*oenc.synthetic_events(B, 600, H, W, 300 + t), H, W, bins)
        out = m(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV), log=True)
        for name, dv, out_flips, in_flips, n in osp.firenet_layerwise_check(neuron, params, captured):
            assert dv < 2e-5, (t, name, dv)
            assert out_flips == 0, (t, name, out_flips)
            flips_in += in_flips
            total += n
        x7 = captured["R2b"][2]
        torch.testing.assert_close(out["flow"][0].detach().cpu(), osp.pred_head(x7, params["pred"]["weight"], params["pred"]["bias"]),
                                   rtol=1e-5, atol=1e-7)
        assert out["activity"]["4:R1b"] > 0.01 and out["activity"]["7:R2b"] > 0.01  # spikes propagate: not vacuous
    for h in handles:
        h.remove()
    assert flips_in <= 1e-5 * total + 2, f"{flips_in} in-band spike flips of {total}"


@pytest.mark.parametrize("neuron", ["lif", "alif"])
def test_state_api_reset_detach_set(neuron):
    import event_flow_b200.models.model as M

    cls = {"lif": M.LIFFireNet, "alif": M.ALIFFireNet}[neuron]
    n_state = 2 if neuron == "lif" else 3
    m = cls(firenet_cfg(2, "cnt", neuron)).to(DEV)
    assert m.states == [None] * 7
    x = torch.randint(0, 3, (1, 2, 32, 32)).float().to(DEV)
    m(x, x)
    s = m.states
    assert len(s) == 7 and s[0].shape == (n_state, 1, 32, 32, 32) and s[0].dtype == torch.float32
    assert set(s[3][1].unique().tolist()) <= {0.0, 1.0}
    before = m.states[0].clone()
    s[0].zero_()  # clones: mutating them must not touch the model (model_util.py:96-102)
    assert torch.equal(m.states[0], before)
    m.detach_states()
    assert all(not t.requires_grad for t in m.states)
    m.states = [torch.zeros_like(t) for t in s]
    out = m(x, x)
    m.reset_states()
    out2 = m(x, x)
    torch.testing.assert_close(out["flow"][0], out2["flow"][0])  # zero state == reset state
    # state round trip: setting the states read back continues the rollout identically
    a = m(x, x)["flow"][0].clone()
    saved = m.states
    b = m(x, x)["flow"][0].clone()
    m.states = saved
    c = m(x, x)["flow"][0].clone()
    assert torch.equal(b, c) and a.shape == b.shape
    with pytest.raises(AttributeError):
        cls(firenet_cfg(5, "cnt", neuron)).to(DEV)(x, x)  # cnt encoding needs num_bins == 2 (model.py:240-244)


def test_truncated_bptt_matches_oracle_over_two_windows():
    """detach_states() cuts the gradient at the window boundary exactly like the reference (train_flow.py:170, model.py:211-221)."""
    import event_flow_b200.models.model as M
    from tests.util import oracle_params_of

    B, Hh, Ww, T, bins = 1, 16, 16, 3, 5
    torch.manual_seed(3)
    m = M.LIFFireNet(firenet_cfg(bins, "voxel"))
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        m.pred.conv2d.weight.mul_(20.0)
    params = oracle_params_of(m)
    m = m.to(DEV)
    g = torch.Generator().manual_seed(0)
    xs = [oenc.encode_window(*oenc.synthetic_events(B, 300, Hh, Ww, 500 + t), Hh, Ww, bins)["event_voxel"] for t in range(2 * T)]
    gw = [torch.rand((B, 2, Hh, Ww), generator=g) - 0.5 for _ in range(2 * T)]
    for win in range(2):
        states0 = None if win == 0 else [s.cpu() for s in m.states]  # the window starts from the (detached) states of the previous one
        loss, spikes = 0, []
        for t in range(win * T, (win + 1) * T):
            loss = loss + (m(xs[t].to(DEV), None)["flow"][0] * gw[t].to(DEV)).sum()
            spikes.append([s[1].cpu() for s in m.states])
        m.zero_grad()
        loss.backward()
        _, ref, _ = oracle_bptt_teacher_forced("lif", params, xs[win * T:(win + 1) * T], spikes,
                                               lambda fl: sum((f * gw[win * T + t]).sum() for t, f in enumerate(fl)), states0=states0)
        compare_grads_by_layer(model_grads_by_layer(m), ref, 1e-3)  # unconditional; north-star tolerance
        m.detach_states()


def test_dropin_training_loop_like_train_flow():
    """The loop of train_flow.py:97-171 with this package's classes resolved under the reference's import names."""
    import sys

    import event_flow_b200

    event_flow_b200.install_dropin()
    from loss.flow import EventWarping  # noqa
    from models.model import LIFFireNet  # noqa

    H = W = 64
    B, N, T = 2, 500, 4
    torch.manual_seed(0)
    model = LIFFireNet(firenet_cfg(2, "cnt")).to(DEV)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        model.pred.conv2d.weight.mul_(50.0)
    model.train()
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False, "clip_grad": 100.0},
           "model": {"mask_output": True}, "data": {"window_loss": N * T}}
    loss_function = EventWarping(cfg, DEV)
    optimizer = torch.optim.Adam(model.parameters(), lr=2e-4)
    before = [p.detach().clone() for p in model.parameters()]
    losses = []
    for it in range(2 * T):
        d = oenc.encode_window(*oenc.synthetic_events(B, N, H, W, 900 + it), H, W, 2)
        x = model(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV))
        loss_function.event_flow_association(x["flow"], d["event_list"].to(DEV), d["event_list_pol_mask"].to(DEV), d["event_mask"].to(DEV))
        if loss_function.num_events >= cfg["data"]["window_loss"]:
            loss = loss_function()
            losses.append(loss.item())
            loss.backward()
            torch.nn.utils.clip_grad.clip_grad_norm_(model.parameters(), cfg["loss"]["clip_grad"])
            optimizer.step()
            optimizer.zero_grad()
            model.detach_states()
            loss_function.reset()
    assert len(losses) == 2 and all(l == l and l > 0 for l in losses)
    assert any((a - b.detach()).abs().max() > 0 for a, b in zip(before, model.parameters()))
    for s in [k for k in sys.modules if k.split(".")[0] in ("models", "loss", "utils", "dataloader")]:
        sys.modules.pop(s, None)


@pytest.mark.parametrize("name,recurrent", [("ann_firenet", True), ("ann_fireflownet", False)])
def test_ann_firenet_forward_matches_reference_golden(name, recurrent):
    """BASELINE cfg 1: ANN FireNet (ConvLayer_ + ConvGRU) forward, 1 voxel bin, batch 1; and FireFlowNet.  fp32 CUDA cores: 1e-4."""
    import event_flow_b200.models.model as M

    g = load_golden(name)
    bins = g["x_0"].shape[1]
    cls = M.FireNet if recurrent else M.FireFlowNet
    cfg = dict(name="x", encoding="voxel" if bins == 1 else "cnt", round_encoding=False, norm_input=False, num_bins=bins, base_num_channels=32,
               kernel_size=3, activations=["relu", None], mask_output=True, spiking_neuron=None)
    m = cls(cfg)
    m.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("sd_")})
    m = m.to(DEV).eval()
    T = sum(1 for k in g if k.startswith("x_"))
    with torch.no_grad():
        for t in range(T):
            x = g[f"x_{t}"].to(DEV)
            out = m(x, x)
            torch.testing.assert_close(out["flow"][0].cpu(), g[f"flow_{t}"], rtol=1e-4, atol=1e-6)
    if recurrent:
        for i in (1, 4):
            torch.testing.assert_close(m.states[i].cpu(), g[f"state_{i}"], rtol=1e-4, atol=1e-6)
    # BPTT through the ANN cells (ConvLayer_ / ConvGRU): gradients vs the reference's autograd, 1e-3 relative to each tensor's scale
    m.train()
    m.reset_states()
    loss = 0.0
    for t in range(T):
        x = g[f"x_{t}"].to(DEV)
        loss = loss + (m(x, x)["flow"][0] * g[f"gw_{t}"].to(DEV)).sum()
    loss.backward()
    checked = 0
    for nm, q in m.named_parameters():
        if "grad_" + nm in g:
            ref = g["grad_" + nm]
            scale = ref.abs().max().item() + 1e-12
            err = (q.grad.cpu() - ref).abs().max().item()
            assert err <= 1e-3 * scale, f"{nm}: {err:.3e} vs scale {scale:.3e}"
            checked += 1
    assert checked >= 10


# ---------------------------------------------------------------------------------------------------------------------
# config-size BPTT (cfg 3: B=8, 128x128, T=10) through the graph-replayed fast path and the real loss, against oracle autograd
# ---------------------------------------------------------------------------------------------------------------------
def test_config_size_bptt_matches_oracle_autograd():
    """
    The path bench.py times: LIFFireNet fast path (CUDA-graph forward steps, cached + graph-replayed backward calls), the
    EventWarping loss kernels, BPTT over a full cfg-3 window (B=8, 128x128, T=10, 1000 events per step).  Three windows are run so
    that the third one executes entirely from cached / replayed launches; its parameter gradients are compared with the CPU
    oracle's autograd of the SAME window (initial states = the states the CUDA path carried over, spikes and flow values
    teacher-forced -- see tests/util.py:oracle_bptt_teacher_forced for why --, loss = the oracle's event-warping loss):
    rel 1e-3 per parameter tensor.
    """
    import event_flow_b200.models.model as M
    from event_flow_b200.loss.flow import EventWarping
    from oracle import iwe as oiwe
    from tests.util import oracle_params_of

    B, H, W, T, N, bins = 8, 128, 128, 10, 1000, 5
    torch.manual_seed(0)
    m = M.LIFFireNet(firenet_cfg(bins, "voxel"))
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        m.pred.conv2d.weight.mul_(20.0)
    params = oracle_params_of(m)
    m = m.to(DEV).train()
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False}, "model": {"mask_output": True}}
    lossf = EventWarping(cfg, DEV)
    for win in range(3):
        states0 = None if win == 0 else [s.cpu() for s in m.states]
        lossf.reset()
        m.zero_grad(set_to_none=True)
        data, spikes, flows = [], [], []
        for t in range(T):
            d = oenc.encode_window(*oenc.synthetic_events(B, N, H, W, 7000 + 100 * win + t), H, W, bins)
            data.append(d)
            out = m(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV))
            lossf.event_flow_association(out["flow"], d["event_list"].clone().to(DEV), d["event_list_pol_mask"].to(DEV), d["event_mask"].to(DEV))
            if win == 2:
                spikes.append([s[1].cpu() for s in m.states])
                flows.append(out["flow"][0].detach().cpu())
        loss = lossf()
        loss.backward()
        if win < 2:
            m.detach_states()
    assert sum(s[6].mean().item() for s in spikes) > 0.01 * T, "vacuous: no spikes reach the prediction layer"

    def oracle_loss(flows):
        evs = []
        for t, d in enumerate(data):
            e = d["event_list"].clone()
            e[:, :, 0] += t
            evs.append(e)
        return oiwe.event_warping_loss(torch.cat(evs, 1), torch.cat([d["event_list_pol_mask"] for d in data], 1), torch.arange(T).repeat_interleave(N),
                                       [torch.stack(flows, 1)], torch.cat([d["event_mask"] for d in data], 1), (H, W), weight=0.001, passes=T)

    loss_o, ref, flows_o = oracle_bptt_teacher_forced("lif", params, [d["event_voxel"] for d in data], spikes, oracle_loss, states0=states0,
                                                      forced_flows=flows)
    for f, fo in zip(flows, flows_o):
        torch.testing.assert_close(f, fo, rtol=1e-5, atol=1e-7)
    assert_rel(loss, loss_o, 1e-5, "loss of the config-size window")
    worst = compare_grads_by_layer(model_grads_by_layer(m), ref, 1e-3)
    print(f"config-size BPTT: worst relative gradient error {worst:.2e}")


# ---------------------------------------------------------------------------------------------------------------------
# T4: free-running rollout statistics on an MVSEC-shape structured stream (256x256, known affine flow): activity and AEE
# ---------------------------------------------------------------------------------------------------------------------
def affine_dot_stream(B, H, W, T, n_events, seed):
    """
    Random dots translating with a known affine flow u(x,y) = a + A [x,y] (pixels per window): per step T event tuples
    (ts,y,x,p) [B,N,4] on integer pixels and the ground-truth flow map [B,2,H,W] (x, y channels), dt_gt = dt_input.
    """
    g = torch.Generator().manual_seed(seed)
    a = (torch.rand((B, 2), generator=g) - 0.5) * 6.0               # translation, px / window
    A = (torch.rand((B, 2, 2), generator=g) - 0.5) * 0.02           # shear / zoom
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    gt = torch.stack([a[:, 0, None, None] + A[:, 0, 0, None, None] * (xx - W / 2) + A[:, 0, 1, None, None] * (yy - H / 2),
                      a[:, 1, None, None] + A[:, 1, 0, None, None] * (xx - W / 2) + A[:, 1, 1, None, None] * (yy - H / 2)], 1)
    n_dots = 400
    px = torch.rand((B, n_dots), generator=g) * (W - 40) + 20
    py = torch.rand((B, n_dots), generator=g) * (H - 40) + 20
    pol = (torch.randint(0, 2, (B, n_dots), generator=g) * 2 - 1).float()
    steps = []
    for _ in range(T):
        ts = torch.sort(torch.rand((B, n_events), generator=g))[0]
        ts = (ts - ts[:, :1]) / (ts[:, -1:] - ts[:, :1])
        which = torch.randint(0, n_dots, (B, n_events), generator=g)
        x0, y0, p = torch.gather(px, 1, which), torch.gather(py, 1, which), torch.gather(pol, 1, which)
        u = a[:, 0:1] + A[:, 0, 0:1] * (x0 - W / 2) + A[:, 0, 1:2] * (y0 - H / 2)
        v = a[:, 1:2] + A[:, 1, 0:1] * (x0 - W / 2) + A[:, 1, 1:2] * (y0 - H / 2)
        xs = torch.round(x0 + u * ts).clamp(0, W - 1)
        ys = torch.round(y0 + v * ts).clamp(0, H - 1)
        steps.append(torch.stack([ts, ys, xs, p], 2))
        ud = a[:, 0:1] + A[:, 0, 0:1] * (px - W / 2) + A[:, 0, 1:2] * (py - H / 2)
        vd = a[:, 1:2] + A[:, 1, 0:1] * (px - W / 2) + A[:, 1, 1:2] * (py - H / 2)
        px, py = (px + ud).clamp(10, W - 10), (py + vd).clamp(10, H - 10)
    return steps, gt


def test_rollout_statistics_and_aee_on_mvsec_shape_stream():
    """
    SURVEY T4 / north-star: a FREE-RUNNING T=20 rollout on a 256x256 (configs/eval_MVSEC.yml:16) structured stream.  Pointwise
    parity of a free-running spiking net is impossible for any implementation that is not bit-identical in summation order (SURVEY
    7.3), the statistics are robust: per-layer activity (model(..., log=True), models/model.py:268-284) within 1e-2 absolute of
    the oracle's rollout at every step, AEE (loss/flow.py:582-628, through the drop-in AEE class and ef_aee) averaged over the
    rollout within 1e-3 relative of the oracle's.
    """
    import event_flow_b200.models.model as M
    from event_flow_b200.loss.flow import AEE
    from oracle import iwe as oiwe
    from tests.util import oracle_params_of

    B, H, W, T, N, bins = 2, 256, 256, 20, 4000, 5
    torch.manual_seed(2)
    m = M.LIFFireNet(firenet_cfg(bins, "voxel"))
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        m.pred.conv2d.weight.mul_(20.0)
    params = oracle_params_of(m)
    m = m.to(DEV).eval()
    steps, gt = affine_dot_stream(B, H, W, T, N, seed=21)
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"overwrite_intermediate": False}}
    metric = AEE(cfg, DEV, flow_scaling=max(H, W))
    states = [None] * 7
    aee_mine, aee_ref, worst_act = [], [], 0.0
    names = ["1:head", "2:G1", "3:R1a", "4:R1b", "5:G2", "6:R2a", "7:R2b"]
    with torch.no_grad():
        for t, ev in enumerate(steps):
            d = oenc.encode_window(ev[:, :, 0], ev[:, :, 1], ev[:, :, 2], ev[:, :, 3], H, W, bins)
            out = m(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV), log=True)
            flow_o, states, acts = osp.firenet_step("lif", params, states, d["event_voxel"])
            for name, a in zip(names, acts):
                worst_act = max(worst_act, abs(out["activity"][name] - a.ne(0).float().mean().item()))
            inputs = {"event_list": d["event_list"], "event_list_pol_mask": d["event_list_pol_mask"], "event_mask": d["event_mask"], "gtflow": gt,
                      "dt_input": torch.ones(B), "dt_gt": torch.ones(B)}
            metric.reset()
            metric.event_flow_association(out["flow"], inputs)
            aee, _ = metric()
            aee_o, _ = oiwe.aee(flow_o, gt, d["event_mask"][:, 0], torch.ones(B, 1, 1, 1), torch.ones(B, 1, 1, 1), float(max(H, W)))
            aee_mine.append(aee.mean().item())
            aee_ref.append(aee_o.mean().item())
    assert min(out["activity"][n] for n in names) > 0.005, "vacuous rollout: a layer went silent"
    mean_mine, mean_ref = sum(aee_mine) / T, sum(aee_ref) / T
    rel = abs(mean_mine - mean_ref) / mean_ref
    print(f"T4: worst per-layer activity difference {worst_act:.2e}; averaged AEE {mean_mine:.5f} vs oracle {mean_ref:.5f} (rel {rel:.2e})")
    assert worst_act <= 1e-2
    assert rel <= 1e-3


# ---------------------------------------------------------------------------------------------------------------------
# checkpointing a model that has run on the fast path; the arena's recycling contract
# ---------------------------------------------------------------------------------------------------------------------
def test_model_pickles_and_deepcopies_after_fast_path_training(tmp_path):
    """
    The reference checkpoints by pickling the whole module (utils/utils.py:36, mlflow.pytorch.log_model).  After a forward +
    backward on the fast path the model holds CUDA graphs, ctypes argument structs and activation slabs: none of them may end up
    in the pickle, and the neuron states must survive in the reference's stacked format.
    """
    import copy

    import event_flow_b200.models.model as M

    torch.manual_seed(0)
    m = M.LIFFireNet(firenet_cfg(5, "voxel"))
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        m.pred.conv2d.weight.mul_(20.0)
    m = m.to(DEV)
    xs = [oenc.encode_window(*oenc.synthetic_events(2, 400, 32, 48, 60 + t), 32, 48, 5)["event_voxel"].to(DEV) for t in range(5)]
    for _ in range(2):  # second window: backward runs from cached / graph-replayed calls
        loss = sum(m(x, None)["flow"][0].square().sum() for x in xs[:2])
        loss.backward()
        m.detach_states()
    path = tmp_path / "model.pth"
    torch.save(m, path)
    state_bytes = sum(t.numel() * 4 for t in m.states) + sum(p.numel() * 4 for p in m.parameters())
    assert path.stat().st_size < state_bytes + 500_000, "run-time caches (activation slabs, graphs) leaked into the checkpoint"
    loaded = torch.load(path, weights_only=False)
    twin = copy.deepcopy(m)
    for other in (loaded, twin):  # the neuron states travel in the reference's stacked format
        assert all(torch.equal(x, y) for x, y in zip(other.states, m.states)) and other.states[3].shape == (2, 2, 32, 32, 48)
    a = m(xs[2], None)["flow"][0]
    for other in (loaded, twin):
        b = other(xs[2], None)["flow"][0]
        assert torch.equal(a, b), "a restored model continues the rollout bit-identically (same states, same kernels)"
        assert set(other.state_dict()) == set(m.state_dict())


def test_arena_refuses_backward_after_its_activations_were_recycled():
    """fast._Arena contract: backward of window w must run before window w+1 completes; violating it raises instead of silently
    differentiating overwritten activations."""
    import event_flow_b200.models.model as M

    torch.manual_seed(0)
    m = M.LIFFireNet(firenet_cfg(5, "voxel")).to(DEV)
    xs = [oenc.encode_window(*oenc.synthetic_events(1, 300, 32, 32, 80 + t), 32, 32, 5)["event_voxel"].to(DEV) for t in range(2)]
    loss_a = sum(m(x, None)["flow"][0].square().sum() for x in xs)
    m.detach_states()
    loss_b = sum(m(x, None)["flow"][0].square().sum() for x in xs)  # window w+1 completes: it rewrote the slot window w started from ...
    m.detach_states()
    loss_c = sum(m(x, None)["flow"][0].square().sum() for x in xs)  # ... and window w+2 rewrites window w's own slots
    with pytest.raises(RuntimeError, match="overwritten"):
        loss_a.backward()
    loss_c.backward()  # the current window is fine
    del loss_b
