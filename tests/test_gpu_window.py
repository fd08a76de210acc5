"""
GPU: the windowed, time-fused entry point of the LIF FireNet (model.forward_window -> ef_lif_conv_fwd_window) against the step-by-step
path of the same model (train_flow.py:97-141 / models/model.py:255-265 call the model once per step).  The fused launches run the same
arithmetic in the same order with the state kept in registers instead of memory, so flows and states must be BIT-equal (gradients: same backward on the same activations);
the step-by-step path itself is pinned against the reference in test_gpu_model.py.
"""
import pytest
import torch

from oracle import encodings as oenc
from tests.util import firenet_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(bins=5, seed=0):
    from event_flow_b200.models.model import LIFFireNet

    torch.manual_seed(seed)
    m = LIFFireNet(firenet_cfg(bins, "voxel"))
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        m.pred.conv2d.weight.mul_(20.0)
    return m.to(DEV)


def _windows(B, N, H, W, T, bins, seed):
    return [oenc.encode_window(*oenc.synthetic_events(B, N, H, W, seed + t), H, W, bins) for t in range(T)]


@pytest.mark.parametrize("shape", [(2, 32, 48, 4), (1, 40, 36, 3), (3, 128, 128, 10)])
def test_window_forward_is_the_stepwise_forward(shape):
    """Two consecutive windows (the second starts from a non-zero state), no gradient: flows and final states bit-equal."""
    B, H, W, T = shape
    wins = _windows(B, 400, H, W, 2 * T, 5, 70)
    a, b = _model(), _model()
    with torch.no_grad():
        for w in range(2):
            part = wins[w * T:(w + 1) * T]
            vox = torch.stack([d["event_voxel"] for d in part]).to(DEV)
            cnt = torch.stack([d["event_cnt"] for d in part]).to(DEV)
            outs = a.forward_window(vox, cnt)
            assert len(outs) == T
            for t, d in enumerate(part):
                ref = b(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV))["flow"][0]
                assert torch.equal(outs[t]["flow"][0], ref), f"window {w} step {t}"
            for i, (sa, sb) in enumerate(zip(a.states, b.states)):
                assert torch.equal(sa, sb), f"window {w}: state of layer {i}"
    assert outs[-1]["flow"][0].abs().max() > 0  # spikes reach the prediction layer


def test_window_then_steps_continue_the_same_sequence():
    """A window through forward_window followed by ordinary steps: the state hand-over works in both directions."""
    B, H, W, T = 2, 32, 48, 3
    wins = _windows(B, 400, H, W, 3 * T, 5, 90)
    a, b = _model(), _model()
    with torch.no_grad():
        flows_b = [b(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV))["flow"][0] for d in wins]
        flows_a = [a(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV))["flow"][0] for d in wins[:T]]
        part = wins[T:2 * T]
        outs = a.forward_window(torch.stack([d["event_voxel"] for d in part]).to(DEV), torch.stack([d["event_cnt"] for d in part]).to(DEV))
        flows_a += [o["flow"][0] for o in outs]
        flows_a += [a(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV))["flow"][0] for d in wins[2 * T:]]
    for t, (fa, fb) in enumerate(zip(flows_a, flows_b)):
        assert torch.equal(fa, fb), f"step {t}"


@pytest.mark.parametrize("shape", [(2, 32, 48, 4), (2, 64, 64, 10)])
def test_window_bptt_gradients_are_the_stepwise_gradients(shape):
    """Truncated BPTT over two windows through forward_window + EventWarping: saved activations bit-equal, loss / gradients to the atomics' noise floor."""
    from event_flow_b200.loss.flow import EventWarping

    B, H, W, T = shape
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
                outs = model.forward_window(torch.stack([d["event_voxel"] for d in part]).to(DEV),
                                            torch.stack([d["event_cnt"] for d in part]).to(DEV))
            else:
                outs = [model(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV)) for d in part]
            for d, out in zip(part, outs):
                lossf.event_flow_association(out["flow"], d["event_list"].clone().to(DEV), d["event_list_pol_mask"].to(DEV), d["event_mask"].to(DEV))
            loss = lossf()
            bank = model._arena.banks[model._arena.parity]  # what the backward will read: membrane potentials and spikes of every step
            acts = [v[:T].clone() for v in bank.v] + [z[1:T + 1].clone() for z in bank.zs] + [bank.flow[:T].clone(), bank.x_cl[:T].clone()]
            loss.backward()
            lossf.reset()
            model.detach_states()
            res.append((loss.item(), {n: p.grad.clone() for n, p in model.named_parameters()}, acts))
        return res

    ra, rb = run(True), run(False)
    for w, ((la, ga, aa), (lb, gb, ab)) in enumerate(zip(ra, rb)):
        for k, (x, y) in enumerate(zip(aa, ab)):
            assert torch.equal(x, y), f"window {w}: saved activation {k} differs"
        # same launches on bit-identical activations (test above); what differs run to run is the summation order of the atomics in the
        # loss kernels and the per-channel reductions (measured floor: profiles/r02_loss_gradient_noise_floor.txt, ~1e-6 .. 1e-5)
        assert abs(la - lb) <= 1e-6 * abs(lb), f"window {w}: loss {la} vs {lb}"
        for n in gb:
            scale = gb[n].abs().max().item() + 1e-30
            assert (ga[n] - gb[n]).abs().max().item() <= 5e-5 * scale, f"window {w}: gradient of {n} differs by {(ga[n] - gb[n]).abs().max().item():.3e}"


def test_window_entry_refuses_to_start_inside_a_window():
    B, H, W, T = 1, 32, 32, 2
    wins = _windows(B, 200, H, W, T + 1, 5, 130)
    m = _model()
    d = wins[0]
    m(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV))  # gradients enabled: a BPTT window is open
    with pytest.raises(RuntimeError):
        m.forward_window(torch.stack([w["event_voxel"] for w in wins[1:]]).to(DEV), torch.stack([w["event_cnt"] for w in wins[1:]]).to(DEV))


def test_window_c_abi_rejects_bad_arguments():
    from event_flow_b200 import _lib as L

    q = L.LifConvWindowParams()
    q.B, q.T, q.H, q.W = 1, 2, 16, 18  # W not a multiple of 4
    assert L.lib().ef_lif_conv_fwd_window(q, None) < 0
    assert L.lib().ef_lif_conv_fwd_window(None, None) < 0


def test_staged_training_loop_is_the_stepwise_training_loop():
    """event_flow_b200.train.train_windows(staged=True) vs the reference-shaped per-step loop: same window losses, same parameters."""
    from event_flow_b200.loss.flow import EventWarping
    from event_flow_b200.train import SyntheticEventStream, build_trainer, train_windows

    H, W, B, T, N = 32, 48, 2, 4, 300
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False, "clip_grad": 100.0},
           "model": {"mask_output": True}, "optimizer": {"lr": 2e-4}}

    def run(staged, n_windows):
        model = _model()
        trainer = build_trainer(model, cfg)
        loader = SyntheticEventStream(B, N, (H, W), 5, DEV, seq_len=3 * T + 2, n_items=8 * T)  # a new recording starts inside the 4th window
        losses = train_windows(model, EventWarping(cfg, DEV), trainer, loader, T * N, n_windows, staged=staged)
        return losses, trainer.flat_param.clone()

    (la, pa), (lb, pb) = run(True, 1), run(False, 1)
    # first optimiser step: same loss; Adam's first update is lr * sign(g) whatever |g| is, so the run-to-run noise of the atomics
    # (1e-6 of the gradient scale) may flip the update of a parameter whose gradient is ~0: almost all parameters agree to rounding,
    # none differs by more than 2 lr
    assert abs(la[0] - lb[0]) <= 1e-6 * abs(lb[0])
    d = (pa - pb).abs()
    assert d.max().item() <= 2.1 * 2e-4 and (d > 1e-6).float().mean().item() < 0.01, (d.max().item(), (d > 1e-6).float().mean().item())
    # five windows with a new recording inside the 4th: same number of optimiser steps, losses of the same trajectory family (spiking
    # networks amplify the differences above)
    (la, pa), (lb, pb) = run(True, 5), run(False, 5)
    assert len(la) == len(lb) == 5
    for a, b in zip(la, lb):
        assert abs(a - b) <= 0.05 * abs(b), (la, lb)


@pytest.mark.parametrize("T", [1, 13])
def test_window_lengths_one_step_and_longer_than_the_bank(T):
    """T = 1 (the fused launch degenerates to a step) and T = 13 (> the default bank capacity of 12: the arena grows)."""
    B, H, W = 2, 32, 48
    wins = _windows(B, 300, H, W, T, 5, 150)
    a, b = _model(), _model()
    vox = torch.stack([d["event_voxel"] for d in wins]).to(DEV)
    outs = a.forward_window(vox, None)
    for t, d in enumerate(wins):
        assert torch.equal(outs[t]["flow"][0], b(d["event_voxel"].to(DEV), None)["flow"][0]), f"step {t}"
    sum(o["flow"][0].square().sum() for o in outs).backward()
    b_loss = None
    b.zero_grad(set_to_none=True)
    b.reset_states()
    b_loss = sum(b(d["event_voxel"].to(DEV), None)["flow"][0].square().sum() for d in wins)
    b_loss.backward()
    for (n, pa), pb in zip(a.named_parameters(), b.parameters()):
        assert (pa.grad - pb.grad).abs().max().item() <= 2e-5 * (pb.grad.abs().max().item() + 1e-30), n


def test_window_starts_from_states_set_through_the_state_api_and_soft_reset():
    """model.states = ... (reference format, [2,B,C,H,W] per layer) before a window; soft-reset cells (hard_reset=False)."""
    from event_flow_b200.models.model import LIFFireNet

    B, H, W, T = 2, 32, 48, 3
    cfg = firenet_cfg(5, "voxel")
    cfg["spiking_neuron"] = dict(cfg["spiking_neuron"], hard_reset=False)
    wins = _windows(B, 400, H, W, 2 * T, 5, 170)

    def mk():
        torch.manual_seed(1)
        m = LIFFireNet(dict(cfg, spiking_neuron=dict(cfg["spiking_neuron"])))
        with torch.no_grad():
            for n, p in m.named_parameters():
                if n.endswith("ff.weight") or n.endswith("rec.weight"):
                    p.mul_(2.5)
            m.pred.conv2d.weight.mul_(20.0)
        return m.to(DEV)

    a, b, c = mk(), mk(), mk()
    with torch.no_grad():
        for d in wins[:T]:
            c(d["event_voxel"].to(DEV), None)
        states = c.states  # clones in the reference's format
        a.states = [s.clone() for s in states]
        b.states = [s.clone() for s in states]
        part = wins[T:]
        outs = a.forward_window(torch.stack([d["event_voxel"] for d in part]).to(DEV), None)
        for t, d in enumerate(part):
            ref = b(d["event_voxel"].to(DEV), None)["flow"][0]
            assert torch.equal(outs[t]["flow"][0], ref), f"step {t}"
            assert torch.equal(ref, c(d["event_voxel"].to(DEV), None)["flow"][0])
        for sa, sb in zip(a.states, b.states):
            assert torch.equal(sa, sb)
