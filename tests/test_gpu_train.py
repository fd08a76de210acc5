"""GPU: fused clip + Adam against torch's own clip_grad_norm_ + Adam; 1-GPU DP trainer step; bench-sized forward sanity."""
import pytest
import torch

from oracle import encodings as oenc
from tests.util import firenet_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("clip", [100.0, 0.5])
def test_clip_adam_matches_torch(clip):
    from event_flow_b200.parallel import DataParallelTrainer

    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Conv2d(8, 4, 1)).to(DEV)
    ref = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Conv2d(8, 4, 1)).to(DEV)
    ref.load_state_dict(lin.state_dict())
    tr = DataParallelTrainer(lin, lr=2e-4, clip_grad=clip)
    opt = torch.optim.Adam(ref.parameters(), lr=2e-4)
    for it in range(5):
        x = torch.randn(4, 3, 12, 12, device=DEV)
        (lin(x) ** 2).sum().backward()
        (ref(x) ** 2).sum().backward()
        total = torch.nn.utils.clip_grad_norm_(ref.parameters(), clip)
        tr.step()
        torch.testing.assert_close(tr.grad_norm(), total, rtol=1e-5, atol=0)
        opt.step()
        opt.zero_grad()
        for a, b in zip(lin.parameters(), ref.parameters()):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7)
            assert a.grad.abs().max() == 0


def test_trainer_step_on_firenet_changes_parameters_and_keeps_grad_views():
    from event_flow_b200.loss.flow import EventWarping
    from event_flow_b200.models.model import LIFFireNet
    from event_flow_b200.parallel import DataParallelTrainer

    Hh = Ww = 64
    torch.manual_seed(0)
    m = LIFFireNet(firenet_cfg(5, "voxel"))
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        m.pred.conv2d.weight.mul_(50.0)
    m = m.to(DEV)
    tr = DataParallelTrainer(m)
    cfg = {"loader": {"resolution": [Hh, Ww]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False}, "model": {"mask_output": True}}
    L = EventWarping(cfg, DEV)
    before = tr.flat_param.clone()
    for t in range(3):
        d = oenc.encode_window(*oenc.synthetic_events(2, 500, Hh, Ww, 40 + t), Hh, Ww, 5)
        out = m(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV))
        L.event_flow_association(out["flow"], d["event_list"].to(DEV), d["event_list_pol_mask"].to(DEV), d["event_mask"].to(DEV))
    L().backward()
    assert tr.flat_grad.abs().max() > 0
    tr.step()
    assert (tr.flat_param - before).abs().max() > 0 and tr.flat_grad.abs().max() == 0
    assert torch.isfinite(tr.flat_param).all()


def test_cached_backward_arguments_reproduce_the_first_window():
    """
    The fast path builds the argument structs of a step's backward calls once per (arena slot, sweep position) and replays
    them afterwards.  Windows 3 and 4 (same arena banks as windows 1 and 2) must give the gradients of the uncached windows.
    """
    from event_flow_b200.loss.flow import EventWarping
    from event_flow_b200.models.model import LIFFireNet
    from oracle import encodings as oenc
    from tests.util import firenet_cfg

    torch.manual_seed(0)
    H, W, B, T = 32, 48, 2, 3
    model = LIFFireNet(firenet_cfg(5, "voxel"))
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        model.pred.conv2d.weight.mul_(20.0)
    model = model.to(DEV)
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False}, "model": {"mask_output": True}}
    lossf = EventWarping(cfg, DEV)
    wins = [oenc.encode_window(*oenc.synthetic_events(B, 300, H, W, 50 + t), H, W, 5) for t in range(T)]

    def window():
        model.zero_grad(set_to_none=True)
        model.reset_states()
        lossf.reset()
        for d in wins:
            out = model(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV))
            lossf.event_flow_association(out["flow"], d["event_list"].clone().to(DEV), d["event_list_pol_mask"].to(DEV), d["event_mask"].to(DEV))
        loss = lossf()
        loss.backward()
        model.detach_states()
        return loss.item(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    ref_loss, ref = window()
    for k in range(4):
        loss, grads = window()
        assert loss == ref_loss
        for n, g in grads.items():
            scale = ref[n].abs().max().item() + 1e-20
            assert (g - ref[n]).abs().max().item() <= 1e-5 * scale, f"window {k + 2}: {n}"


def test_weight_images_of_the_cell_path_follow_the_fused_optimizer():
    """
    PLIF FireNet (cells on the tensor-core convolution + neuron kernel; weight images cached per weight tensor): after a fused
    clip + Adam step -- which writes the parameters behind torch's version counters -- the next forward must use the NEW weights.
    """
    import copy

    from event_flow_b200.models.model import PLIFFireNet
    from event_flow_b200.parallel import DataParallelTrainer

    H = W = 32
    torch.manual_seed(0)
    cfg = firenet_cfg(5, "voxel", neuron="plif")
    m = PLIFFireNet(cfg)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        m.pred.conv2d.weight.mul_(50.0)
    m = m.to(DEV)
    tr = DataParallelTrainer(m, lr=1e-2)  # a large step: stale weights would be obvious
    d = oenc.encode_window(*oenc.synthetic_events(2, 400, H, W, 3), H, W, 5)
    vox = d["event_voxel"].to(DEV)
    m(vox, None)["flow"][0].square().sum().backward()
    tr.step()
    m.reset_states()
    with torch.no_grad():
        after = m(vox, None)["flow"][0].clone()
    twin = copy.deepcopy(m)  # fresh run-time state, same (updated) parameters
    twin.reset_states()
    for cell in twin.modules():
        cell.__dict__.pop("_x_kind", None)  # the twin runs the fused CUDA-core kernel: no weight images at all
    with torch.no_grad():
        ref = twin(vox, None)["flow"][0]
    assert ref.abs().max() > 0
    assert (after - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


@pytest.mark.parametrize("windowed", [True, False])
def test_gradient_sink_equals_autograd_accumulation(windowed):
    """
    With a DataParallelTrainer the window backward accumulates the parameter gradients straight into the trainer's flat buffer (no clone
    + add per parameter): the buffer must hold what autograd's own accumulation gives for the same window on a twin model, also when
    TWO windows are accumulated before the optimiser step (p.grad += semantics).
    """
    import copy

    from event_flow_b200.loss.flow import EventWarping
    from event_flow_b200.models.model import LIFFireNet
    from event_flow_b200.parallel import DataParallelTrainer

    H, W, B, T = 32, 48, 2, 3
    torch.manual_seed(0)
    a = LIFFireNet(firenet_cfg(5, "voxel"))
    with torch.no_grad():
        for n, p in a.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(2.5)
        a.pred.conv2d.weight.mul_(20.0)
    a = a.to(DEV)
    b = copy.deepcopy(a)
    tr = DataParallelTrainer(a)
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001, "overwrite_intermediate": False}, "model": {"mask_output": True}}
    wins = [oenc.encode_window(*oenc.synthetic_events(B, 300, H, W, 60 + t), H, W, 5) for t in range(2 * T)]

    def run(m, sink_expected):
        lossf = EventWarping(cfg, DEV)
        for w in range(2):
            part = wins[w * T:(w + 1) * T]
            if windowed:
                outs = m.forward_window(torch.stack([d["event_voxel"] for d in part]).to(DEV), torch.stack([d["event_cnt"] for d in part]).to(DEV))
            else:
                outs = [m(d["event_voxel"].to(DEV), d["event_cnt"].to(DEV)) for d in part]
            for d, out in zip(part, outs):
                lossf.event_flow_association(out["flow"], d["event_list"].clone().to(DEV), d["event_list_pol_mask"].to(DEV), d["event_mask"].to(DEV))
            lossf().backward()
            assert m._fast.carry.into_sink is sink_expected
            lossf.reset()
            m.detach_states()  # no optimiser step in between: the second window ACCUMULATES

    run(a, True)
    run(b, False)
    ref = torch.cat([p.grad.reshape(-1) for p in b.parameters()])
    assert ref.abs().max() > 0
    assert (tr.flat_grad - ref).abs().max().item() <= 5e-5 * ref.abs().max().item()
    for p in a.parameters():  # still views of the flat buffer
        assert p.grad.data_ptr() >= tr.flat_grad.data_ptr() and p.grad.data_ptr() < tr.flat_grad.data_ptr() + 4 * tr.flat_grad.numel()
