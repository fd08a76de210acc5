"""
The no-grad step of the cell-by-cell models replayed as a CUDA graph (event_flow_b200/graphed.py) against the same model stepping
launch by launch: same flows and states over sequences with resets, caller-set states, interleaved eager steps (grad mode), weight
updates, hooks; and the graphs stay out of checkpoints.
"""
import copy
import pickle

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

FIRE = dict(name="x", encoding="voxel", round_encoding=False, norm_input=False, num_bins=5, base_num_channels=32, kernel_size=3,
            activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron={})
ANN = dict(name="x", encoding="voxel", round_encoding=False, norm_input=False, num_bins=5, base_num_channels=32, kernel_size=3,
           activations=["relu", None], mask_output=True, spiking_neuron=None)
UNET = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=8, kernel_size=3,
            activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron={})
UNET_ANN = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=8, kernel_size=3,
                activations=["relu", None], mask_output=True, spiking_neuron=None)
MODELS = [("FireNet", ANN), ("FireFlowNet", ANN), ("RNNFireNet", ANN), ("LeakyFireNet", ANN), ("PLIFFireNet", FIRE), ("ALIFFireNet", FIRE),
          ("XLIFFireNet", FIRE), ("EVFlowNet", UNET_ANN), ("RecEVFlowNet", UNET_ANN), ("E2VID", UNET_ANN), ("PLIFRecEVFlowNet", UNET)]


def build_pair(cls, cfg, gain=2.0):
    import event_flow_b200.models.model as M

    torch.manual_seed(3)
    a = getattr(M, cls)(dict(cfg))
    with torch.no_grad():
        for n, p in a.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(gain)
    a = a.to(DEV).eval()
    b = copy.deepcopy(a)
    b.__dict__["_graph_off"] = True  # the launch-by-launch twin
    return a, b


def inputs(cfg, B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    vox = (torch.rand(B, cfg["num_bins"], H, W, generator=g) * 4 - 2) * (torch.rand(B, cfg["num_bins"], H, W, generator=g) < 0.3)
    cnt = torch.randint(0, 3, (B, 2, H, W), generator=g).float()
    return vox.to(DEV), cnt.to(DEV)


def leaves(states):
    from event_flow_b200.graphed import _flat

    return _flat(states, [])


def same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x.shape == y.shape and torch.equal(x, y)


def n_graphs(model):
    from event_flow_b200.graphed import _StepGraph

    return sum(isinstance(g, _StepGraph) for g in model.__dict__.get("_step_graphs", {}).values())


@pytest.mark.parametrize("cls,cfg", MODELS, ids=[m[0] for m in MODELS])
def test_graphed_steps_equal_launch_by_launch_steps(cls, cfg):
    a, b = build_pair(cls, cfg)
    B, H, W = 2, 32, 32
    kept = []
    with torch.no_grad():
        for k in range(12):
            if k == 7:  # a new sequence: the first-step graph and the later-step graph are both reused
                a.reset_states(), b.reset_states()
            vox, cnt = inputs(cfg, B, H, W, k)
            fa, fb = a(vox, cnt)["flow"], b(vox, cnt)["flow"]
            same(fa, fb)
            kept.append((fa, [f.clone() for f in fa]))
            if hasattr(a, "states"):
                same(leaves(a.states), leaves(b.states))
    assert not a.__dict__.get("_graph_off"), a.__dict__.get("_graph_error")
    assert n_graphs(a) >= 1 and n_graphs(b) == 0
    for live, copy_ in kept:  # outputs handed out earlier are not overwritten by later replays
        same(live, copy_)


def test_caller_set_states_and_interleaved_grad_steps():
    cls, cfg = "ALIFFireNet", FIRE
    a, b = build_pair(cls, cfg)
    B, H, W = 2, 32, 32
    with torch.no_grad():
        for k in range(5):
            vox, cnt = inputs(cfg, B, H, W, k)
            same(a(vox, cnt)["flow"], b(vox, cnt)["flow"])
    assert n_graphs(a) >= 1
    # states handed in by the caller (the reference's state API)
    st = b.states
    a.states = copy.deepcopy(st)
    with torch.no_grad():
        vox, cnt = inputs(cfg, B, H, W, 50)
        same(a(vox, cnt)["flow"], b(vox, cnt)["flow"])
    # a training step in between: its saved tensors must survive the replays that follow, and its gradients equal the twin's
    a.train(), b.train()
    vox, cnt = inputs(cfg, B, H, W, 60)
    la = a(vox, cnt)["flow"][0].square().sum()
    lb = b(vox, cnt)["flow"][0].square().sum()
    a.eval(), b.eval()
    with torch.no_grad():
        sa, sb = a.states, b.states
        for k in range(3):
            v2, c2 = inputs(cfg, B, H, W, 70 + k)
            same(a(v2, c2)["flow"], b(v2, c2)["flow"])
        a.states, b.states = sa, sb
    la.backward(), lb.backward()
    for (n, p), q in zip(a.named_parameters(), b.parameters()):
        assert (p.grad is None) == (q.grad is None), n
        if p.grad is not None:
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-7), n


def test_weight_update_is_a_new_situation_and_hooks_disable_replay():
    cls, cfg = "FireNet", ANN
    a, b = build_pair(cls, cfg)
    B, H, W = 1, 32, 32
    with torch.no_grad():
        for k in range(4):
            vox, cnt = inputs(cfg, B, H, W, k)
            same(a(vox, cnt)["flow"], b(vox, cnt)["flow"])
        for m in (a, b):
            for p in m.parameters():
                p.mul_(1.25)
        for k in range(4):
            vox, cnt = inputs(cfg, B, H, W, 10 + k)
            same(a(vox, cnt)["flow"], b(vox, cnt)["flow"])
        calls = []
        h = a.G1.register_forward_hook(lambda m, i, o: calls.append(1))
        for k in range(3):
            vox, cnt = inputs(cfg, B, H, W, 20 + k)
            same(a(vox, cnt)["flow"], b(vox, cnt)["flow"])
        h.remove()
    assert len(calls) == 3 and n_graphs(a) >= 1


def test_models_with_graphs_pickle_and_deepcopy():
    cls, cfg = "PLIFFireNet", FIRE
    a, b = build_pair(cls, cfg)
    with torch.no_grad():
        for k in range(4):
            vox, cnt = inputs(cfg, 2, 32, 32, k)
            a(vox, cnt), b(vox, cnt)
    assert n_graphs(a) >= 1
    c = pickle.loads(pickle.dumps(a))
    d = copy.deepcopy(a)
    assert "_step_graphs" not in c.__dict__ and "_step_graphs" not in d.__dict__
    with torch.no_grad():
        vox, cnt = inputs(cfg, 2, 32, 32, 9)
        ref = b(vox, cnt)["flow"]
        same(c(vox, cnt)["flow"], ref)
        same(d(vox, cnt)["flow"], ref)
        same(a(vox, cnt)["flow"], ref)


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("c1,c2,cout", [(5, 0, 32), (32, 32, 64), (24, 8, 16)])
def test_latency_shape_of_the_ann_convolution_matches_torch(stride, c1, c2, cout, monkeypatch):
    """Small inference launches of ef_conv_ann_fwd (8 output channels per CTA, input channels split over 4 thread groups) vs F.conv2d."""
    import torch.nn.functional as F

    from event_flow_b200 import ops

    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    g = torch.Generator().manual_seed(c1 + 7 * stride)
    B, H, W = 1, 45, 70  # ragged against the 16x16 tiles
    x1 = torch.randn(B, c1, H, W, generator=g).to(DEV)
    x2 = torch.randn(B, c2, H, W, generator=g).to(DEV) if c2 else None
    sc = torch.rand(B, c2, H, W, generator=g).to(DEV) if c2 else None
    w = (torch.randn(cout, c1 + c2, 3, 3, generator=g) * 0.1).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    res = torch.randn(B, cout, Ho, Wo, generator=g).to(DEV)
    h, u = torch.randn(B, cout, Ho, Wo, generator=g).to(DEV), torch.rand(B, cout, Ho, Wo, generator=g).to(DEV)
    with torch.no_grad():
        got = ops.conv_ann(x1, w, b, "relu", x2=x2, x2_scale=sc, residual=res, blend_h=h, blend_u=u, stride=stride)
    xin = x1 if x2 is None else torch.cat([x1, x2 * sc], 1)
    o = torch.relu(F.conv2d(xin.double(), w.double(), b.double(), stride=stride, padding=1) + res.double())
    want = (h.double() * (1 - u.double()) + o * u.double()).float()
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    # the tracked path (sequential channel order) agrees to rounding
    w2 = w.clone().requires_grad_(True)
    tracked = ops.conv_ann(x1, w2, b, "relu", x2=x2, x2_scale=sc, residual=res, blend_h=h, blend_u=u, stride=stride)
    torch.testing.assert_close(got, tracked.detach(), rtol=1e-5, atol=1e-5)
