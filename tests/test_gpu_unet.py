"""
GPU parity of the spiking recurrent EV-FlowNet family (BASELINE config 4 is this model at 256x256):
per-cell teacher-forced comparison with the CPU oracle on the reference's golden weights, the resampling kernels against
torch, the state API, and a larger-shape run (256x256, base 32 channels) checked the same way.
"""
import glob
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import encodings as oenc
from oracle import spiking as osp
from oracle import unet as ounet
from tests.conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"
UNETS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "unet_*.npz")))
CLASSES = {"lif": "SpikingRecEVFlowNet", "plif": "PLIFRecEVFlowNet", "alif": "ALIFRecEVFlowNet", "xlif": "XLIFRecEVFlowNet"}


def unet_cfg(neuron, base=4):
    return dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=base, kernel_size=3,
                activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron=None if neuron == "lif" else {})


def hook_cells(model, tc=False):
    """
    (name, x_in, state_in, out, state_out) of every spiking cell, in execution order, as CPU tensors.  tc=False pins the cell-by-cell
    path (forward hooks on the cells); tc=True records the same tuples from the tensor-core inference path (fast_unet._capture).
    """
    from event_flow_b200.models.spiking_submodules import _SpikingConvCell

    trace, handles = [], []
    net = getattr(model, "net", None)
    if net is not None and hasattr(net, "num_encoders"):
        net.__dict__["_use_tc"] = tc
        if tc:
            net.__dict__["_capture"], net.__dict__["_capture_prefix"] = trace, model.net_attr + "."
            return trace, handles
    for name, mod in model.named_modules():
        if isinstance(mod, _SpikingConvCell):
            def hook(m, inputs, kwargs, output, name=name):
                st = inputs[1] if len(inputs) > 1 else None
                res = kwargs.get("residual", inputs[2] if len(inputs) > 2 else 0)
                trace.append((name, inputs[0].detach().cpu(), None if st is None else st.detach().cpu(),
                              res.detach().cpu() if torch.is_tensor(res) else 0, output[0].detach().cpu(), output[1].detach().cpu(), m.stride))
            handles.append(mod.register_forward_hook(hook, with_kwargs=True))
    return trace, handles


def check_trace(neuron, sd, trace):
    worst, flips_in, n_tot = 0.0, 0, 0
    for name, x_in, st_in, res, out, st_out, stride in trace:
        p = ounet.cell_params(sd, name + ".", neuron)
        out_o, st_o = osp.cell_step(neuron, x_in, st_in, p, stride=stride, residual=res)
        if neuron in ("lif", "plif"):
            thr = p["thresh"].clamp_min(0.01)
        else:
            thr = p["t0"].clamp_min(0.01) + p["t1"].clamp_min(0) * st_o[2]
        # fp32 summation-order noise, magnitude scaled (tests/util.py: 3e-6 * max|v| was measured for sums of <= 576 terms);
        # it grows with the square root of the number of accumulated terms (the 1024 -> 256 decoder adds 9216 per output)
        k_terms = 9 * (p["ff"].shape[1] + (p["rec"].shape[1] if "rec" in p else 0))
        tol = max(2e-5, 3e-6 * st_o[0].abs().max().item() * max(1.0, (k_terms / 576.0) ** 0.5))
        dv = (st_out[0] - st_o[0]).abs().max().item()
        assert dv <= tol, f"{name}: max|dv| {dv:.2e} > {tol:.2e}"
        near = (st_o[0] - thr).abs() < tol
        diff = st_out[1] != st_o[1]
        assert int((diff & ~near).sum()) == 0, f"{name}: spike flips outside the tolerance band"
        assert torch.equal(out[~diff], out_o[~diff]), f"{name}: output (spikes + residual)"
        if st_o.shape[0] == 3:
            assert (st_out[2] - st_o[2]).abs().max().item() <= 1e-5, f"{name}: trace state"
        worst, flips_in, n_tot = max(worst, dv / tol), flips_in + int((diff & near).sum()), n_tot + diff.numel()
    return worst, flips_in, n_tot


@pytest.mark.parametrize("name", UNETS)
def test_unet_cells_match_oracle_on_reference_weights(name):
    import event_flow_b200.models.model as M

    g = load_golden(name)
    neuron = name.split("_")[1]
    m = getattr(M, CLASSES[neuron])(unet_cfg(neuron))
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd_")}
    m.load_state_dict(sd)  # checkpoint compatibility: the reference's state_dict loads unchanged
    m = m.to(DEV)
    trace, handles = hook_cells(m)
    T = len([k for k in g if k.startswith("x_")])
    with torch.no_grad():
        for t in range(T):
            out = m(None, g["x_%d" % t].to(DEV))
    assert len(trace) == 16 * T and len(out["flow"]) == 4
    worst, flips_in, n_tot = check_trace(neuron, sd, trace)
    # the last step's flows: every scale upsampled to the input resolution, same shapes as the reference's
    for i in range(4):
        assert out["flow"][i].shape == g["flow_%d_%d" % (T - 1, i)].shape
    states = m.states
    assert len(states) == 10 and states[0].shape == g["state_0"].shape and states[9].shape == g["state_9"].shape
    print(f"{name}: worst |dv|/tol {worst:.2f}, {flips_in}/{n_tot} borderline flips inside the band")


def test_resampling_kernels_match_torch():
    from event_flow_b200 import ops

    g = torch.Generator().manual_seed(3)
    for shape in ((2, 5, 7, 9), (1, 3, 1, 1), (2, 8, 16, 24)):
        x = torch.randn(shape, generator=g)
        up = ops.upsample_bilinear2x(x.to(DEV)).cpu()
        ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        assert (up - ref).abs().max().item() <= 2e-7 * ref.abs().max().item()
        z = (torch.rand(shape, generator=g) < 0.4).float()  # spikes: products with 0.25 / 0.75 are exact
        assert torch.equal(ops.upsample_bilinear2x(z.to(DEV)).cpu(), F.interpolate(z, scale_factor=2, mode="bilinear", align_corners=False))
        for f in (1, 2, 4, 8):
            assert torch.equal(ops.upsample_nearest(x.to(DEV), f, f).cpu(), F.interpolate(x, scale_factor=(float(f), float(f))))


def test_unet_state_api_and_cropping():
    import event_flow_b200.models.model as M

    torch.manual_seed(0)
    m = M.SpikingRecEVFlowNet(unet_cfg("lif")).to(DEV)
    m.init_cropping(40, 24)  # 24 x 40 is padded to 32 x 48 and cropped back
    ts, ys, xs, ps = oenc.synthetic_events(1, 600, 24, 40, 5)
    cnt = oenc.encode_window(ts, ys, xs, ps, 24, 40, 2)["event_cnt"].to(DEV)
    with torch.no_grad():
        out = m(None, cnt)
    assert [tuple(f.shape) for f in out["flow"]] == [(1, 2, 24, 40)] * 4
    st = m.states
    assert len(st) == 10 and tuple(st[0].shape) == (2, 2, 1, 8, 16, 24) and tuple(st[9].shape) == (2, 1, 4, 32, 48)
    m.detach_states()
    m.states = st
    m.reset_states()
    assert all(s is None for s in m.states)
    with pytest.raises(AttributeError):  # bad encoding: same exception as models/model.py:497-499
        M.SpikingRecEVFlowNet(dict(unet_cfg("lif"), encoding="nope")).to(DEV)(cnt, cnt)


@pytest.mark.parametrize("neuron", ["lif", "plif", "alif", "xlif"])
@pytest.mark.parametrize("rec", [False, True])
def test_stride2_and_wide_cell_backward_matches_oracle_autograd(neuron, rec):
    """Cell-level gradients for the U-Net shapes: stride 2 (encoders), many channels, residual input (resblocks)."""
    import event_flow_b200.models.spiking_submodules as S

    torch.manual_seed(7)
    stride = 1 if rec else 2
    cin, c, H, W = (12, 12, 10, 14) if rec else (6, 20, 13, 18)
    cls = {("lif", False): S.ConvLIF, ("plif", False): S.ConvPLIF, ("alif", False): S.ConvALIF, ("xlif", False): S.ConvXLIF,
           ("lif", True): S.ConvLIFRecurrent, ("plif", True): S.ConvPLIFRecurrent, ("alif", True): S.ConvALIFRecurrent,
           ("xlif", True): S.ConvXLIFRecurrent}[(neuron, rec)]
    cell = cls(cin, c, 3) if rec else cls(cin, c, 3, stride)
    with torch.no_grad():
        cell.ff.weight.mul_(2.0)
    g = torch.Generator().manual_seed(5)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    x = ((torch.rand((2, cin, H, W), generator=g) < 0.4).float() * (torch.rand((2, cin, H, W), generator=g) + 0.5)).requires_grad_(True)
    n_state = 2 if neuron == "lif" else 3
    st = torch.rand((n_state, 2, c, Ho, Wo), generator=g)
    st[1] = (st[1] < 0.3).float()
    st.requires_grad_(True)
    res = torch.rand((2, c, Ho, Wo), generator=g).requires_grad_(True)
    gw_out, gw_st = torch.randn((2, c, Ho, Wo), generator=g), torch.randn((n_state, 2, c, Ho, Wo), generator=g) * 0.3
    # oracle (CPU autograd)
    p = {"ff": cell.ff.weight.detach().clone().requires_grad_(True)}
    if rec:
        p["rec"] = cell.rec.weight.detach().clone().requires_grad_(True)
    for n in ounet.CELL_PARAM_NAMES[neuron]:
        p[n] = getattr(cell, n).detach().clone().requires_grad_(True)
    out_o, st_o = osp.cell_step(neuron, x, st, p, stride=stride, residual=0 if rec else res)  # recurrent cells take no residual
    ((out_o * gw_out).sum() + (st_o * gw_st).sum()).backward()
    # CUDA path
    cell = cell.to(DEV)
    xg, stg, resg = (t.detach().to(DEV).requires_grad_(True) for t in (x, st, res))
    out, st_n = cell(xg, stg) if rec else cell(xg, stg, residual=resg)
    ((out * gw_out.to(DEV)).sum() + (st_n * gw_st.to(DEV)).sum()).backward()
    assert not (st_n[1].cpu() != st_o[1]).any(), "spike flip in the gradient test case (pick another seed)"

    def chk(mine, ref, what):
        scale = ref.abs().max().item() + 1e-12
        err = (mine.cpu() - ref).abs().max().item()
        assert err <= 1e-3 * scale, f"{what}: {err:.3e} vs scale {scale:.3e}"

    chk(xg.grad, x.grad, "g_x")
    chk(stg.grad, st.grad, "g_state")
    if not rec:
        chk(resg.grad, res.grad, "g_residual")
    chk(cell.ff.weight.grad, p["ff"].grad, "g_w_ff")
    if rec:
        chk(cell.rec.weight.grad, p["rec"].grad, "g_w_rec")
    for n in ounet.CELL_PARAM_NAMES[neuron]:
        if getattr(cell, n).requires_grad:
            chk(getattr(cell, n).grad, p[n].grad, "g_" + n)


def test_resampling_backward_matches_torch_autograd():
    from event_flow_b200 import ops

    g = torch.Generator().manual_seed(4)
    for shape in ((2, 3, 5, 7), (1, 2, 1, 1), (1, 4, 8, 2)):
        x = torch.randn(shape, generator=g, requires_grad=True)
        w = torch.randn((shape[0], shape[1], 2 * shape[2], 2 * shape[3]), generator=g)
        (F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False) * w).sum().backward()
        xg = x.detach().to(DEV).requires_grad_(True)
        (ops.upsample_bilinear2x(xg) * w.to(DEV)).sum().backward()
        assert (xg.grad.cpu() - x.grad).abs().max().item() <= 1e-6 * x.grad.abs().max().item()
        for f in (2, 4):
            x2 = torch.randn(shape, generator=g, requires_grad=True)
            w2 = torch.randn((shape[0], shape[1], f * shape[2], f * shape[3]), generator=g)
            (F.interpolate(x2, scale_factor=(float(f), float(f))) * w2).sum().backward()
            xg2 = x2.detach().to(DEV).requires_grad_(True)
            (ops.upsample_nearest(xg2, f, f) * w2.to(DEV)).sum().backward()
            assert (xg2.grad.cpu() - x2.grad).abs().max().item() <= 1e-5 * x2.grad.abs().max().item()


@pytest.mark.parametrize("name", UNETS)
def test_unet_bptt_gradients_match_reference_golden(name):
    """Training through the U-Net: BPTT gradients of a linear functional of all flow scales of 3 steps vs the reference's autograd."""
    import event_flow_b200.models.model as M

    g = load_golden(name)
    neuron = name.split("_")[1]
    m = getattr(M, CLASSES[neuron])(unet_cfg(neuron))
    m.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("sd_")})
    m = m.to(DEV)
    T = len([k for k in g if k.startswith("x_")])
    loss = 0.0
    for t in range(T):
        out = m(None, g["x_%d" % t].to(DEV))
        loss = loss + sum((f * g["gw_%d_%d" % (t, i)].to(DEV)).sum() for i, f in enumerate(out["flow"]))
    loss.backward()
    m.detach_states()
    # threshold chaos: a single borderline spike flip changes the trajectory, and with it the gradients; the
    # comparison is only meaningful on an identical spike trajectory, which the final states witness
    states = m.states
    spikes = lambda t: t[:, 1] if t.dim() == 6 else t[1]  # noqa: E731  (encoders / resblocks stack two cell states)
    same = all(torch.equal(spikes(states[i]).cpu(), spikes(g["state_%d" % i])) for i in (0, 3, 5, 9))
    if not same:
        pytest.skip("a borderline spike flipped on this box: trajectories differ, gradients are not comparable")
    assert abs(loss.item() - g["loss"].item()) <= 1e-4 * abs(g["loss"].item()) + 1e-5
    checked = 0
    for nm, q in m.named_parameters():
        if "grad_" + nm in g:
            ref = g["grad_" + nm]
            scale = ref.abs().max().item() + 1e-12
            err = (q.grad.cpu() - ref).abs().max().item()
            assert err <= 1e-3 * scale, f"{nm}: {err:.3e} vs scale {scale:.3e}"
            checked += 1
    assert checked >= 20


@pytest.mark.parametrize("tc", [True, False])
def test_unet_full_size_cells_match_oracle(tc):
    """
    BASELINE config 4 shape per GPU reduced to B=1: 256x256, base 32 channels (2->64->128->256->512), two steps; every cell of every
    step teacher-forced against the oracle.  tc=True: the tensor-core inference path (general tcgen05 cell kernel, multi-source
    decoders, cl upsampling); tc=False: the fp32 cell-by-cell path (what training uses).
    """
    import event_flow_b200.models.model as M

    torch.manual_seed(3)
    m = M.SpikingRecEVFlowNet(unet_cfg("lif", base=32))
    with torch.no_grad():
        for nm, q in m.named_parameters():
            if nm.endswith("ff.weight") or nm.endswith("rec.weight"):
                q.mul_(3.0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    trace, handles = hook_cells(m, tc=tc)
    with torch.no_grad():
        for t in range(2):
            ts, ys, xs, ps = oenc.synthetic_events(1, 50000, 256, 256, 40 + t)
            out = m(None, oenc.encode_window(ts, ys, xs, ps, 256, 256, 2)["event_cnt"].to(DEV))
    assert [tuple(f.shape) for f in out["flow"]] == [(1, 2, 256, 256)] * 4
    assert len(trace) == 16 * 2, "every cell of both steps must have been recorded"
    worst, flips_in, n_tot = check_trace("lif", sd, trace)
    # the last decoder's spikes through the prediction layer = the finest flow map
    last = trace[-1]
    flow_o = osp.pred_head(last[4], sd[m.net_attr + ".preds.3.conv2d.weight"], sd[m.net_attr + ".preds.3.conv2d.bias"])
    f64 = torch.tanh(F.conv2d(last[4].double(), sd[m.net_attr + ".preds.3.conv2d.weight"].double(), sd[m.net_attr + ".preds.3.conv2d.bias"].double()))
    print("flow vs fp64:", (out["flow"][3].cpu() - f64).abs().max().item(), "oracle (CPU fp32) vs fp64:", (flow_o - f64).abs().max().item())
    torch.testing.assert_close(out["flow"][3].cpu().double(), f64, rtol=1e-5, atol=1e-6)
    # state API: the reference's layout whichever path ran
    st = m.states
    assert len(st) == 10 and st[0].shape == (2, 2, 1, 64, 128, 128) and st[9].shape == (2, 1, 32, 256, 256)
    assert torch.equal(st[9][1].cpu(), last[4]) and torch.equal(st[9].cpu(), last[5])
    print(f"full size (tc={tc}): worst |dv|/tol {worst:.2f}, {flips_in}/{n_tot} borderline flips")


def test_unet_tc_path_continues_the_cell_path_rollout():
    """Switching paths mid-sequence (states converted at the API boundary): step 2 on the tensor-core path from the states the cell path
    left equals, cell by cell within the T1 band, step 2 on the cell path."""
    import event_flow_b200.models.model as M

    torch.manual_seed(4)
    m = M.SpikingRecEVFlowNet(unet_cfg("lif", base=32))
    with torch.no_grad():
        for nm, q in m.named_parameters():
            if nm.endswith("ff.weight") or nm.endswith("rec.weight"):
                q.mul_(3.0)
    m = m.to(DEV)
    xs = [oenc.encode_window(*oenc.synthetic_events(1, 20000, 128, 128, 60 + t), 128, 128, 2)["event_cnt"].to(DEV) for t in range(2)]
    net = m.net
    with torch.no_grad():
        net.__dict__["_use_tc"] = False
        m(None, xs[0])
        saved = m.states
        a = m(None, xs[1])["flow"]
        m.states = saved
        net.__dict__["_use_tc"] = True
        b = m(None, xs[1])["flow"]
    for fa, fb in zip(a, b):
        assert (fa - fb).abs().max().item() <= 1e-3 * max(1e-6, fa.abs().max().item())


def test_ann_evflownet_forward_matches_reference_golden_and_oracle():
    """EV-FlowNet (ANN twin): flows vs the reference's golden output, and every 3x3 layer teacher-forced against the oracle."""
    import event_flow_b200.models.model as M

    g = load_golden("annunet_evflownet")
    cfg = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=4, kernel_size=3,
               activations=["relu", None], mask_output=True, spiking_neuron=None)
    m = M.EVFlowNet(cfg)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd_")}
    m.load_state_dict(sd)
    m = m.to(DEV)
    with torch.no_grad():
        out = m(None, g["x"].to(DEV))
    assert len(out["flow"]) == 4
    for i in range(4):
        ref = g["flow_%d" % i]
        assert out["flow"][i].shape == ref.shape
        err = (out["flow"][i].cpu() - ref).abs().max().item()
        assert err <= 1e-3 * ref.abs().max().item() + 1e-6, f"flow scale {i}: {err:.3e}"  # north-star tolerance 1e-3 rel; measured ~1e-6
    # training: gradients of a linear functional of all four flow scales vs the reference's autograd
    loss = sum((f * g["gw_%d" % i].to(DEV)).sum() for i, f in enumerate(m(None, g["x"].to(DEV))["flow"]))
    loss.backward()
    checked = 0
    for nm, q in m.named_parameters():
        if "grad_" + nm in g:
            ref = g["grad_" + nm]
            scale = ref.abs().max().item() + 1e-12
            assert (q.grad.cpu() - ref).abs().max().item() <= 1e-3 * scale, nm
            checked += 1
    assert checked >= 20
    m.reset_states(), m.detach_states()
    m.init_cropping(40, 24)
    with torch.no_grad():
        o2 = m(None, g["x"][:, :, :24, :40].contiguous().to(DEV))
    assert [tuple(f.shape) for f in o2["flow"]] == [(1, 2, 24, 40)] * 4


def test_ann_recevflownet_rollout_and_gradients_match_reference_golden():
    """RecEVFlowNet (stride-2 conv + ConvGRU encoders): 3-step rollout, final flows / states and BPTT gradients vs the reference."""
    import event_flow_b200.models.model as M

    g = load_golden("annunet_recevflownet")
    cfg = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=4, kernel_size=3,
               activations=["relu", None], mask_output=True, spiking_neuron=None)
    m = M.RecEVFlowNet(cfg)
    m.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("sd_")})
    m = m.to(DEV)
    T = len([k for k in g if k.startswith("x_")])
    with torch.no_grad():
        for t in range(T):
            out = m(None, g["x_%d" % t].to(DEV))
    for i in range(4):
        ref = g["flow_%d" % i]
        assert (out["flow"][i].cpu() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item() + 1e-6
        ref = g["state_%d" % i]
        assert (m.states[i].cpu() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item() + 1e-6
    m.reset_states()
    loss = 0.0
    for t in range(T):
        loss = loss + sum((f * g["gw_%d_%d" % (t, i)].to(DEV)).sum() for i, f in enumerate(m(None, g["x_%d" % t].to(DEV))["flow"]))
    loss.backward()
    m.detach_states()
    checked = 0
    for nm, q in m.named_parameters():
        if "grad_" + nm in g:
            ref = g["grad_" + nm]
            assert (q.grad.cpu() - ref).abs().max().item() <= 1e-3 * (ref.abs().max().item() + 1e-12), nm
            checked += 1
    assert checked >= 40


ZOO_CLASSES = {"rnnfirenet": "RNNFireNet", "leakyfirenet": "LeakyFireNet", "leakyfireflownet": "LeakyFireFlowNet",
               "rnnrecevflownet": "RNNRecEVFlowNet", "leakyrecevflownet": "LeakyRecEVFlowNet", "e2vid": "E2VID"}


@pytest.mark.parametrize("name", sorted(ZOO_CLASSES))
def test_ann_zoo_rollout_and_gradients_match_reference_golden(name):
    """SURVEY 8 f4 (ConvRecurrent / ConvLeaky* / ConvLSTM models): 3-step rollout and BPTT gradients vs the reference's numbers."""
    import event_flow_b200.models.model as M

    g = load_golden("annzoo_" + name)
    cls = ZOO_CLASSES[name]
    fire = "Fire" in cls
    cfg = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=8 if fire else 4, kernel_size=3,
               activations=["relu", None], mask_output=True, spiking_neuron={} if "Leaky" in cls else None)
    m = getattr(M, cls)(cfg)
    m.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("sd_")})
    m = m.to(DEV)
    T = len([k for k in g if k.startswith("x_")])
    with torch.no_grad():
        for t in range(T):
            out = m(None, g["x_%d" % t].to(DEV))
    for i, f in enumerate(out["flow"]):
        ref = g["flow_%d" % i]
        assert (f.cpu() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item() + 1e-7, f"flow {i}"
    m.reset_states()
    loss = 0.0
    for t in range(T):
        loss = loss + sum((f * g["gw_%d_%d" % (t, i)].to(DEV)).sum() for i, f in enumerate(m(None, g["x_%d" % t].to(DEV))["flow"]))
    loss.backward()
    m.detach_states()
    checked = 0
    for nm, q in m.named_parameters():
        if "grad_" + nm in g:
            ref = g["grad_" + nm]
            assert (q.grad.cpu() - ref).abs().max().item() <= 1e-3 * (ref.abs().max().item() + 1e-12), nm
            checked += 1
    assert checked >= 8


# ---- optional layers of the ANN U-Nets: normalisation and transposed-convolution decoders (models/submodules.py:45-49, 86-137) ----------------
# The reference's classes are thin wrappers of torch modules (nn.ConvTranspose2d, nn.BatchNorm2d, nn.InstanceNorm2d, torch activations), so
# the same torch modules on the same weights ARE the reference here; wiring against the live reference: tests/test_host_wiring_cpu.py.
@pytest.mark.parametrize("norm", [None, "BN", "IN"])
@pytest.mark.parametrize("act", ["relu", "tanh", None])
def test_transposed_conv_layer_matches_torch(norm, act, monkeypatch):
    from event_flow_b200.models.submodules import TransposedConvLayer

    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)  # the comparison is against torch's fp32 convolution, not its TF32 default
    torch.manual_seed(3)
    mine = TransposedConvLayer(6, 10, 3, activation=act, norm=norm).to(DEV).train()
    conv = torch.nn.ConvTranspose2d(6, 10, 3, stride=2, padding=1, output_padding=1, bias=norm != "BN").to(DEV)
    conv.load_state_dict(mine.transposed_conv2d.state_dict())
    nl = {None: None, "BN": torch.nn.BatchNorm2d(10), "IN": torch.nn.InstanceNorm2d(10, track_running_stats=True)}[norm]
    nl = None if nl is None else nl.to(DEV).train()
    x = torch.randn(2, 6, 9, 13, device=DEV, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    out = mine(x)
    ref = conv(x2)
    ref = ref if nl is None else nl(ref)
    ref = ref if act is None else getattr(torch, act)(ref)
    assert out.shape == ref.shape == (2, 10, 18, 26)
    assert (out - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    g = torch.randn_like(ref)
    out.backward(g), ref.backward(g)
    scale = max(p.grad.abs().max().item() for p in conv.parameters())
    assert (x.grad - x2.grad).abs().max().item() <= 1e-4 * x2.grad.abs().max().item()
    for (n, a), b in zip(mine.transposed_conv2d.named_parameters(), conv.parameters()):
        assert (a.grad - b.grad).abs().max().item() <= 1e-4 * max(b.grad.abs().max().item(), 1e-2 * scale), n
    if nl is not None:  # running statistics follow torch's modules
        assert torch.allclose(mine.norm_layer.running_mean, nl.running_mean, atol=1e-5) and torch.allclose(mine.norm_layer.running_var, nl.running_var, rtol=1e-4)


@pytest.mark.parametrize("norm", ["BN", "IN"])
def test_normalised_ann_layers_match_torch(norm, monkeypatch):
    import torch.nn.functional as F

    from event_flow_b200.models.submodules import ConvLayer, ResidualBlock, UpsampleConvLayer

    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    torch.manual_seed(4)
    mk = (lambda c: torch.nn.BatchNorm2d(c)) if norm == "BN" else (lambda c: torch.nn.InstanceNorm2d(c, track_running_stats=True))
    x = torch.randn(3, 8, 12, 16, device=DEV)
    for stride, k, act in ((1, 3, "relu"), (2, 3, "relu"), (1, 1, "tanh")):
        layer = ConvLayer(8, 12, k, stride=stride, activation=act, norm=norm).to(DEV).train()
        nl = mk(12).to(DEV).train()
        ref = getattr(torch, act)(nl(F.conv2d(x, layer.conv2d.weight, layer.conv2d.bias, stride, k // 2)))
        out = layer(x)
        assert out.shape == ref.shape and (out - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), (stride, k)
    up = UpsampleConvLayer(8, 6, 3, activation="relu", norm=norm).to(DEV).train()
    nl = mk(6).to(DEV).train()
    ref = torch.relu(nl(F.conv2d(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), up.conv2d.weight, up.conv2d.bias, 1, 1)))
    assert (up(x) - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    rb = ResidualBlock(8, 8, activation="relu", norm=norm).to(DEV).train()
    n1, n2 = mk(8).to(DEV).train(), mk(8).to(DEV).train()
    o1 = torch.relu(n1(F.conv2d(x, rb.conv1.weight, rb.conv1.bias, 1, 1)))
    o2 = torch.relu(n2(F.conv2d(o1, rb.conv2.weight, rb.conv2.bias, 1, 1)) + x)
    a2, a1 = rb(x)
    assert (a1 - o1).abs().max().item() <= 2e-5 * max(1.0, o1.abs().max().item()) and (a2 - o2).abs().max().item() <= 2e-5 * max(1.0, o2.abs().max().item())


def test_recevflownet_with_norm_and_transposed_decoders_trains():
    """RecEVFlowNet(norm="BN", use_upsample_conv=False): forward, backward and an optimiser step run on the CUDA path (values: the tests above)."""
    import event_flow_b200.models.model as M

    torch.manual_seed(0)
    cfg = dict(name="RecEVFlowNet", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=8, kernel_size=3,
               activations=["relu", None], mask_output=True, spiking_neuron=None, norm="BN", use_upsample_conv=False)
    m = M.RecEVFlowNet(cfg).to(DEV).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    cnt = torch.randint(0, 3, (2, 2, 64, 64), device=DEV).float()
    before = [p.detach().clone() for p in m.parameters()]
    flows = m(None, cnt)["flow"]
    assert len(flows) == 4 and all(f.shape == (2, 2, 64, 64) and torch.isfinite(f).all() for f in flows)
    sum(f.square().mean() for f in flows).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    opt.step()
    assert any((a - b).abs().max() > 0 for a, b in zip(m.parameters(), before))


# (the gated form -- ConvGRU -- is stride 1 by construction: its blend operands live at the input resolution)
@pytest.mark.parametrize("stride,gated", [(1, True), (1, False), (2, False)])
@pytest.mark.parametrize("act", ["tanh", "relu", "sigmoid", None])
def test_conv_ann_every_input_gradient_matches_torch_autograd(stride, act, gated, monkeypatch):
    """
    ops.conv_ann with every option (second input with gate product, residual, gated blend, stride 2): the value and the gradient of EVERY
    tensor argument against torch's own ops and autograd -- the backward here is kernels only (ef_ann_gate_bwd, ef_ann_cat_scale,
    ef_conv3x3_bwd_s, ef_ann_scale_bwd).
    """
    import torch.nn.functional as F

    from event_flow_b200 import ops

    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    g = torch.Generator().manual_seed(9)
    B, C1, C2, Co, H, W = 2, 5, 7, 9, 19, 22
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    mk = lambda *s: torch.randn(*s, generator=g).to(DEV).requires_grad_(True)  # noqa: E731
    big = torch.rand(B, 2 * C2, H, W, generator=g).to(DEV).requires_grad_(True)  # gate tensor: its channel halves are slices (non-dense batch stride)
    t = {"x1": mk(B, C1, H, W), "x2": mk(B, C2, H, W), "w": mk(Co, C1 + C2, 3, 3), "b": mk(Co), "res": mk(B, Co, Ho, Wo)}
    if gated:
        t["h"], t["u"] = mk(B, Co, H, W), torch.rand(B, Co, H, W, generator=g).to(DEV).requires_grad_(True)
    scale = big[:, C2:] if gated else None
    out = ops.conv_ann(t["x1"], t["w"], t["b"], act, x2=t["x2"], x2_scale=scale, residual=t["res"], blend_h=t.get("h"), blend_u=t.get("u"), stride=stride)
    g_out = torch.randn(out.shape, generator=g).to(DEV)
    out.backward(g_out)
    mine = {k: v.grad.clone() for k, v in t.items()}
    mine["big"] = None if big.grad is None else big.grad.clone()
    for v in list(t.values()) + [big]:
        v.grad = None
    x = torch.cat([t["x1"], t["x2"] * scale if gated else t["x2"]], 1)
    ref = F.conv2d(x, t["w"], t["b"], stride, 1) + t["res"]
    ref = ref if act is None else getattr(torch, act)(ref)
    if gated:
        ref = t["h"] * (1 - t["u"]) + ref * t["u"]
    assert out.shape == ref.shape == (B, Co, Ho, Wo)
    assert (out - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    ref.backward(g_out)
    for k, v in t.items():
        assert (mine[k] - v.grad).abs().max().item() <= 1e-4 * (v.grad.abs().max().item() + 1e-12), k
    if gated:
        assert (mine["big"] - big.grad).abs().max().item() <= 1e-4 * big.grad.abs().max().item()
