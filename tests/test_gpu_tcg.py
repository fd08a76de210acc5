"""
GPU parity of the general-channel tensor-core cell (ef_lif_conv_fwd_g: K / N blocked, several concatenated sources, streamed
weights) against the CPU oracle: the cell shapes of the EV-FlowNet family.  Tolerances as T1: |dv| <= max(2e-5, 3e-6 max|v|)
(x sqrt(K/576) for longer sums), spikes exact outside that band.
"""
import math

import pytest
import torch

from oracle import spiking as osp
from tests.util import spike_band_compare

pytestmark = pytest.mark.gpu
DEV = "cuda"


def lif_params(cin, c, rec, seed, gain=2.0):
    g = torch.Generator().manual_seed(seed)
    p = {"ff": (torch.rand((c, cin, 3, 3), generator=g) * 2 - 1) * math.sqrt(1 / cin) * gain,
         "leak": torch.randn((c, 1, 1), generator=g) * 0.1 - 4.0, "thresh": torch.randn((c, 1, 1), generator=g) * 0.1 + 0.8}
    if rec:
        p["rec"] = (torch.rand((c, c, 3, 3), generator=g) * 2 - 1) * math.sqrt(1 / c) * gain
    return p


def pad_cl(x, ops):
    """fp32 NCHW with any channel count -> cl bf16 with the channels zero-padded to a multiple of 32."""
    B, C, H, W = x.shape
    Cp = (C + 31) // 32 * 32
    if Cp != C:
        x = torch.cat([x, torch.zeros(B, Cp - C, H, W)], 1)
    return ops.pack_cl(x.to(DEV))


def check(v, z, ns_o, thr, K):
    scale = max(1.0, math.sqrt(K / 576))
    v_atol = scale * max(2e-5, 3e-6 * ns_o[0].abs().max().item())
    spike_band_compare(v.cpu(), z, ns_o[0], ns_o[1], thr, v_atol=v_atol)


@pytest.mark.parametrize("C,shape", [(64, (2, 37, 52)), (128, (1, 32, 32)), (512, (1, 16, 16))])
@pytest.mark.parametrize("hard", [True, False])
@pytest.mark.parametrize("with_state", [True, False])
def test_recurrent_cell_general_channels(C, shape, hard, with_state):
    """ConvLIFRecurrent C -> C (encoder stages of the spiking U-Net): the recurrent convolution is a second input source."""
    from event_flow_b200 import ops

    B, H, W = shape
    g = torch.Generator().manual_seed(C + H)
    p = lif_params(C, C, True, C)
    x = (torch.rand((B, C, H, W), generator=g) < 0.25).float()
    st = None
    if with_state:
        st = torch.rand((2, B, C, H, W), generator=g) * 1.2 - 0.1
        st[1] = (st[1] < 0.3).float()
    pd = {k: v.to(DEV).contiguous() for k, v in p.items()}
    x_cl = ops.pack_cl(x.to(DEV))
    v_in = z_in = None
    srcs, wsrcs = [x_cl], [(pd["ff"], 0, C, False)]
    if st is not None:
        v_in, z_in = st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV))
        srcs.append(z_in)
        wsrcs.append((pd["rec"], 0, C, False))
    img = ops.split_weights_g(wsrcs, C)
    v, z, _ = ops.lif_step_g(srcs, v_in, z_in, img, pd["leak"].reshape(-1), pd["thresh"].reshape(-1), C, hard_reset=hard)
    out_o, ns_o = osp.cell_step("lif", x, st, p, hard_reset=hard)
    zf = ops.unpack_cl(z).cpu()
    check(v, zf, ns_o, p["thresh"].clamp_min(0.01), 9 * C * (2 if st is not None else 1))
    assert zf.mean() > 0.01


@pytest.mark.parametrize("hard", [True, False])
def test_decoder_cell_three_sources_with_fractional_flow(hard):
    """
    Decoder stage (unet.py:451-462): input = cat[upsampled prediction (2 fractional fp32 channels), upsampled x, upsampled skip];
    the concat is never built: three sources, the prediction as an exact hi/mid/lo split with the weight rows repeated per slot.
    """
    from event_flow_b200 import ops

    B, H, W, Cx, C = 2, 40, 48, 64, 32
    g = torch.Generator().manual_seed(3)
    p = lif_params(2 + 2 * Cx, C, False, 5, gain=1.5)
    pred = torch.tanh(torch.randn((B, 2, H, W), generator=g))                           # fractional
    xu = torch.randint(0, 33, (B, Cx, H, W), generator=g).float() / 16.0 * (torch.rand((B, Cx, H, W), generator=g) < 0.3)   # k/16
    sk = torch.randint(0, 17, (B, Cx, H, W), generator=g).float() / 16.0 * (torch.rand((B, Cx, H, W), generator=g) < 0.3)
    st = torch.rand((2, B, C, H, W), generator=g) * 1.2 - 0.1
    st[1] = (st[1] < 0.3).float()
    pd = {k: v.to(DEV).contiguous() for k, v in p.items()}
    srcs = [ops.pack_split_cl(pred.to(DEV)), ops.pack_cl(xu.to(DEV)), ops.pack_cl(sk.to(DEV))]
    img = ops.split_weights_g([(pd["ff"], 0, 2, True), (pd["ff"], 2, Cx, False), (pd["ff"], 2 + Cx, Cx, False)], C)
    v_in, z_in = st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV))
    v, z, _ = ops.lif_step_g(srcs, v_in, z_in, img, pd["leak"].reshape(-1), pd["thresh"].reshape(-1), C, hard_reset=hard)
    out_o, ns_o = osp.cell_step("lif", torch.cat([pred, xu, sk], 1), st, p, hard_reset=hard)
    zf = ops.unpack_cl(z).cpu()
    check(v, zf, ns_o, p["thresh"].clamp_min(0.01), 9 * (2 + 2 * Cx))
    assert zf.mean() > 0.01


def test_residual_cell_and_channel_padding():
    """Second cell of a spiking residual block (spiking_submodules.py:933-975): out = spikes + block input; C = 96 (three output
    blocks), a source whose channel count is not a multiple of 32 (zero-padded tensor, `n` real channels in the weight slice)."""
    from event_flow_b200 import ops

    B, H, W, C, Cin = 2, 24, 36, 96, 80
    g = torch.Generator().manual_seed(8)
    p = lif_params(Cin, C, False, 9)
    x = (torch.rand((B, Cin, H, W), generator=g) < 0.3).float()
    res = (torch.rand((B, C, H, W), generator=g) < 0.4).float()
    st = torch.rand((2, B, C, H, W), generator=g) * 1.2 - 0.1
    st[1] = (st[1] < 0.3).float()
    pd = {k: v.to(DEV).contiguous() for k, v in p.items()}
    img = ops.split_weights_g([(pd["ff"], 0, Cin, False)], C)
    v, z, out = ops.lif_step_g([pad_cl(x, ops)], st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV)), img, pd["leak"].reshape(-1),
                               pd["thresh"].reshape(-1), C, hard_reset=True, residual_cl=ops.pack_cl(res.to(DEV)))
    out_o, ns_o = osp.cell_step("lif", x, st, p, hard_reset=True, residual=res)
    zf = ops.unpack_cl(z).cpu()
    check(v, zf, ns_o, p["thresh"].clamp_min(0.01), 9 * Cin)
    assert torch.equal(ops.unpack_cl(out).cpu(), zf + res)


@pytest.mark.parametrize("cin,C,shape", [(2, 64, (2, 64, 96)), (64, 128, (2, 32, 48)), (96, 64, (1, 40, 56))])
@pytest.mark.parametrize("with_state", [True, False])
def test_stride2_cell_on_space_to_depth_input(cin, C, shape, with_state):
    """Stride-2 ConvLIF (the U-Net encoders) as a stride-1 tensor-core cell over the space-to-depth input; the first encoder's fp32
    counts additionally as an exact hi/mid/lo split.  Against the oracle's stride-2 cell."""
    from event_flow_b200 import ops

    B, H, W = shape
    g = torch.Generator().manual_seed(cin + H)
    p = lif_params(cin, C, False, cin + 1, gain=2.5)
    if cin == 2:
        x = torch.randint(0, 6, (B, cin, H, W), generator=g).float() * (torch.rand((B, cin, H, W), generator=g) < 0.5)
        x[0, 0, 0, 0], x[0, 1, 3, 5] = 300.0, 1000.0  # counts beyond the bf16 integer range: the split keeps them exact
        src = ops.pack_split_s2d_cl(x.to(DEV))
    else:
        x = (torch.rand((B, cin, H, W), generator=g) < 0.3).float()
        src = ops.space_to_depth_cl(pad_cl(x, ops)) if cin % 32 else ops.space_to_depth_cl(ops.pack_cl(x.to(DEV)))
    Ho, Wo = H // 2, W // 2
    st = None
    if with_state:
        st = torch.rand((2, B, C, Ho, Wo), generator=g) * 1.2 - 0.1
        st[1] = (st[1] < 0.3).float()
    pd = {k: v.to(DEV).contiguous() for k, v in p.items()}
    n_real = cin if cin % 32 == 0 or cin == 2 else (cin + 31) // 32 * 32  # a zero-padded tensor: the padded channels are part of the s2d layout
    w = pd["ff"]
    if n_real != cin:  # weights of the padded channels are zero
        w = torch.cat([w, torch.zeros(C, n_real - cin, 3, 3, device=DEV)], 1).contiguous()
    img = ops.split_weights_g([(w, 0, n_real, cin == 2, True)], C)
    v_in = z_in = None
    if st is not None:
        v_in, z_in = st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV))
    v, z, _ = ops.lif_step_g([src], v_in, z_in, img, pd["leak"].reshape(-1), pd["thresh"].reshape(-1), C, hard_reset=True, s2d=True)
    out_o, ns_o = osp.cell_step("lif", x, st, p, hard_reset=True, stride=2)
    zf = ops.unpack_cl(z).cpu()
    check(v, zf, ns_o, p["thresh"].clamp_min(0.01), 9 * cin)
    assert zf.mean() > 0.01


@pytest.mark.parametrize("neuron", ["lif", "alif"])
@pytest.mark.parametrize("cin,C,stride,rec,shape", [(64, 96, 1, True, (2, 24, 32)), (32, 64, 2, False, (2, 32, 48)), (66, 32, 1, False, (1, 20, 24)),
                                                    (128, 128, 1, True, (1, 16, 16))])
def test_data_gradient_as_convolution_on_the_general_kernel(neuron, cin, C, stride, rec, shape):
    """
    Backward of a cell step with other channel counts than 32 (the U-Net family): neuron backward + the data gradients as plain
    convolutions of g_I (two bf16 terms, zero-inserted for stride 2, flipped / transposed weights) on ef_lif_conv_fwd_g, against the
    CUDA-core data gradient of ef_lif_conv_bwd on the same tensors: 1e-4 of each gradient's scale (g_I carries 16 significant bits).
    """
    from event_flow_b200 import ops

    B, H, W = shape
    g = torch.Generator().manual_seed(cin + C + stride)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    x = (torch.rand(B, cin, H, W, generator=g) < 0.2).float()
    n_state = 2 if neuron == "lif" else 3
    st = torch.randn(n_state, B, C, Ho, Wo, generator=g) * 0.5
    st[1] = (st[1] > 0.3).float()
    if n_state == 3:
        st[2] = st[2].abs() * 0.3
    w = {"ff": (torch.rand(C, cin, 3, 3, generator=g) * 2 - 1) * math.sqrt(1 / cin) * 2.0}
    if rec:
        w["rec"] = (torch.rand(C, C, 3, 3, generator=g) * 2 - 1) * math.sqrt(1 / C) * 2.0
    chan = {n: (torch.randn(C, 1, 1, generator=g) * 0.1 + (0.8 if n in ("thresh", "t0") else (-4.0 if n == "leak" else 0.1))).to(DEV)
            for n in ops.param_names(neuron)}
    g_out, g_ns = torch.randn(B, C, Ho, Wo, generator=g).to(DEV), torch.randn(n_state, B, C, Ho, Wo, generator=g).to(DEV)
    res = {}
    for tc in (True, False):
        ops.TCG_BACKWARD = tc
        try:
            xd, sd = x.to(DEV).requires_grad_(True), st.to(DEV).requires_grad_(True)
            ws = {k: v.to(DEV).requires_grad_(True) for k, v in w.items()}
            n0 = ops.L.lib().ef_launch_count() if hasattr(ops.L.lib(), "ef_launch_count") else 0
            out, ns = ops.cell_step(neuron, xd, sd, ws["ff"], ws.get("rec"), chan, hard_reset=True, stride=stride)
            torch.autograd.backward([out, ns], [g_out, g_ns])
            res[tc] = (xd.grad.cpu(), sd.grad.cpu(), {k: v.grad.cpu() for k, v in ws.items()})
        finally:
            ops.TCG_BACKWARD = True
    for a, b, what in ((res[True][0], res[False][0], "g_x"), (res[True][1], res[False][1], "g_state")) + tuple(
            (res[True][2][k], res[False][2][k], "g_" + k) for k in w):
        assert (a - b).abs().max().item() <= 1e-4 * (b.abs().max().item() + 1e-20), what
    assert not torch.equal(res[True][0], res[False][0])  # (the two paths really are different kernels)


@pytest.mark.parametrize("cin,C,stride,rec,kind,residual,shape", [
    (64, 96, 1, True, "spikes", False, (2, 24, 32)),       # recurrent block of an encoder
    (32, 64, 2, False, "spikes", False, (2, 32, 48)),      # stride-2 encoder convolution (space-to-depth source)
    (64, 64, 1, False, "spikes", True, (1, 20, 24)),       # second cell of a residual block
    (66, 32, 1, False, ("mixed", 2), False, (2, 16, 32)),  # decoder: [flow prediction (fractional), x, skip]
    (128, 128, 1, True, "spikes", False, (1, 16, 16))])
@pytest.mark.parametrize("with_state", [True, False])
def test_training_forward_of_general_lif_cells_on_the_tensor_cores(cin, C, stride, rec, kind, residual, shape, with_state):
    """
    ops.cell_step under autograd for LIF cells with other channel counts than 32 (x_kind vouching for the input): the fused general
    tcgen05 kernel against the fused CUDA-core kernel and the oracle -- membrane to summation-order noise, spikes exact outside the band,
    gradients (which read v_out, never the emitted spikes) to 1e-4 of their scale.
    """
    from event_flow_b200 import ops

    B, H, W = shape
    g = torch.Generator().manual_seed(cin + 3 * C + stride)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    x = (torch.rand(B, cin, H, W, generator=g) < 0.2).float() * torch.randint(1, 3, (B, cin, H, W), generator=g).float() / 16 * 16
    if type(kind) is tuple:
        x[:, :kind[1]] = torch.randn(B, kind[1], H, W, generator=g)
    st = None
    if with_state:
        st = torch.randn(2, B, C, Ho, Wo, generator=g) * 0.5
        st[1] = (st[1] > 0.3).float()
    params = lif_params(cin, C, rec, seed=5)
    res_t = (torch.rand(B, C, Ho, Wo, generator=g) < 0.3).float() if residual else None
    chan = {"leak": params["leak"].to(DEV), "thresh": params["thresh"].to(DEV)}
    g_out, g_ns = torch.randn(B, C, Ho, Wo, generator=g).to(DEV), torch.randn(2, B, C, Ho, Wo, generator=g).to(DEV)
    g_ns[1] = 0
    res = {}
    for tc in (True, False):
        ops.TCG_FORWARD = tc
        try:
            xd = x.to(DEV).requires_grad_(True)
            sd = None if st is None else st.to(DEV).requires_grad_(True)
            ws = {k: params[k].to(DEV).requires_grad_(True) for k in ("ff", "rec") if k in params}
            rd = None if res_t is None else res_t.to(DEV)
            n0 = ops.L.LAUNCHES
            out, ns = ops.cell_step("lif", xd, sd, ws["ff"], ws.get("rec"), chan, hard_reset=True, stride=stride, x_kind=kind, residual=rd)
            torch.autograd.backward([out, ns], [g_out, g_ns])
            res[tc] = (out.detach().cpu(), ns.detach().cpu(), xd.grad.cpu(), None if sd is None else sd.grad.cpu(), {k: v.grad.cpu() for k, v in ws.items()})
        finally:
            ops.TCG_FORWARD = True
    out_o, ns_o = osp.cell_step("lif", x, st, params, hard_reset=True, stride=stride, residual=0 if res_t is None else res_t)
    thr = params["thresh"].clamp_min(0.01)
    K = 9 * (cin + (C if rec and with_state else 0))
    scale = max(1.0, math.sqrt(K / 576))
    for other in (res[False][1], ns_o):
        spike_band_compare(res[True][1][0], res[True][1][1], other[0], other[1], thr, v_atol=scale * max(2e-5, 3e-6 * other[0].abs().max().item()))
    extra = 0 if res_t is None else res_t
    assert torch.equal(res[True][0], res[True][1][1] + extra)
    assert not torch.equal(res[True][1][0], res[False][1][0])  # (different kernels: only the summation order may differ)
    for a, b, what in ((res[True][2], res[False][2], "g_x"), (res[True][3], res[False][3], "g_state")) + tuple(
            (res[True][4][k], res[False][4][k], "g_" + k) for k in res[False][4]):
        if b is not None:
            assert (a - b).abs().max().item() <= 1e-4 * (b.abs().max().item() + 1e-20), what


@pytest.mark.parametrize("neuron", ["plif", "alif", "xlif"])
@pytest.mark.parametrize("cin,C,rec,kind,residual", [(64, 96, True, "spikes", False), (64, 64, False, "spikes", True), (66, 32, False, ("mixed", 2), False)])
def test_other_neuron_kinds_on_the_general_kernel(neuron, cin, C, rec, kind, residual):
    """
    PLIF / ALIF / XLIF cells with other channel counts than 32 under autograd: the general tensor-core kernel as a pure convolution +
    ef_lif_neuron_fwd on its current (and the tensor-core gradients where they apply) against the fused CUDA-core kernels on the same
    tensors: state to summation-order noise, gradients to 1e-4 of their scale.
    """
    from event_flow_b200 import ops

    B, H, W = 2, 20, 24
    g = torch.Generator().manual_seed(cin + C + len(neuron))
    x = (torch.rand(B, cin, H, W, generator=g) < 0.2).float()
    if type(kind) is tuple:
        x[:, :kind[1]] = torch.randn(B, kind[1], H, W, generator=g)
    st = torch.randn(3, B, C, H, W, generator=g) * 0.5
    st[1] = (st[1] > 0.3).float()
    st[2] = st[2].abs() * 0.3
    w = {"ff": (torch.rand(C, cin, 3, 3, generator=g) * 2 - 1) * math.sqrt(1 / cin) * 2.0}
    if rec:
        w["rec"] = (torch.rand(C, C, 3, 3, generator=g) * 2 - 1) * math.sqrt(1 / C) * 2.0
    base = {"thresh": 0.8, "t0": 0.8, "leak": -4.0, "leak_v": -4.0}
    chan = {n: (torch.randn(C, 1, 1, generator=g) * 0.1 + base.get(n, 0.1)).to(DEV) for n in ops.param_names(neuron)}
    res_t = (torch.rand(B, C, H, W, generator=g) < 0.3).float().to(DEV) if residual else None
    g_out, g_ns = torch.randn(B, C, H, W, generator=g).to(DEV), torch.randn(3, B, C, H, W, generator=g).to(DEV)
    g_ns[1] = 0
    res = {}
    for tc in (True, False):
        ops.TCG_FORWARD = tc
        try:
            xd, sd = x.to(DEV).requires_grad_(True), st.to(DEV).requires_grad_(True)
            ws = {k: v.to(DEV).requires_grad_(True) for k, v in w.items()}
            out, ns = ops.cell_step(neuron, xd, sd, ws["ff"], ws.get("rec"), chan, hard_reset=True, x_kind=kind, residual=res_t)
            torch.autograd.backward([out, ns], [g_out, g_ns])
            res[tc] = (out.detach().cpu(), ns.detach().cpu(), xd.grad.cpu(), sd.grad.cpu(), {k: v.grad.cpu() for k, v in ws.items()})
        finally:
            ops.TCG_FORWARD = True
    v_t, v_c = res[True][1][0], res[False][1][0]
    assert not torch.equal(v_t, v_c)
    assert (v_t - v_c).abs().max().item() <= max(2e-5, 3e-6 * v_c.abs().max().item())
    assert (res[True][1][1] != res[False][1][1]).float().mean().item() < 1e-4  # (borderline spikes may flip with the summation order)
    for a, b, what in ((res[True][2], res[False][2], "g_x"), (res[True][3], res[False][3], "g_state")) + tuple(
            (res[True][4][k], res[False][4][k], "g_" + k) for k in w):
        assert (a - b).abs().max().item() <= 1e-4 * (b.abs().max().item() + 1e-20), what
