"""
GPU parity of the tcgen05 (tensor-core) fused conv + LIF kernel: against the fp32 CUDA-core kernel on identical cl inputs,
and against the CPU oracle.  Same tolerances as T1: |dv| <= 2e-5, spikes exact outside |v - thresh| < 1e-5.
"""
import pytest
import torch

from oracle import spiking as osp
from tests.util import spike_band_compare

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_case(B, H, W, rec, seed, with_state=True, density=0.3):
    g = torch.Generator().manual_seed(seed)
    params = osp.init_firenet_params("lif", 32, 32, seed=seed, weight_gain=2.0)["G1" if rec else "R1a"]
    x = (torch.rand((B, 32, H, W), generator=g) < density).float()
    st = None
    if with_state:
        st = torch.rand((2, B, 32, H, W), generator=g) * 1.2 - 0.1
        st[1] = (st[1] < 0.3).float()
    return params, x, st


@pytest.mark.parametrize("rec", [False, True])
@pytest.mark.parametrize("hard", [True, False])
@pytest.mark.parametrize("shape", [(1, 16, 8), (2, 37, 52), (8, 128, 128), (3, 16, 20), (1, 5, 3), (2, 33, 47)])
@pytest.mark.parametrize("with_state", [True, False])
def test_tc_kernel_matches_cuda_core_kernel_and_oracle(rec, hard, shape, with_state):
    from event_flow_b200 import ops

    B, H, W = shape
    params, x, st = make_case(B, H, W, rec, seed=B * 7 + H, with_state=with_state)
    pd = {k: v.to(DEV).contiguous() for k, v in params.items()}
    x_cl = ops.pack_cl(x.to(DEV))
    v_in = z_in = None
    if st is not None:
        v_in, z_in = st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV))
    w_split = ops.split_weights(pd["ff"], pd.get("rec"))
    assert w_split is not None
    leak, thresh = pd["leak"].reshape(-1), pd["thresh"].reshape(-1)
    v_tc, z_tc = ops.lif_step_cl(x_cl, v_in, z_in, pd["ff"], pd.get("rec"), leak, thresh, hard_reset=hard, w_split=w_split)
    v_cc, z_cc = ops.lif_step_cl(x_cl, v_in, z_in, pd["ff"], pd.get("rec"), leak, thresh, hard_reset=hard, w_split=None)
    torch.cuda.synchronize()
    thr = params["thresh"].clamp_min(0.01)
    z_tc_f, z_cc_f = ops.unpack_cl(z_tc).cpu(), ops.unpack_cl(z_cc).cpu()
    spike_band_compare(v_tc.cpu(), z_tc_f, v_cc.cpu(), z_cc_f, thr)
    out_o, ns_o = osp.cell_step("lif", x, st, params, hard_reset=hard)
    dv, _, in_flips = spike_band_compare(v_tc.cpu(), z_tc_f, ns_o[0], ns_o[1], thr)
    assert z_tc_f.mean() > 0.02  # spikes present: not a vacuous comparison


def test_weight_split_is_exact():
    from event_flow_b200 import ops

    g = torch.Generator().manual_seed(0)
    w = (torch.rand((32, 32, 3, 3), generator=g) * 2 - 1) * 0.35
    w[0, 0, 0, 0], w[0, 0, 0, 1], w[0, 0, 0, 2] = 1.0, 3.0e-5, -0.333333343267
    wr = torch.randn((32, 32, 3, 3), generator=g)
    sp = ops.split_weights(w.to(DEV), wr.to(DEV)).cpu()
    assert sp.numel() == 2 * 9 * 96 * 32 + 8  # + 16 bytes reserved for the scales of the (disabled) two-term variant
    sp = sp[:2 * 9 * 96 * 32]
    raw = ((sp.view(2, 9, 12, 8, 4, 8).to(torch.int32) & 0xFFFF) << 16).view(torch.float32)  # conv, tap, atom, row, phys chunk, k%8
    vals = torch.empty(2, 9, 12, 8, 4, 8)
    for r in range(8):  # undo the 64-byte swizzle: logical chunk = physical chunk ^ ((row >> 1) & 3)
        for c in range(4):
            vals[:, :, :, r, c] = raw[:, :, :, r, c ^ ((r >> 1) & 3)]
    vals = vals.reshape(2, 9, 3, 32, 32).permute(0, 2, 1, 3, 4)  # -> [conv, split, tap, n, ci]
    for cv, ref in enumerate((w, wr)):
        total = (vals[cv, 2] + vals[cv, 1]) + vals[cv, 0]  # lo + mid + hi, exact in fp32
        assert torch.equal(total.permute(1, 2, 0).reshape(32, 32, 3, 3), ref)


@pytest.mark.parametrize("cin,hard,with_state", [(5, True, True), (2, True, False), (5, False, True), (8, True, True)])
def test_head_kernel_matches_generic_and_oracle(cin, hard, with_state):
    """Dedicated head kernel (few fractional input channels -> 32 LIF channels, cl spikes out)."""
    from event_flow_b200 import ops

    B, H, W = 3, 37, 70
    g = torch.Generator().manual_seed(cin)
    params = osp.init_firenet_params("lif", cin, 32, seed=2, weight_gain=3.0)["head"]
    x = torch.randn((B, cin, H, W), generator=g) * (torch.rand((B, cin, H, W), generator=g) < 0.4)
    st = None
    if with_state:
        st = torch.rand((2, B, 32, H, W), generator=g) * 1.2 - 0.1
        st[1] = (st[1] < 0.3).float()
    pd = {k: v.to(DEV).contiguous() for k, v in params.items()}
    v_in = z_in = None
    if st is not None:
        v_in, z_in = st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV))
    v_h, z_h = ops.lif_step_cl(None, v_in, z_in, pd["ff"], None, pd["leak"].reshape(-1), pd["thresh"].reshape(-1), hard_reset=hard,
                               x_f32=x.to(DEV).contiguous())
    out_o, ns_o = osp.cell_step("lif", x, st, params, hard_reset=hard)
    spike_band_compare(v_h.cpu(), ops.unpack_cl(z_h).cpu(), ns_o[0], ns_o[1], params["thresh"].clamp_min(0.01))
    assert ops.unpack_cl(z_h).mean() > 0.02


@pytest.mark.parametrize("rec", [False, True])
def test_tc_kernel_repeated_launches_are_bit_identical(rec):
    """
    Race detector.  The kernel is a 10-warp producer / MMA / epilogue pipeline over mbarriers; a missing dependency shows up
    as a few wrong values in a few launches (it did during development: tools/tc_determinism.py).  Launch a multi-tile-per-CTA
    problem repeatedly: every launch must be bit-identical to the first and match the CUDA-core kernel.
    """
    from event_flow_b200 import ops

    B, H, W = 16, 128, 128  # 2048 tiles on 148 persistent CTAs: ~14 pipelined tiles per CTA
    params, x, st = make_case(B, H, W, rec, seed=5)
    pd = {k: v.to(DEV).contiguous() for k, v in params.items()}
    x_cl, v_in, z_in = ops.pack_cl(x.to(DEV)), st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV))
    ws = ops.split_weights(pd["ff"], pd.get("rec"))
    args = (x_cl, v_in, z_in, pd["ff"], pd.get("rec"), pd["leak"].reshape(-1), pd["thresh"].reshape(-1))
    v_cc, z_cc = ops.lif_step_cl(*args, hard_reset=True, w_split=None)
    v0, z0 = ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
    assert (v0 - v_cc).abs().max() < 1e-4 and (z0 != z_cc).float().mean() < 1e-5
    for _ in range(30):
        v, z = ops.lif_step_cl(*args, hard_reset=True, w_split=ws)
        assert torch.equal(v, v0) and torch.equal(z, z0)


@pytest.mark.parametrize("rec", [False, True])
@pytest.mark.parametrize("shape", [(1, 16, 8), (2, 37, 52), (8, 128, 128), (1, 5, 3)])
@pytest.mark.parametrize("with_state", [True, False])
def test_tc_weight_gradient_matches_cuda_core_and_fp64(rec, shape, with_state):
    """
    Weight gradient on tcgen05 (MN-major operands, per-CTA partial sums, fixed-order reduction) against the CUDA-core
    kernel on the same g_I terms and against an fp64 correlation of the same operands.  Tolerance 1e-4 relative to the
    largest entry (fp32 accumulation order differs; the operands are identical bf16 values).  Three calls under the
    accumulate / finalize protocol must give three times the gradient of one call.
    """
    import torch.nn.functional as F

    from event_flow_b200 import ops

    B, H, W = shape
    params, x, st = make_case(B, H, W, rec, seed=B * 11 + W, with_state=with_state)
    pd = {k: v.to(DEV).contiguous() for k, v in params.items()}
    g = torch.Generator().manual_seed(3)
    x_cl = ops.pack_cl(x.to(DEV))
    v_in = z_in = None
    if st is not None:
        v_in, z_in = st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV))
    leak, thresh = pd["leak"].reshape(-1), pd["thresh"].reshape(-1)
    v_out, _ = ops.lif_step_cl(x_cl, v_in, z_in, pd["ff"], pd.get("rec"), leak, thresh)
    g_out = torch.randn((B, 32, H, W), generator=g).to(DEV)
    g_v = torch.randn((B, 32, H, W), generator=g).to(DEV) * 0.1
    args = (x_cl, v_in, z_in, v_out, g_out, g_v, None, pd["ff"], pd.get("rec"), leak, thresh)
    tc = ops.lif_bwd_cl(*args, tc_wgrad=True)
    cc = ops.lif_bwd_cl(*args, tc_wgrad=False)
    tc3 = ops.lif_bwd_cl(*args, tc_wgrad=True, steps=3)
    torch.cuda.synchronize()
    names = ["g_w_ff"] + (["g_w_rec"] if rec else [])
    gI = tc["gI"].double().permute(0, 3, 1, 2)  # [B,co,H,W]
    for n in names:
        src = x if n == "g_w_ff" else (st[1] if st is not None else torch.zeros_like(x))
        xp = F.pad(src.double().to(DEV), (1, 1, 1, 1))
        ref = torch.stack([torch.stack([torch.einsum("bchw,bkhw->kc", xp[:, :, dy:dy + H, dx:dx + W], gI) for dx in range(3)], -1)
                           for dy in range(3)], -2)  # [co,ci,dy,dx]
        scale = ref.abs().max().item() + 1e-12
        assert (tc[n].double() - ref).abs().max().item() <= 1e-4 * scale, n
        assert (cc[n].double() - ref).abs().max().item() <= 1e-4 * scale, n
        assert (tc3[n].double() - 3 * ref).abs().max().item() <= 3e-4 * scale, n
        assert scale > 1e-3 or st is None
    for n in ("g_x", "g_v_in"):
        assert torch.equal(tc[n], cc[n]), n
    for n in ("g_leak", "g_thresh"):  # block sums land through atomics: order-dependent rounding (relative to the largest entry)
        assert (tc[n] - cc[n]).abs().max().item() <= 1e-4 * cc[n].abs().max().item() + 1e-6, n


def test_tc_weight_gradient_is_bit_reproducible():
    from event_flow_b200 import ops

    B, H, W = 8, 128, 128
    params, x, st = make_case(B, H, W, True, seed=9)
    pd = {k: v.to(DEV).contiguous() for k, v in params.items()}
    x_cl, v_in, z_in = ops.pack_cl(x.to(DEV)), st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV))
    leak, thresh = pd["leak"].reshape(-1), pd["thresh"].reshape(-1)
    v_out, _ = ops.lif_step_cl(x_cl, v_in, z_in, pd["ff"], pd["rec"], leak, thresh)
    g_out = torch.randn((B, 32, H, W), generator=torch.Generator().manual_seed(1)).to(DEV)
    args = (x_cl, v_in, z_in, v_out, g_out, None, None, pd["ff"], pd["rec"], leak, thresh)
    first = ops.lif_bwd_cl(*args, tc_wgrad=True)
    for _ in range(10):
        again = ops.lif_bwd_cl(*args, tc_wgrad=True)
        assert torch.equal(again["g_w_ff"], first["g_w_ff"]) and torch.equal(again["g_w_rec"], first["g_w_rec"])


@pytest.mark.parametrize("cin", [1, 2, 5, 10])
def test_head_layer_on_tensor_cores_split_input(cin):
    """
    Head layer (fractional fp32 voxel inputs, Cin <= 10) through the tensor-core cell kernel: ef_pack_split_cl writes the exact
    hi/mid/lo bf16 split of the input, ef_split_weights_head the weight image with w[.,c] in the three slots of c -> nine exact
    partial products per (input, weight) pair, fp32 accumulate.  Split exactness bit for bit; membrane / spikes vs the CPU oracle
    and vs the fp32 CUDA-core head kernel within the T1 band; weight gradient through ef_lif_bwd_window vs oracle autograd.
    """
    from event_flow_b200 import _lib as L
    from event_flow_b200 import ops

    B, H, W = 2, 37, 52
    g = torch.Generator().manual_seed(cin)
    params = osp.init_firenet_params("lif", cin, 32, seed=cin, weight_gain=3.0)["head"]
    x = torch.randn((B, cin, H, W), generator=g) * (torch.rand((B, cin, H, W), generator=g) < 0.4)
    x[0, 0, 0, :4] = torch.tensor([1.0, 3.0e-5, -0.333333343267, 1.0e-30])[: min(4, W)]
    st = torch.rand((2, B, 32, H, W), generator=g) * 1.2 - 0.1
    st[1] = (st[1] < 0.3).float()
    pd = {k: v.to(DEV).contiguous() for k, v in params.items()}
    x_cl = ops.pack_split_cl(x.to(DEV))
    assert x_cl.shape == (B, H, W, 32) and x_cl.dtype == torch.bfloat16
    SL = 8 if cin <= 8 else 10
    parts = x_cl.float().cpu()
    rebuilt = parts[..., 0:cin] + parts[..., SL:SL + cin] + parts[..., 2 * SL:2 * SL + cin]  # (hi + mid) + lo is exact in fp32
    assert torch.equal(rebuilt.permute(0, 3, 1, 2), x) or torch.equal((parts[..., 0:cin].double() + parts[..., SL:SL + cin].double()
                                                                      + parts[..., 2 * SL:2 * SL + cin].double()).float().permute(0, 3, 1, 2), x)
    assert parts[..., 3 * SL:].abs().max() == 0
    w_head = ops.split_weights_head(pd["ff"])
    leak, thresh = pd["leak"].reshape(-1), pd["thresh"].reshape(-1)
    v_in, z_in = st[0].to(DEV).contiguous(), ops.pack_cl(st[1].to(DEV))
    v_tc, z_tc = ops.lif_step_cl(x_cl, v_in, z_in, pd["ff"], None, leak, thresh, hard_reset=True, w_split=w_head)
    v_cc, z_cc = ops.lif_step_cl(None, v_in, z_in, pd["ff"], None, leak, thresh, hard_reset=True, x_f32=x.to(DEV))
    thr = params["thresh"].clamp_min(0.01)
    z_tc_f = ops.unpack_cl(z_tc).cpu()
    spike_band_compare(v_tc.cpu(), z_tc_f, v_cc.cpu(), ops.unpack_cl(z_cc).cpu(), thr)
    xo = x.clone()
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out_o, ns_o = osp.cell_step("lif", xo, st, po, hard_reset=True)
    spike_band_compare(v_tc.cpu(), z_tc_f, ns_o[0].detach(), ns_o[1].detach(), thr)
    assert z_tc_f.mean() > 0.02
    # weight gradient of the head on the tensor cores (one-step window)
    g_out = torch.rand((B, 32, H, W), generator=g)
    (out_o * g_out).sum().backward()
    q = L.LifBwdWindowParams()
    q.B, q.T, q.H, q.W, q.hard_reset, q.surrogate, q.act_width = B, 1, H, W, 1, 0, 10.0
    g_w = torch.zeros_like(pd["ff"])
    g_leak, g_thresh = torch.zeros(32, device=DEV), torch.zeros(32, device=DEV)
    gI_hi = torch.empty((B, H, W, 32), device=DEV, dtype=torch.bfloat16)
    gI_mid = torch.empty_like(gI_hi)
    wg = torch.empty(L.lib().ef_lif_wgrad_partial_elems(B, H, W, 0), device=DEV)
    god = g_out.to(DEV)
    q.Cin, q.x_cl, q.z_prev_cl, q.v, q.v_prev, q.g_out = cin, L.ptr(x_cl), L.ptr(z_in), L.ptr(v_tc), L.ptr(v_in), L.ptr(god)
    q.leak, q.thresh, q.gI_hi, q.gI_mid, q.wg_partial = L.ptr(leak), L.ptr(thresh), L.ptr(gI_hi), L.ptr(gI_mid), L.ptr(wg)
    q.g_w_ff, q.g_leak, q.g_thresh = L.ptr(g_w), L.ptr(g_leak), L.ptr(g_thresh)
    L.call("ef_lif_bwd_window", q)
    from tests.util import assert_rel

    assert_rel(g_w, po["ff"].grad, 1e-3, "head g_w_ff (tensor cores)")
    assert_rel(g_leak, po["leak"].grad.reshape(-1), 1e-3, "head g_leak")
    assert_rel(g_thresh, po["thresh"].grad.reshape(-1), 1e-3, "head g_thresh")


# ---- 32-channel cells of every neuron kind: convolution on the tensor cores + neuron update on its current (ops._CellStep, x_kind) -------------
def _cell_case(neuron, layer, B, H, W, seed, with_state, bins=5):
    g = torch.Generator().manual_seed(seed)
    params = osp.init_firenet_params(neuron, bins, 32, seed=seed, weight_gain=2.0)[layer]
    if layer == "head":
        x = torch.randn((B, bins, H, W), generator=g) * (torch.rand((B, bins, H, W), generator=g) < 0.4).float()  # fractional voxel-like input
    else:
        x = (torch.rand((B, 32, H, W), generator=g) < 0.3).float()
    st = None
    if with_state:
        n = 2 if neuron == "lif" else 3
        st = torch.rand((n, B, 32, H, W), generator=g) * 1.2 - 0.1
        st[1] = (st[1] < 0.3).float()
        if n == 3:
            st[2] = st[2].abs() * 0.3
    return params, x, st


@pytest.mark.parametrize("neuron", ["lif", "plif", "alif", "xlif"])
@pytest.mark.parametrize("layer", ["head", "R1a", "G1"])
@pytest.mark.parametrize("hard", [True, False])
@pytest.mark.parametrize("with_state", [True, False])
def test_tensor_core_conv_plus_neuron_update_matches_generic_kernel_and_oracle(neuron, layer, hard, with_state):
    """
    ops.cell_step(x_kind="spikes" | "split") = tcgen05 convolution (exact products) + ef_lif_neuron_fwd, against the fused CUDA-core kernel
    on the same tensors and against the oracle: membrane / trace to summation-order noise, spikes exact outside the band; then the
    backward (same kernel for both, fed by each path's saved state) to 1e-4.
    """
    from event_flow_b200 import ops

    B, H, W = 2, 36, 44
    params, x, st = _cell_case(neuron, layer, B, H, W, seed=11, with_state=with_state)
    pd = {k: v.to(DEV).contiguous() for k, v in params.items()}
    chan = {n: pd[n] for n in ops.param_names(neuron)}
    kind = "split" if layer == "head" else "spikes"
    res = {}
    for tag, xk in (("tc", kind), ("cc", None)):
        xd = x.to(DEV).requires_grad_(True)
        sd = None if st is None else st.to(DEV).requires_grad_(True)
        ws = {k: pd[k].clone().requires_grad_(True) for k in ("ff", "rec") if k in pd}
        out, ns = ops.cell_step(neuron, xd, sd, ws["ff"], ws.get("rec"), chan, hard_reset=hard, x_kind=xk)
        gen = torch.Generator().manual_seed(5)
        g_out, g_ns = torch.randn(out.shape, generator=gen).to(DEV), torch.randn(ns.shape, generator=gen).to(DEV)
        g_ns[1] = 0  # (the spikes inside the state are the same tensor values as `out`)
        torch.autograd.backward([out, ns], [g_out, g_ns])
        res[tag] = (out.detach().cpu(), ns.detach().cpu(), xd.grad.cpu(), None if sd is None else sd.grad.cpu(), {k: w.grad.cpu() for k, w in ws.items()})
    out_o, ns_o = osp.cell_step(neuron, x, st, params, hard_reset=hard)
    if neuron in ("lif", "plif"):
        thr = params["thresh"].clamp_min(0.01)
    else:
        thr = params["t0"].clamp_min(0.01) + params["t1"].clamp_min(0) * ns_o[2]
    for other_v, other_z, other_aux in ((res["cc"][1][0], res["cc"][1][1], res["cc"][1][2] if neuron != "lif" else None),
                                        (ns_o[0], ns_o[1], ns_o[2] if neuron != "lif" else None)):
        _, _, in_flips = spike_band_compare(res["tc"][1][0], res["tc"][1][1], other_v, other_z, thr)
        if other_aux is not None:
            assert (res["tc"][1][2] - other_aux).abs().max().item() <= 1e-6 * max(1.0, other_aux.abs().max().item())
    assert torch.equal(res["tc"][0], res["tc"][1][1])  # out = spikes (no residual)
    # The backward of ONE cell step reads the inputs, the new membrane potential / trace and the parameters -- never the emitted spikes --
    # so the gradients of the two paths are comparable UNCONDITIONALLY (a borderline spike that flipped changes nothing they read).
    for a, b, what in ((res["tc"][2], res["cc"][2], "g_x"), (res["tc"][3], res["cc"][3], "g_state")) + tuple(
            (res["tc"][4][k], res["cc"][4][k], "g_" + k) for k in res["cc"][4]):
        if b is not None:
            assert (a - b).abs().max().item() <= 1e-4 * (b.abs().max().item() + 1e-20), what
    assert res["tc"][1][1].mean() > 0.01


def test_neuron_step_c_abi_rejects_bad_arguments():
    from event_flow_b200 import _lib as L

    assert L.lib().ef_lif_neuron_fwd(None, None, None) < 0
    p = L.LifConvParams()
    assert L.lib().ef_lif_neuron_fwd(p, None, None) < 0
