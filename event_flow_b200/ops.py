"""
Host-side operators over libeventflow.so: tensor allocation, argument marshalling and torch.autograd glue.
Every function here launches CUDA kernels through the C ABI (event_flow_b200/_lib.py); none has a CPU path.
"""

import ctypes as C
import weakref

import torch

from . import _lib as L

_N_STATE = {"lif": 2, "plif": 3, "alif": 3, "xlif": 3}
# per-channel parameter names of each neuron kind, in the order of the C struct fields
#                 leak      thresh    leak_aux   add_pt    t0    t1
_PARAM_FIELDS = {
    "lif": ("leak", "thresh", None, None, None, None),
    "plif": ("leak_v", "thresh", "leak_pt", "add_pt", None, None),
    "alif": ("leak_v", None, "leak_t", None, "t0", "t1"),
    "xlif": ("leak_v", None, "leak_pt", None, "t0", "t1"),
}
_STRUCT_FIELDS = ("leak", "thresh", "leak_aux", "add_pt", "t0", "t1")


def param_names(neuron):
    """Names of the per-channel parameters of a neuron kind (reference attribute names)."""
    return tuple(n for n in _PARAM_FIELDS[neuron] if n is not None)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.EventFlowError("event_flow_b200 has no CPU path: tensors must live on a CUDA device (got %s)" % t.device)


def _c(t):
    return None if t is None else t.contiguous()


def _need_f32(*ts):
    for t in ts:
        if t is not None and t.dtype != torch.float32:
            raise L.EventFlowError("event_flow_b200 kernels read fp32 tensors; got %s (cast with .float() first)" % t.dtype)


def _fill_cell_params(p, neuron, x, state_in, w_ff, w_rec, chan, residual, state_out, out, hard_reset, surrogate, width, stride):
    B, Cin, H, W = x.shape
    Cout = w_ff.shape[0]
    p.B, p.Cin, p.C, p.H, p.W = B, Cin, Cout, H, W
    p.ksize, p.stride = w_ff.shape[-1], stride
    p.neuron, p.hard_reset = L.NEURON_CODES[neuron], int(bool(hard_reset))
    p.surrogate, p.act_width = L.SURROGATE_CODES[surrogate], float(width)
    p.x = L.ptr(x)
    if state_in is not None:
        planes = L.planes(state_in)
        p.v_in, p.z_in = planes[0], planes[1]
        if len(planes) > 2:
            p.aux_in = planes[2]
    p.w_ff, p.w_rec = L.ptr(w_ff), L.ptr(w_rec)
    for field, name in zip(_STRUCT_FIELDS, _PARAM_FIELDS[neuron]):
        if name is not None:
            setattr(p, field, L.ptr(chan[name]))
    p.residual = L.ptr(residual)
    planes = L.planes(state_out)
    p.v_out, p.z_out = planes[0], planes[1]
    if len(planes) > 2:
        p.aux_out = planes[2]
    p.out = L.ptr(out)


TCG_BACKWARD = True  # data gradients of cells with C % 32 == 0 outputs on the general tensor-core kernel (tests switch it off to compare)


def _tcg_dgrad_ok(neuron, x, w_ff, stride, wanted):
    """The convolution data gradients of this cell step can run as plain convolutions on the general tcgen05 kernel (_tcg_dgrad)."""
    Cout, Cin = w_ff.shape[0], w_ff.shape[1]
    H, W = x.shape[2], x.shape[3]
    return (TCG_BACKWARD and wanted and x.is_cuda and neuron in ("lif", "alif") and w_ff.shape[-1] == 3 and Cout % 32 == 0 and Cout >= 32
            and Cin >= 16 and stride in (1, 2) and W % 4 == 0 and ((W - 1) // stride + 1) % 4 == 0 and (stride == 1 or (H % 2 == 0 and W % 2 == 0)))


def _tcg_consts(dev, C):
    """Per-channel constants that turn the fused LIF kernel into a plain convolution: leak = -inf (sigmoid = 0), threshold 1."""
    key = (dev, C)
    hit = _TC_CONSTS.get(key)
    if hit is None:
        hit = _TC_CONSTS[key] = (torch.full((C,), float("-inf"), device=dev), torch.ones(C, device=dev))
    return hit


def _dgrad_image(w, rec_w=None):
    """
    Weight image of the DATA GRADIENT of a 3x3 convolution run as a convolution of g_I = hi + mid on the general tensor-core kernel:
    the flipped, transposed kernel (rows padded to a multiple of 32 input channels; with `rec_w` the recurrent convolution's gradient
    rides along as further output channels), once for each of the two g_I sources.  Cached per weight version like the forward images.
    Returns (image, padded channel count of the feed-forward part, total output channels).
    """
    key = (w.data_ptr(), w._version, None if rec_w is None else (rec_w.data_ptr(), rec_w._version), WEIGHT_EPOCH, "dgrad")
    slot = (id(w), "dgrad", rec_w is not None)
    hit = _TC_IMAGES.get(slot)
    if hit is None or hit[0] != key or hit[2]() is not w or (rec_w is not None and hit[3]() is not rec_w):
        Cout, Cin = w.shape[0], w.shape[1]
        cpad = (Cin + 31) // 32 * 32
        wd = torch.zeros((cpad, Cout, 3, 3), device=w.device, dtype=torch.float32)
        wd[:Cin] = w.detach().flip(2, 3).transpose(0, 1)
        if rec_w is not None:
            wd = torch.cat([wd, rec_w.detach().flip(2, 3).transpose(0, 1)], 0)
        wd = wd.contiguous()
        image = split_weights_g([(wd, 0, Cout, False), (wd, 0, Cout, False)], wd.shape[0])
        if len(_TC_IMAGES) > 128:
            _TC_IMAGES.clear()
        hit = _TC_IMAGES[slot] = (key, (image, cpad, wd.shape[0]), weakref.ref(w), None if rec_w is None else weakref.ref(rec_w))
    return hit[1]


def _grad_terms(g, H, W):
    """g fp32 NCHW [B,C,Hs,Ws] as two bf16 channels-last terms [2,B,H,W,C] (16 significant bits), zero-inserted to (H, W) when g is the
    output gradient of a stride-2 convolution: the operand of the tensor-core data- and weight-gradient kernels."""
    B, Cg, Hs, Ws = g.shape
    terms = torch.empty((2, B, H, W, Cg), device=g.device, dtype=torch.bfloat16)
    hi, mid = L.planes(terms)
    L.LAUNCHES += 1
    L.check(L.lib().ef_split2_pack_cl(L.ptr(g), hi, mid, B, Cg, H, W, Hs, Ws, L.stream()), "ef_split2_pack_cl")
    return terms


def _tcg_conv_of_grad(terms, image, c_total):
    """conv3x3(g, flipped kernel) on the general tensor-core kernel; returns fp32 NCHW [B,c_total,H,W]."""
    neg_inf, ones = _tcg_consts(terms.device, c_total)
    v, _, _ = lif_step_g([terms[0], terms[1]], None, None, image, neg_inf, ones, c_total)
    return v


def _tcg_dgrad(terms, w_ff, w_rec, want_x, want_z):
    """Data gradients of a cell step's convolutions: (g_x [B,Cin,H,W] | None, recurrent part of g_z_in [B,C,Ho,Wo] | None).  A recurrent
    cell has stride 1: `terms` serves both convolutions."""
    Cin = w_ff.shape[1]
    g_x = g_z = None
    if want_x and want_z:  # one launch: the recurrent gradient as extra output channels
        image, cpad, ctot = _dgrad_image(w_ff, w_rec)
        v = _tcg_conv_of_grad(terms, image, ctot)
        return v[:, :Cin].contiguous(), v[:, cpad:].contiguous()
    if want_x:
        image, cpad, ctot = _dgrad_image(w_ff)
        v = _tcg_conv_of_grad(terms, image, ctot)
        g_x = v if cpad == Cin else v[:, :Cin].contiguous()
    if want_z:
        image, cpad, ctot = _dgrad_image(w_rec)
        g_z = _tcg_conv_of_grad(terms, image, ctot)
    return g_x, g_z


def _tcg_wgrad(x_cl, terms, g_w, ci_off):
    """g_w[:, ci_off : ci_off + cin] += weight gradient of a 3x3 convolution on the tensor cores (ef_wgrad_tcg); x_cl [B,H,W,cin] exact in bf16."""
    _, B, H, W, cout = terms.shape
    cin = x_cl.shape[3]
    partial = torch.empty(L.lib().ef_wgrad_tcg_partial_elems(B, H, W, cin, cout), device=x_cl.device, dtype=torch.float32)
    hi, mid = L.planes(terms)
    L.LAUNCHES += 2
    L.check(L.lib().ef_wgrad_tcg(L.ptr(x_cl), hi, mid, B, H, W, cin, cout, L.ptr(partial), L.ptr(g_w), g_w.shape[1], int(ci_off), L.stream()),
            "ef_wgrad_tcg")


TCG_FORWARD = True  # forward of LIF cells with C % 32 == 0 outputs under autograd on the general tensor-core kernel (tests switch it off)


def _tcg_fwd_ok(neuron, x, state_in, w_ff, w_rec, stride, x_kind, residual):
    """
    This cell step can run on the general tcgen05 kernel (ef_lif_conv_fwd_g; LIF: fused, other neurons: its convolution) with fp32-exact products: the caller vouches for the
    input -- "spikes": every channel exact in bf16 (spikes, sums of spikes, their bilinear x2 upsampling); ("mixed", n): the first
    n <= 10 channels are arbitrary fp32 values (they enter as their exact three-way split), the rest exact in bf16.
    """
    if not (TCG_FORWARD and x.is_cuda and w_ff.shape[-1] == 3 and w_ff.shape[0] % 32 == 0):
        return False
    if neuron != "lif" and stride != 1:  # (the other neuron kinds: convolution on this kernel + ef_lif_neuron_fwd, stride 1 only)
        return False
    Cin, H, W = x.shape[1], x.shape[2], x.shape[3]
    if x_kind == "spikes":
        if Cin % 32 != 0:
            return False
        if stride == 2:
            return w_rec is None and H % 2 == 0 and W % 2 == 0 and (W // 2) % 4 == 0
        return stride == 1 and W % 4 == 0
    if type(x_kind) is tuple and x_kind[0] == "mixed":
        n = x_kind[1]
        return stride == 1 and W % 4 == 0 and 0 < n <= L.EF_HEAD_MAX_CIN and (Cin - n) % 32 == 0 and Cin > n and w_rec is None
    return False


def _handed(t, name):
    """The channels-last companion a producing cell attached to `t`, if `t` was not modified since (torch's version counter)."""
    hit = getattr(t, name, None)
    return hit[0] if hit is not None and hit[1] == t._version else None


def _fwdg_image(w_ff, w_rec, wsrcs, tag):
    """Cached ef_split_weights_g image of a cell's weights for one source list (keyed like the other weight images)."""
    key = (w_ff.data_ptr(), w_ff._version, None if w_rec is None else (w_rec.data_ptr(), w_rec._version), WEIGHT_EPOCH, tag)
    slot = (id(w_ff), "fwdg", tag)
    hit = _TC_IMAGES.get(slot)
    if hit is None or hit[0] != key or hit[2]() is not w_ff or (w_rec is not None and hit[3]() is not w_rec):
        image = split_weights_g(wsrcs, w_ff.shape[0])
        if len(_TC_IMAGES) > 128:
            _TC_IMAGES.clear()
        hit = _TC_IMAGES[slot] = (key, image, weakref.ref(w_ff), None if w_rec is None else weakref.ref(w_rec))
    return hit[1]


def _tcg_forward(x, state_in, w_ff, w_rec, leak, thresh, residual, state_out, out, hard_reset, stride, x_kind, neuron_call=None):
    """
    The cell step on ef_lif_conv_fwd_g, results written into `state_out` and `out`; returns (out_cl, z_cl, operands for the backward).
    LIF: the fused kernel.  Other neurons (`neuron_call` = their filled ef_lif_conv_params): the kernel as a pure convolution (leak =
    -inf), then ef_lif_neuron_fwd on its current -- what the 32-channel cells do with the 32-channel kernel.
    """
    B, Cin, H, W = x.shape
    C = w_ff.shape[0]
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    s2d = stride == 2
    if x_kind == "spikes":
        x_cl = _handed(x, "_ef_cl")
        if x_cl is None or x_cl.shape != (B, H, W, Cin):
            x_cl = pack_cl(x)
        srcs, wsrcs, tag, n = [space_to_depth_cl(x_cl) if s2d else x_cl], [(w_ff, 0, Cin, False, s2d)], ("spikes", s2d), 0
    else:
        n = x_kind[1]
        x_cl = pack_cl(x[:, n:].contiguous())
        srcs = [pack_split_cl(x[:, :n].contiguous()), x_cl]
        wsrcs, tag = [(w_ff, 0, n, True), (w_ff, n, Cin - n, False)], ("mixed", n)
    z_in_cl = v_in = None
    if state_in is not None and (neuron_call is None or w_rec is not None):
        planes = L.planes(state_in)
        v_in = planes[0]
        z_in_cl = _handed(state_in, "_ef_z_cl")
        if z_in_cl is None:
            z_in_cl = pack_cl(state_in[1])
        if w_rec is not None:
            srcs, wsrcs, tag = srcs + [z_in_cl], wsrcs + [(w_rec, 0, C, False)], tag + ("rec",)
    image = _fwdg_image(w_ff, w_rec if tag[-1] == "rec" else None, wsrcs, tag)
    operands = (x_cl, n, z_in_cl if tag[-1] == "rec" else None)
    if neuron_call is not None:
        neg_inf, ones = _tcg_consts(x.device, C)
        cur, _, _ = lif_step_g(srcs, None, None, image, neg_inf, ones, C)
        q = neuron_call
        out_cl = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.bfloat16)
        q.out_cl = L.ptr(out_cl)
        z_cl = out_cl
        if residual is not None:
            z_cl = torch.empty_like(out_cl)
            q.z_out_cl = L.ptr(z_cl)
        L.LAUNCHES += 1
        L.check(L.lib().ef_lif_neuron_fwd(L.C.byref(q), L.ptr(cur), L.stream()), "ef_lif_neuron_fwd")
        return out_cl, z_cl, operands
    res_cl = None
    if residual is not None:
        res_cl = _handed(residual, "_ef_cl")
        if res_cl is None:
            res_cl = pack_cl(residual)
    z_cl = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.bfloat16)
    out_cl = torch.empty_like(z_cl) if res_cl is not None else None
    p = L.LifConvGParams()
    p.B, p.H, p.W, p.C, p.n_src, p.hard_reset, p.s2d = B, Ho, Wo, C, len(srcs), int(bool(hard_reset)), int(s2d)
    for i, t in enumerate(srcs):
        p.src[i], p.src_c[i] = L.ptr(t), t.shape[3]
    p.v_in, p.z_in_cl, p.residual_cl = v_in, L.ptr(z_in_cl), L.ptr(res_cl)
    p.leak, p.thresh, p.w_image = L.ptr(leak), L.ptr(thresh), L.ptr(image)
    so = L.planes(state_out)
    p.v_out, p.z_out_cl, p.out_cl = so[0], L.ptr(z_cl), L.ptr(out_cl)
    L.call("ef_lif_conv_fwd_g", p)
    L.LAUNCHES += 1
    L.check(L.lib().ef_unpack_cl(L.ptr(z_cl), so[1], B, C, Ho, Wo, L.stream()), "ef_unpack_cl")
    if out_cl is not None:
        L.LAUNCHES += 1
        L.check(L.lib().ef_unpack_cl(L.ptr(out_cl), L.ptr(out), B, C, Ho, Wo, L.stream()), "ef_unpack_cl")
    else:
        out.copy_(state_out[1])
    # operands = (x_cl, n, z_in_cl): the bf16 tensors the weight gradients can run on -- the exact channels of the input (from channel n
    # on) and the previous spikes of a recurrent cell
    return (out_cl if out_cl is not None else z_cl), z_cl, operands


class _CellStep(torch.autograd.Function):
    """One fused conv + neuron step on fp32 NCHW tensors (the reference cells' own tensor contract)."""

    @staticmethod
    def forward(ctx, meta, x, state_in, w_ff, w_rec, residual, *chan_vals):
        neuron, hard_reset, surrogate, width, stride, x_kind, _ = meta
        names = param_names(neuron)
        chan = {n: _c(v) for n, v in zip(names, chan_vals)}  # [C,1,1] (or [C]) contiguous: the kernels read C values at the pointer
        x, state_in, w_ff, w_rec, residual = _c(x), _c(state_in), _c(w_ff), _c(w_rec), _c(residual)
        _need_cuda(x, state_in, w_ff, w_rec, residual, *chan.values())
        B, _, H, W = x.shape
        Cout = w_ff.shape[0]
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        state_out = torch.empty((_N_STATE[neuron], B, Cout, Ho, Wo), device=x.device, dtype=torch.float32)
        out = torch.empty((B, Cout, Ho, Wo), device=x.device, dtype=torch.float32)  # z (+ residual); a tensor of its own
        p = L.LifConvParams()
        _fill_cell_params(p, neuron, x, state_in, w_ff, w_rec, chan, residual, state_out, out, hard_reset, surrogate, width, stride)
        if _tc_conv_ok(x, w_ff, w_rec, stride, x_kind):
            # 32-channel cell of ANY neuron kind: the convolution on the tensor cores (exact products), the neuron update on its current
            cur, x_cl, z_in_cl = _tc_conv_current(x, state_in, w_ff, w_rec, x_kind)
            if x_kind == "spikes":  # the backward runs its convolution gradients on the tensor cores from the same operands
                ctx.tc_operands = (x_cl, z_in_cl)
            # the same spikes also in the internal format, for the next cell's / next step's tensor-core convolution (no re-packing)
            out_cl = torch.empty((B, Ho, Wo, Cout), device=x.device, dtype=torch.bfloat16)
            p.out_cl = L.ptr(out_cl)
            z_cl = out_cl
            if residual is not None:
                z_cl = torch.empty_like(out_cl)
                p.z_out_cl = L.ptr(z_cl)
            L.LAUNCHES += 1
            L.check(L.lib().ef_lif_neuron_fwd(C.byref(p), L.ptr(cur), L.stream()), "ef_lif_neuron_fwd")
            out._ef_cl, state_out._ef_z_cl = (out_cl, out._version), (z_cl, state_out._version)
        elif _tcg_fwd_ok(neuron, x, state_in, w_ff, w_rec, stride, x_kind, residual):
            # LIF cell of the U-Net family (other channel counts, stride 2, mixed inputs): the fused general tensor-core kernel
            out_cl, z_cl, ctx.tcg_operands = _tcg_forward(x, state_in, w_ff, w_rec, chan.get("leak"), chan.get("thresh"), residual, state_out, out,
                                                          hard_reset, stride, x_kind, neuron_call=None if neuron == "lif" else p)
            out._ef_cl, state_out._ef_z_cl = (out_cl, out._version), (z_cl, state_out._version)
        else:
            L.call("ef_lif_conv_fwd", p, tag=(x.shape[1], Cout, w_rec is not None))
        ctx.meta = meta
        ctx.params = p  # the backward starts from a copy of this argument block (same tensors: all of them are saved below)
        ctx.chan_names = names
        if not hasattr(ctx, "tc_operands"):
            ctx.tc_operands = None
        if not hasattr(ctx, "tcg_operands"):
            ctx.tcg_operands = None
        ctx.save_for_backward(x, state_in, w_ff, w_rec, residual, state_out, *[chan[n] for n in names])
        ctx.chan_shapes = [v.shape for v in chan_vals]
        ctx.has_residual = residual is not None
        return out, state_out

    @staticmethod
    def backward(ctx, g_out, g_state):
        neuron, hard_reset, surrogate, width, stride, _, detach = ctx.meta
        x, state_in, w_ff, w_rec, residual, state_out, *chan_vals = ctx.saved_tensors
        names = ctx.chan_names
        chan = dict(zip(names, chan_vals))
        q = L.LifConvBwdParams()
        q.f = ctx.params  # (struct copy; the forward's output / residual pointers in it are never read by the backward kernels)
        if q.f.x != x.data_ptr() or q.f.v_out != state_out.data_ptr() or q.f.w_ff != w_ff.data_ptr():
            # the saved tensors came back in other storage (saved-tensor hooks: offloading, ...): rebuild the argument block
            q.f = L.LifConvParams()
            _fill_cell_params(q.f, neuron, x, state_in, w_ff, w_rec, chan, None, state_out, None, hard_reset, surrogate, width, stride)
        B, Cin, H, W = x.shape
        S, _, Cout, Ho, Wo = state_out.shape
        dev = x.device
        g_out = _c(g_out)
        g_state = _c(g_state)
        q.g_out = L.ptr(g_out)
        if g_state is not None:
            planes = L.planes(g_state)
            q.g_v_out, q.g_z_out = planes[0], planes[1]
            if S > 2:
                q.g_aux_out = planes[2]
        scratch = torch.empty((B, Cout, Ho, Wo), device=dev, dtype=torch.float32)
        q.scratch_gI = L.ptr(scratch)
        need = ctx.needs_input_grad  # (meta, x, state_in, w_ff, w_rec, residual, *chan)
        g_x = torch.empty_like(x) if need[1] else None
        q.g_x = L.ptr(g_x)
        scratch_p = scratch_up = scratch_p_up = None
        if g_x is not None and neuron in ("plif", "xlif"):
            scratch_p = torch.empty((B, Ho, Wo), device=dev, dtype=torch.float32)
            q.scratch_gP = L.ptr(scratch_p)
        if stride == 2:  # the stride-1 gradient kernels run on the zero-inserted output gradient (csrc/lif_conv_bwd.cu)
            scratch_up = torch.empty((B, Cout, H, W), device=dev, dtype=torch.float32)
            q.scratch_gI_up = L.ptr(scratch_up)
            if scratch_p is not None:
                scratch_p_up = torch.empty((B, H, W), device=dev, dtype=torch.float32)
                q.scratch_gP_up = L.ptr(scratch_p_up)
        g_state_in = None
        if state_in is not None and need[2]:
            g_state_in = torch.empty_like(state_in)
            planes = L.planes(g_state_in)
            q.g_v_in, q.g_z_in = planes[0], planes[1]
            if S > 2:
                q.g_aux_in = planes[2]
        # every parameter gradient of this step out of ONE zero-filled buffer (one fill launch instead of one per tensor)
        fields = [f for f, n in zip(_STRUCT_FIELDS, _PARAM_FIELDS[neuron]) if n is not None]
        n_ff = w_ff.numel() if need[3] else 0
        n_rec = w_rec.numel() if (w_rec is not None and need[4]) else 0
        n_chan = sum(1 for i in range(len(fields)) if need[6 + i])
        flat = torch.zeros(n_ff + n_rec + n_chan * Cout, device=dev, dtype=torch.float32)
        g_w_ff = flat[:n_ff].view(w_ff.shape) if n_ff else None
        q.g_w_ff = L.ptr(g_w_ff)
        g_w_rec = flat[n_ff:n_ff + n_rec].view(w_rec.shape) if n_rec else None
        q.g_w_rec = L.ptr(g_w_rec)
        g_chan, o = [], n_ff + n_rec
        for i, field in enumerate(fields):
            if need[6 + i]:
                g = flat[o:o + Cout]
                o += Cout
                setattr(q, "g_" + field, L.ptr(g))
                g_chan.append(g.view(ctx.chan_shapes[i]))
            else:
                g_chan.append(None)
        q.reset_grad = int(not detach and g_state_in is not None)
        if ctx.tc_operands is not None and (need[1] or need[3] or need[4] or g_state_in is not None):
            # 32 -> 32 cell on spike inputs: neuron backward here, the convolution gradients on the tensor cores (any neuron kind)
            x_cl, z_in_cl = ctx.tc_operands
            if scratch_p is None and neuron in ("plif", "xlif"):
                scratch_p = torch.empty((B, Ho, Wo), device=dev, dtype=torch.float32)
                q.scratch_gP = L.ptr(scratch_p)
            q.neuron_only = 1
            L.call("ef_lif_conv_bwd", q)
            rec = w_rec is not None
            t = L.Conv32BwdTcParams()
            t.B, t.H, t.W, t.has_rec = B, H, W, int(rec)
            t.gI, t.x_cl, t.z_in_cl = L.ptr(scratch), L.ptr(x_cl), L.ptr(z_in_cl)
            t.w_bwd = L.ptr(_weight_image(w_ff, w_rec, "bwd"))
            gI_split = torch.empty((2, B, H, W, 32), device=dev, dtype=torch.bfloat16)
            t.gI_hi, t.gI_mid = L.planes(gI_split)
            g_x_tc = g_x if g_x is not None else torch.empty_like(x)
            t.g_x = L.ptr(g_x_tc)
            g_z_tmp = None
            if rec and z_in_cl is not None and g_state_in is not None:
                g_z_tmp = torch.empty((B, 32, H, W), device=dev, dtype=torch.float32)
                t.g_z_in, t.g_z_tmp = q.g_z_in, L.ptr(g_z_tmp)
            partial = None
            if g_w_ff is not None or g_w_rec is not None:
                partial = torch.empty(L.lib().ef_lif_wgrad_partial_elems(B, H, W, int(rec)), device=dev, dtype=torch.float32)
                t.wg_partial, t.wg_flags = L.ptr(partial), L.EF_WG_FINALIZE
                t.g_w_ff, t.g_w_rec = L.ptr(g_w_ff), L.ptr(g_w_rec)
            if scratch_p is not None:
                t.gP_sum, t.x_f32 = L.ptr(scratch_p), L.ptr(x)
            L.call("ef_conv32_bwd_tc", t)
        elif _tcg_dgrad_ok(neuron, x, w_ff, stride, g_x is not None or (w_rec is not None and g_state_in is not None)):
            # other channel counts (the U-Net family): neuron backward, data gradients as plain convolutions of g_I on the general
            # tensor-core kernel, weight gradients on the CUDA cores
            q.neuron_only = 1
            L.call("ef_lif_conv_bwd", q)
            want_z = w_rec is not None and g_state_in is not None
            terms = _grad_terms(scratch, H, W)  # at the input resolution (zero-inserted for stride 2); a recurrent cell has stride 1
            tc_x, tc_z = _tcg_dgrad(terms, w_ff, w_rec if want_z else None, g_x is not None, want_z)
            if tc_x is not None:
                g_x = tc_x
            if tc_z is not None:
                g_state_in[1] += tc_z
            # weight gradients: on the tensor cores from the bf16 operands the forward ran on, else on the CUDA cores
            x_exact_cl, n_real, z_in_cl = ctx.tcg_operands if ctx.tcg_operands is not None else (None, 0, None)
            if g_w_ff is not None:
                if x_exact_cl is not None:
                    _tcg_wgrad(x_exact_cl, terms, g_w_ff, n_real)
                    if n_real:  # the fractional channels of a mixed input: a narrow convolution of their own on the CUDA cores
                        g_small = torch.zeros((Cout, n_real, 3, 3), device=dev, dtype=torch.float32)
                        L.LAUNCHES += 1
                        L.check(L.lib().ef_conv3x3_bwd(L.ptr(scratch), L.ptr(x[:, :n_real].contiguous()), L.ptr(w_ff), None, L.ptr(g_small), B, n_real, Cout,
                                                       H, W, L.stream()), "ef_conv3x3_bwd")
                        g_w_ff[:, :n_real] = g_small
                else:
                    L.LAUNCHES += 1
                    L.check(L.lib().ef_conv3x3_bwd_s(L.ptr(scratch), L.ptr(x), L.ptr(w_ff), None, L.ptr(g_w_ff), L.ptr(scratch_up), B, Cin, Cout, H, W,
                                                     int(stride), L.stream()), "ef_conv3x3_bwd_s")
            if g_w_rec is not None and state_in is not None:
                if z_in_cl is not None:
                    _tcg_wgrad(z_in_cl, terms, g_w_rec, 0)
                else:
                    L.LAUNCHES += 1
                    L.check(L.lib().ef_conv3x3_bwd(L.ptr(scratch), L.planes(state_in)[1], L.ptr(w_rec), None, L.ptr(g_w_rec), B, Cout, Cout, Ho, Wo,
                                                   L.stream()), "ef_conv3x3_bwd")
        else:
            L.call("ef_lif_conv_bwd", q)
        g_res = g_out if (ctx.has_residual and need[5]) else None
        return (None, g_x, g_state_in, g_w_ff, g_w_rec, g_res, *g_chan)


_TC_CONSTS = {}


def _tc_conv_ok(x, w_ff, w_rec, stride, x_kind):
    """The convolution of this cell step can run on the tcgen05 kernel: 32 output channels, 3x3, stride 1, inputs the caller vouches to be
    exactly representable in bf16 ("spikes": 32 channels of spikes / residual sums) or few fractional channels ("split": exact three-way
    bf16 split, feed-forward cells only)."""
    if x_kind is None or stride != 1 or w_ff.shape[0] != 32 or w_ff.shape[-1] != 3 or x.shape[-1] % 4 != 0 or not x.is_cuda:
        return False
    if x_kind == "spikes":
        return x.shape[1] == 32
    return x_kind == "split" and x.shape[1] <= L.EF_HEAD_MAX_CIN and w_rec is None


def _weight_image(w_ff, w_rec, kind):
    """
    bf16 operand image of a cell's weights for the tensor-core kernels -- kind "spikes" / "split": forward (ef_split_weights /
    ef_split_weights_head), "bwd": data gradient (ef_split_weights_bwd).  Rebuilt when the weight tensors changed (torch's version
    counters; WEIGHT_EPOCH for in-place updates behind them); keyed on the identity of the weight tensor OBJECTS, held weakly: a freed
    tensor's address may be reused by other values.
    """
    key = (w_ff.data_ptr(), w_ff._version, None if w_rec is None else (w_rec.data_ptr(), w_rec._version), WEIGHT_EPOCH, kind)
    slot = (id(w_ff), kind == "bwd")
    hit = _TC_IMAGES.get(slot)
    if hit is None or hit[0] != key or hit[2]() is not w_ff or (w_rec is not None and hit[3]() is not w_rec):
        if kind == "bwd":
            image = split_weights_bwd(w_ff, w_rec)
        else:
            image = split_weights_head(w_ff) if kind == "split" else split_weights(w_ff, w_rec)
        if len(_TC_IMAGES) > 128:
            _TC_IMAGES.clear()
        hit = _TC_IMAGES[slot] = (key, image, weakref.ref(w_ff), None if w_rec is None else weakref.ref(w_rec))
    return hit[1]


def _tc_conv_current(x, state_in, w_ff, w_rec, x_kind):
    """cur = conv(x, w_ff) (+ conv(z_in, w_rec)) [B,32,H,W] fp32 on the tensor cores: the fused LIF kernel run as a pure convolution --
    leak = -inf makes sigmoid(leak) = 0, so its membrane output is (1 - 0) * current exactly, whatever state it is given."""
    dev = x.device
    consts = _TC_CONSTS.get(dev)
    if consts is None:
        consts = _TC_CONSTS[dev] = (torch.full((32,), float("-inf"), device=dev), torch.ones(32, device=dev))
    neg_inf, ones = consts
    image = _weight_image(w_ff, w_rec, x_kind)
    # input / previous spikes in the internal format: handed over by the cell that produced them (attributes of the tensors), else packed here
    # (a hand-over is only trusted while the fp32 tensor it mirrors is unmodified: torch's version counter)
    def handed(t, name):
        hit = getattr(t, name, None)
        return hit[0] if hit is not None and hit[1] == t._version else None

    x_cl = handed(x, "_ef_cl")
    if x_kind == "split":
        x_cl = pack_split_cl(x)
    elif x_cl is None or x_cl.shape[:3] != (x.shape[0], x.shape[2], x.shape[3]):
        x_cl = pack_cl(x)
    v_in = z_cl = None
    if w_rec is not None and state_in is not None:
        v_in = state_in[0]  # (only has to be finite: it is multiplied by sigmoid(-inf) = 0)
        z_cl = handed(state_in, "_ef_z_cl")
        if z_cl is None:
            z_cl = pack_cl(state_in[1])
    cur, _ = lif_step_cl(x_cl, v_in, z_cl, w_ff, w_rec, neg_inf, ones, hard_reset=True, w_split=image)
    return cur, x_cl, z_cl


WEIGHT_EPOCH = 0
_TC_IMAGES = {}


def invalidate_weight_images():
    """Call after updating parameters outside torch's version tracking (the fused Adam kernel writes through raw pointers)."""
    global WEIGHT_EPOCH
    WEIGHT_EPOCH += 1


def cell_step(neuron, x, state, w_ff, w_rec, chan, *, hard_reset, surrogate="arctanspike", width=10.0, stride=1, residual=None, x_kind=None,
              detach=True):
    """
    Fused forward of a spiking conv cell.  Mirrors `cell.forward(input_, prev_state, residual)` of
    models/spiking_submodules.py: returns (out, new_state) with new_state = stack([v, z(, trace)]).
    :param chan: dict of per-channel parameters named as in the reference module (leak, thresh, leak_v, ...)
    :param detach: False = the reset term is differentiable (gradient reaches the previous spikes through it; backward only)
    :param x_kind: None (anything: CUDA-core kernel), "spikes" (the caller vouches that x holds spikes / small integer sums, i.e. is exact
                   in bf16) or "split" (<= 10 fractional channels): the convolution of a 32-channel cell then runs on the tensor cores
    """
    if neuron not in _N_STATE:
        raise ValueError(neuron)
    if torch.is_tensor(residual) is False:
        residual = None if (residual is None or residual == 0) else torch.as_tensor(residual)
    meta = (neuron, bool(hard_reset), surrogate, float(width), int(stride), x_kind, bool(detach))
    vals = [chan[n] for n in param_names(neuron)]
    return _CellStep.apply(meta, x, state, w_ff, w_rec, residual, *vals)


class _Pred(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x, w, b = _c(x), _c(w), _c(b)
        _need_cuda(x, w, b)
        B, Cin, H, W = x.shape
        Cout = w.shape[0]
        y = torch.empty((B, Cout, H, W), device=x.device, dtype=torch.float32)
        p = L.PredParams()
        p.B, p.Cin, p.Cout, p.H, p.W = B, Cin, Cout, H, W
        p.x, p.w, p.b, p.y = L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y)
        L.call("ef_pred_fwd", p)
        ctx.save_for_backward(x, w, b, y)
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, w, b, y = ctx.saved_tensors
        g_y = _c(g_y)
        B, Cin, H, W = x.shape
        Cout = w.shape[0]
        p = L.PredParams()
        p.B, p.Cin, p.Cout, p.H, p.W = B, Cin, Cout, H, W
        p.x, p.w, p.b, p.y, p.g_y = L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), L.ptr(g_y)
        g_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        g_w = torch.zeros_like(w)
        g_b = torch.zeros_like(b)
        p.g_x, p.g_w, p.g_b = L.ptr(g_x), L.ptr(g_w), L.ptr(g_b)
        L.call("ef_pred_bwd", p)
        return g_x, g_w, g_b


def pred_head(x, weight, bias):
    """tanh(conv1x1(x) + b): ConvLayer(kernel_size=1, activation="tanh") of models/model.py:197-199."""
    return _Pred.apply(x, weight.reshape(weight.shape[0], -1), bias)


# ---------------------------------------------------------------------------------------------------------------------
# event-warping loss
# ---------------------------------------------------------------------------------------------------------------------
# Workspaces of the loss kernels are recycled: the kernels need the first 16 words zero when a buffer is first used and leave
# them zero (include/eventflow.h), so a buffer is cleared exactly once.  A buffer travels with the autograd node of its
# forward call (the backward reads the accumulator images) and goes back to the pool when the backward has been enqueued --
# reuse is ordered by the stream, hence the stream in the key.
_WS_POOL = {}


def _ws_take(n, dev):
    key = (n, dev, torch.cuda.current_stream(dev).cuda_stream)
    pool = _WS_POOL.setdefault(key, [])
    if pool:
        return pool.pop(), key
    ws = torch.empty(n, device=dev, dtype=torch.float32)
    ws[:16].zero_()
    return ws, key


def _ws_give(ws, key):
    pool = _WS_POOL.setdefault(key, [])
    if len(pool) < 4:
        pool.append(ws)


class _EventWarpingLoss(torch.autograd.Function):
    """Window in map form: everything concatenated (ef_iwe_loss_fwd / ef_iwe_loss_bwd)."""

    @staticmethod
    def forward(ctx, flow_maps, events, pol_mask, event_mask, pass_offsets, meta):
        # flow_maps [S,B,Tm,2,H,W]; events [B,N,4]; pol_mask [B,N,2]; event_mask [B,Tm,H,W]; pass_offsets: list of T+1 ints | None
        T, n_per_pass, flow_scaling, weight, loss_scaling, smoothing_mask, overwrite = meta
        flow_maps, events, pol_mask, event_mask = _c(flow_maps), _c(events), _c(pol_mask), _c(event_mask)
        _need_cuda(flow_maps, events, pol_mask, event_mask)
        _need_f32(flow_maps, events, pol_mask, event_mask)
        S, B, Tm, _, H, W = flow_maps.shape
        p = L.IweLossParams()
        p.S, p.B, p.T, p.T_maps, p.H, p.W = S, B, T, Tm, H, W
        p.n_total, p.n_per_pass = events.shape[1], n_per_pass
        p.flow_scaling, p.weight = float(flow_scaling), float(weight)
        p.loss_scaling, p.smoothing_mask, p.overwrite_intermediate = int(loss_scaling), int(smoothing_mask), int(overwrite)
        ws, key = _ws_take(L.lib().ef_iwe_loss_workspace_elems(S, B, H, W), flow_maps.device)
        loss = torch.empty((), device=flow_maps.device, dtype=torch.float32)
        p.events, p.pol_mask, p.flow_maps, p.event_mask = L.ptr(events), L.ptr(pol_mask), L.ptr(flow_maps), L.ptr(event_mask)
        p.workspace, p.loss = L.ptr(ws), L.ptr(loss)
        ctx.offsets = None
        if pass_offsets is not None:  # host array, read during the call only
            ctx.offsets = (C.c_int32 * (T + 1))(*[int(v) for v in pass_offsets])
            p.pass_offsets = C.cast(ctx.offsets, C.c_void_p)
        L.call("ef_iwe_loss_fwd", p)
        ctx.p, ctx.ws, ctx.key = p, ws, key
        ctx.save_for_backward(flow_maps, events, pol_mask, event_mask)
        if not (torch.is_grad_enabled() and flow_maps.requires_grad):
            _ws_give(ws, key)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        flow_maps = ctx.saved_tensors[0]
        p = ctx.p
        g_loss = g_loss.contiguous().to(torch.float32)
        g_maps = torch.empty_like(flow_maps)
        p.g_loss, p.g_flow_maps = L.ptr(g_loss), L.ptr(g_maps)
        L.call("ef_iwe_loss_bwd", p)
        p.g_loss = p.g_flow_maps = None
        _ws_give(ctx.ws, ctx.key)
        return g_maps, None, None, None, None, None


def event_warping_loss(flow_maps, events, pol_mask, event_mask, *, passes, n_per_pass, flow_scaling, weight, loss_scaling=True,
                       smoothing_mask=True, overwrite_intermediate=False, pass_offsets=None):
    """EventWarping.forward (loss/flow.py:176-301) on a window in map form; differentiable wrt flow_maps."""
    meta = (int(passes), int(n_per_pass), float(flow_scaling), float(weight), bool(loss_scaling), bool(smoothing_mask),
            bool(overwrite_intermediate))
    if pass_offsets is not None and torch.is_tensor(pass_offsets):
        pass_offsets = pass_offsets.tolist()
    return _EventWarpingLoss.apply(flow_maps, events, pol_mask, event_mask, pass_offsets, meta)


class _EventWarpingLossPasses(torch.autograd.Function):
    """
    Window in pass form: the tensors of every pass stay where event_flow_association received them (ef_iwe_loss_fwd_passes /
    ef_iwe_loss_bwd_passes); no torch.cat / torch.stack, no copies.  Differentiable wrt the flow maps.
    """

    @staticmethod
    def forward(ctx, meta, events, pol_masks, masks, *flows):
        S, Tm, T, flow_scaling, weight, loss_scaling, smoothing_mask, overwrite = meta
        B, _, H, W = flows[0].shape
        dev = flows[0].device
        p = L.IweLossPassParams()
        p.S, p.B, p.T, p.T_maps, p.H, p.W = S, B, T, Tm, H, W
        p.flow_scaling, p.weight = float(flow_scaling), float(weight)
        p.loss_scaling, p.smoothing_mask, p.overwrite_intermediate = int(loss_scaling), int(smoothing_mask), int(overwrite)
        keep = []
        for t in range(T):
            e, m = _c(events[t]), _c(pol_masks[t])
            _need_cuda(e, m)
            _need_f32(e, m)
            keep += [e, m]
            p.n_pass[t] = e.shape[1]
            p.events[t], p.pol_mask[t] = L.ptr(e), L.ptr(m)
        for i, f in enumerate(flows):
            f = _c(f)
            _need_cuda(f)
            _need_f32(f)
            keep.append(f)
            p.flow[i] = L.ptr(f)
        if smoothing_mask:
            for m in range(Tm):
                k = _c(masks[m])
                _need_cuda(k)
                _need_f32(k)
                keep.append(k)
                p.event_mask[m] = L.ptr(k)
        ws, key = _ws_take(L.lib().ef_iwe_loss_workspace_elems(S, B, H, W), dev)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        p.workspace, p.loss = L.ptr(ws), L.ptr(loss)
        L.call("ef_iwe_loss_fwd_passes", p)
        ctx.p, ctx.ws, ctx.key, ctx.keep, ctx.shape = p, ws, key, keep, (S * Tm, B, 2, H, W)
        if not (torch.is_grad_enabled() and any(f.requires_grad for f in flows)):
            _ws_give(ws, key)
            ctx.keep = None
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        p = ctx.p
        g_loss = g_loss.contiguous().to(torch.float32)
        g = torch.empty(ctx.shape, device=g_loss.device, dtype=torch.float32)  # one slab, a dense [B,2,H,W] gradient per flow map
        p.g_loss = L.ptr(g_loss)
        for i in range(ctx.shape[0]):
            p.g_flow[i] = L.ptr(g[i])
        L.call("ef_iwe_loss_bwd_passes", p)
        _ws_give(ctx.ws, ctx.key)
        return (None, None, None, None, *g.unbind(0))


def event_warping_loss_passes(flows, events, pol_masks, masks, *, flow_scaling, weight, loss_scaling=True, smoothing_mask=True,
                              overwrite_intermediate=False):
    """
    EventWarping.forward (loss/flow.py:176-301) on a window in pass form.
    :param flows: per scale a list of the window's flow maps [B,2,H,W] (one per pass, or the single final map with overwrite)
    :param events / pol_masks: per pass [B,N_t,4] (ts offset by the pass index) / [B,N_t,2]
    :param masks: per flow map [B,1,H,W]
    """
    S, Tm, T = len(flows), len(flows[0]), len(events)
    if T > L.EF_IWE_MAX_PASSES or S > L.EF_IWE_MAX_SCALES:
        raise L.EventFlowError(f"event-warping loss: at most {L.EF_IWE_MAX_PASSES} passes and {L.EF_IWE_MAX_SCALES} flow scales per window")
    meta = (S, Tm, T, float(flow_scaling), float(weight), bool(loss_scaling), bool(smoothing_mask), bool(overwrite_intermediate))
    flat = [f for per_scale in flows for f in per_scale]
    return _EventWarpingLossPasses.apply(meta, list(events), list(pol_masks), list(masks), *flat)


def iwe_image(events, pol_mask, res, *, flow=None, event_flow=None, tref=1.0, flow_scaling=128.0, round_idx=True):
    """Per-polarity image of warped events [B,2,H,W] (utils/iwe.py:95-153)."""
    f32 = lambda t: None if t is None else _c(t.float())  # noqa: E731  (callers pass float64 event lists straight from the loaders)
    events, pol_mask, flow, event_flow = f32(events), f32(pol_mask), f32(flow), f32(event_flow)
    _need_cuda(events, pol_mask, flow, event_flow)
    B, N = events.shape[:2]
    H, W = res
    out = torch.empty((B, 2, H, W), device=events.device, dtype=torch.float32)
    p = L.IweImageParams()
    p.B, p.N, p.H, p.W, p.round_idx = B, N, H, W, int(round_idx)
    p.tref, p.flow_scaling = float(tref), float(flow_scaling)
    p.events, p.pol_mask, p.flow, p.event_flow, p.iwe = L.ptr(events), L.ptr(pol_mask), L.ptr(flow), L.ptr(event_flow), L.ptr(out)
    L.call("ef_iwe_image", p)
    return out


def encode_events(events, res, num_bins, *, round_ts=False, want=("cnt", "voxel", "mask", "pol_mask")):
    """Event encodings of one batch (dataloader/encodings.py:30-85, base.py:148-222).  events [B,N,4] (ts,y,x,p)."""
    events = _c(events)
    _need_cuda(events)
    B, N = events.shape[:2]
    H, W = res
    dev = events.device
    out = {}
    p = L.EncodeParams()
    p.B, p.N, p.H, p.W, p.num_bins, p.round_ts = B, N, H, W, int(num_bins), int(round_ts)
    p.events = L.ptr(events)
    if "cnt" in want and "voxel" in want and "mask" in want:
        # one allocation for the three images (cnt | voxel | mask): the library zero-fills them with a single memset
        flat = torch.empty(B * (2 + num_bins + 1) * H * W, device=dev, dtype=torch.float32)
        n1, n2 = B * 2 * H * W, B * num_bins * H * W
        out["event_cnt"] = flat[:n1].view(B, 2, H, W)
        out["event_voxel"] = flat[n1:n1 + n2].view(B, num_bins, H, W)
        out["event_mask"] = flat[n1 + n2:].view(B, 1, H, W)
        p.cnt, p.voxel, p.mask = L.ptr(out["event_cnt"]), L.ptr(out["event_voxel"]), L.ptr(out["event_mask"])
    else:
        if "cnt" in want:
            out["event_cnt"] = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
            p.cnt = L.ptr(out["event_cnt"])
        if "voxel" in want:
            out["event_voxel"] = torch.empty((B, num_bins, H, W), device=dev, dtype=torch.float32)
            p.voxel = L.ptr(out["event_voxel"])
        if "mask" in want:
            out["event_mask"] = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
            p.mask = L.ptr(out["event_mask"])
    if "pol_mask" in want:
        out["event_list_pol_mask"] = torch.empty((B, N, 2), device=dev, dtype=torch.float32)
        p.pol_mask = L.ptr(out["event_list_pol_mask"])
    L.call("ef_encode_events", p)
    return out


class _UpsampleBilinear2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        _need_cuda(x)
        B, Cc, H, W = x.shape
        out = torch.empty((B, Cc, 2 * H, 2 * W), device=x.device, dtype=torch.float32)
        L.LAUNCHES += 1
        L.check(L.lib().ef_upsample_bilinear2x(L.ptr(x), L.ptr(out), B * Cc, H, W, L.stream()), "ef_upsample_bilinear2x")
        ctx.shape = (B, Cc, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        B, Cc, H, W = ctx.shape
        g = _c(g)
        g_x = torch.empty((B, Cc, H, W), device=g.device, dtype=torch.float32)
        L.LAUNCHES += 1
        L.check(L.lib().ef_upsample_bilinear2x_bwd(L.ptr(g), L.ptr(g_x), B * Cc, H, W, L.stream()), "ef_upsample_bilinear2x_bwd")
        return g_x


def upsample_bilinear2x(x):
    """F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False) on fp32 NCHW, differentiable."""
    return _UpsampleBilinear2x.apply(x)


class _UpsampleNearest(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fy, fx):
        x = _c(x)
        _need_cuda(x)
        B, Cc, H, W = x.shape
        out = torch.empty((B, Cc, H * fy, W * fx), device=x.device, dtype=torch.float32)
        L.LAUNCHES += 1
        L.check(L.lib().ef_upsample_nearest(L.ptr(x), L.ptr(out), B * Cc, H, W, fy, fx, L.stream()), "ef_upsample_nearest")
        ctx.meta = (B, Cc, H, W, fy, fx)
        return out

    @staticmethod
    def backward(ctx, g):
        B, Cc, H, W, fy, fx = ctx.meta
        g = _c(g)
        g_x = torch.empty((B, Cc, H, W), device=g.device, dtype=torch.float32)
        L.LAUNCHES += 1
        L.check(L.lib().ef_upsample_nearest_bwd(L.ptr(g), L.ptr(g_x), B * Cc, H, W, fy, fx, L.stream()), "ef_upsample_nearest_bwd")
        return g_x, None, None


def upsample_nearest(x, fy, fx):
    """F.interpolate(x, scale_factor=(fy, fx)) (nearest) for integer factors (models/model.py:528-539), differentiable."""
    if fy == 1 and fx == 1:
        return x
    return _UpsampleNearest.apply(x, int(fy), int(fx))


def pack_cl(x):
    """fp32 NCHW -> bf16 channels-last [B,H,W,C] (internal spike format)."""
    x = _c(x)
    _need_cuda(x)
    B, Cc, H, W = x.shape
    out = torch.empty((B, H, W, Cc), device=x.device, dtype=torch.bfloat16)
    L.LAUNCHES += 1
    L.check(L.lib().ef_pack_cl(L.ptr(x), L.ptr(out), B, Cc, H, W, L.stream()), "ef_pack_cl")
    return out


def upsample_bilinear2x_cl(x_cl):
    """Bilinear x2 upsampling of a cl bf16 tensor [B,H,W,C] (exact for spikes / residual sums).  No autograd."""
    x_cl = _c(x_cl)
    _need_cuda(x_cl)
    B, H, W, Cc = x_cl.shape
    out = torch.empty((B, 2 * H, 2 * W, Cc), device=x_cl.device, dtype=torch.bfloat16)
    L.LAUNCHES += 1
    L.check(L.lib().ef_upsample_bilinear2x_cl(L.ptr(x_cl), L.ptr(out), B, H, W, Cc, L.stream()), "ef_upsample_bilinear2x_cl")
    return out


def space_to_depth_cl(x_cl):
    """cl [B,H,W,C] -> [B,H/2,W/2,4C]: the four pixel parities of every channel side by side (input form of the stride-2 tensor-core cell)."""
    x_cl = _c(x_cl)
    _need_cuda(x_cl)
    B, H, W, Cc = x_cl.shape
    out = torch.empty((B, H // 2, W // 2, 4 * Cc), device=x_cl.device, dtype=torch.bfloat16)
    L.LAUNCHES += 1
    L.check(L.lib().ef_space_to_depth_cl(L.ptr(x_cl), L.ptr(out), B, H, W, Cc, L.stream()), "ef_space_to_depth_cl")
    return out


def pack_split_s2d_cl(x):
    """fp32 NCHW [B,Cin,H,W], 4*Cin <= 10 -> cl [B,H/2,W/2,32]: exact hi/mid/lo split of the space-to-depth form (first stride-2 encoder)."""
    x = _c(x)
    _need_cuda(x)
    _need_f32(x)
    B, Cin, H, W = x.shape
    out = torch.empty((B, H // 2, W // 2, 32), device=x.device, dtype=torch.bfloat16)
    L.LAUNCHES += 1
    L.check(L.lib().ef_pack_split_s2d_cl(L.ptr(x), L.ptr(out), B, Cin, H, W, L.stream()), "ef_pack_split_s2d_cl")
    return out


def pack_split_cl(x, out=None):
    """fp32 NCHW [B,Cin<=10,H,W] -> bf16 cl [B,H,W,32] holding the exact hi/mid/lo split of every value (ef_pack_split_cl)."""
    x = _c(x)
    _need_cuda(x)
    _need_f32(x)
    B, Cin, H, W = x.shape
    if out is None:
        out = torch.empty((B, H, W, 32), device=x.device, dtype=torch.bfloat16)
    L.LAUNCHES += 1
    L.check(L.lib().ef_pack_split_cl(L.ptr(x), L.ptr(out), B, Cin, H, W, L.stream()), "ef_pack_split_cl")
    return out


def split_weights_head(w_ff, out=None):
    """Weight image of the head layer for inputs packed by pack_split_cl (ef_split_weights_head)."""
    w_ff = _c(w_ff.detach())
    _need_cuda(w_ff)
    n = L.lib().ef_split_weights_elems(32, 32, 0)
    if out is None or out.numel() != n:
        out = torch.empty(n, device=w_ff.device, dtype=torch.int16)
    L.LAUNCHES += 1
    L.check(L.lib().ef_split_weights_head(L.ptr(w_ff), w_ff.shape[1], L.ptr(out), L.stream()), "ef_split_weights_head")
    return out


def unpack_cl(x):
    """bf16 channels-last [B,H,W,C] -> fp32 NCHW."""
    x = _c(x)
    _need_cuda(x)
    B, H, W, Cc = x.shape
    out = torch.empty((B, Cc, H, W), device=x.device, dtype=torch.float32)
    L.LAUNCHES += 1
    L.check(L.lib().ef_unpack_cl(L.ptr(x), L.ptr(out), B, Cc, H, W, L.stream()), "ef_unpack_cl")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# internal-format (cl spikes) LIF step: the building block of the fast model path
# ---------------------------------------------------------------------------------------------------------------------
def split_weights(w_ff, w_rec=None, out=None):
    """fp32 conv weights -> three exact bf16 terms in the tcgen05 B-operand layout (uint16 tensor); `out` is refilled in place."""
    w_ff, w_rec = _c(w_ff.detach()), (None if w_rec is None else _c(w_rec.detach()))
    _need_cuda(w_ff, w_rec)
    C, Cin = w_ff.shape[:2]
    n = L.lib().ef_split_weights_elems(Cin, C, int(w_rec is not None))
    if n == 0:
        return None
    if out is None or out.numel() != n:
        out = torch.empty(n, device=w_ff.device, dtype=torch.int16)
    L.LAUNCHES += 1
    L.check(L.lib().ef_split_weights(L.ptr(w_ff), L.ptr(w_rec), Cin, C, L.ptr(out), L.stream()), "ef_split_weights")
    return out


def split_weights_bwd(w_ff, w_rec=None, out=None):
    """Flipped / transposed bf16 hi+mid weight image for the tensor-core data gradient (ef_split_weights_bwd)."""
    w_ff, w_rec = _c(w_ff.detach()), (None if w_rec is None else _c(w_rec.detach()))
    _need_cuda(w_ff, w_rec)
    n = L.lib().ef_split_weights_bwd_elems(int(w_rec is not None))
    if out is None or out.numel() != n:
        out = torch.empty(n, device=w_ff.device, dtype=torch.int16)
    L.LAUNCHES += 1
    L.check(L.lib().ef_split_weights_bwd(L.ptr(w_ff), L.ptr(w_rec), L.ptr(out), L.stream()), "ef_split_weights_bwd")
    return out


def split_weights_g(sources, C):
    """
    Weight image of a general tensor-core cell (ef_split_weights_g).  sources: list of (weight [C,c_total,3,3], ch0, n, split) -- the
    channel slices of the conv weights that multiply each input source, in input order.  Returns (image, list of source channel counts).
    """
    n = len(sources)
    arr = (L.WSrc * n)()
    keep = []
    for i, src in enumerate(sources):
        w, ch0, cnt, split = src[:4]
        s2d = src[4] if len(src) > 4 else False
        w = _c(w.detach())
        _need_cuda(w)
        _need_f32(w)
        keep.append(w)
        arr[i].w, arr[i].c_total, arr[i].ch0, arr[i].n, arr[i].split = L.ptr(w), w.shape[1], int(ch0), int(cnt), int(bool(split))
        arr[i].s2d = int(bool(s2d))
    elems = L.lib().ef_split_weights_g_elems(C, n, arr)
    if elems <= 0:
        raise L.EventFlowError("split_weights_g: unsupported shape (C must be a multiple of 32)")
    out = torch.empty(elems, device=keep[0].device, dtype=torch.int16)
    L.LAUNCHES += 1
    L.check(L.lib().ef_split_weights_g(arr, n, C, L.ptr(out), L.stream()), "ef_split_weights_g")
    return out


def lif_step_g(srcs, v_in, z_in_cl, w_image, leak, thresh, C, *, hard_reset=True, residual_cl=None, s2d=False):
    """
    One fused conv + LIF step of a general-channel cell on the tensor cores (ef_lif_conv_fwd_g).  srcs: cl bf16 tensors [B,H,W,c_s]
    (c_s multiples of 32) in the order of the weight image's sources.  Returns (v_out fp32 NCHW, z_out cl, out cl | None).  No autograd.
    """
    B, H, W, _ = srcs[0].shape
    dev = srcs[0].device
    v_out = torch.empty((B, C, H, W), device=dev, dtype=torch.float32)
    z_out = torch.empty((B, H, W, C), device=dev, dtype=torch.bfloat16)
    out = torch.empty_like(z_out) if residual_cl is not None else None
    p = L.LifConvGParams()
    p.B, p.H, p.W, p.C, p.n_src, p.hard_reset, p.s2d = B, H, W, C, len(srcs), int(hard_reset), int(bool(s2d))
    for i, s in enumerate(srcs):
        p.src[i], p.src_c[i] = L.ptr(s), s.shape[3]
    p.v_in, p.z_in_cl, p.residual_cl = L.ptr(v_in), L.ptr(z_in_cl), L.ptr(residual_cl)
    p.leak, p.thresh, p.w_image = L.ptr(leak), L.ptr(thresh), L.ptr(w_image)
    p.v_out, p.z_out_cl, p.out_cl = L.ptr(v_out), L.ptr(z_out), L.ptr(out)
    L.call("ef_lif_conv_fwd_g", p)
    return v_out, z_out, out


def lif_step_cl(x_cl, v_in, z_in_cl, w_ff, w_rec, leak, thresh, *, hard_reset=True, w_split=None, x_f32=None, stride=1):
    """
    One fused conv + LIF step on the internal formats: spikes bf16 channels-last [B,H,W,C], membrane fp32 NCHW.
    With `w_split` (ops.split_weights) and 32->32 channels the tcgen05 kernel runs, otherwise the CUDA-core kernel.
    `x_f32` (fp32 NCHW) may replace x_cl for the first layer.  Returns (v_out, z_out_cl).  No autograd.
    """
    if x_cl is not None:
        B, H, W, Cin = x_cl.shape
    else:
        B, Cin, H, W = x_f32.shape
    C = w_ff.shape[0]
    dev = w_ff.device
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    v_out = torch.empty((B, C, Ho, Wo), device=dev, dtype=torch.float32)
    z_out = torch.empty((B, Ho, Wo, C), device=dev, dtype=torch.bfloat16)
    p = L.LifConvParams()
    p.B, p.Cin, p.C, p.H, p.W = B, Cin, C, H, W
    p.ksize, p.stride, p.neuron, p.hard_reset = 3, int(stride), L.EF_LIF, int(hard_reset)
    p.surrogate, p.act_width = 0, 10.0
    p.x, p.x_cl = L.ptr(x_f32), L.ptr(x_cl)
    p.v_in, p.z_in_cl = L.ptr(v_in), L.ptr(z_in_cl)
    p.w_ff, p.w_rec, p.w_split = L.ptr(w_ff), L.ptr(w_rec), L.ptr(w_split)
    p.leak, p.thresh = L.ptr(leak), L.ptr(thresh)
    p.v_out, p.z_out_cl = L.ptr(v_out), L.ptr(z_out)
    L.call("ef_lif_conv_fwd", p, tag=(Cin, C, w_rec is not None))
    return v_out, z_out


def lif_bwd_cl(x_cl, v_in, z_in_cl, v_out, g_out, g_v_out, g_z_out, w_ff, w_rec, leak, thresh, *, hard_reset=True,
               surrogate="arctanspike", act_width=10.0, tc_wgrad=True, steps=1):
    """
    Backward of one 32->32 LIF cell-step on the internal formats (ef_lif_bwd_tc): tensor-core data gradient and, with
    `tc_wgrad`, the tensor-core weight gradient (`steps` > 1 repeats the call to exercise the accumulate / finalize
    protocol: the weight gradients are then `steps` times those of one call).  Returns a dict of gradients.  No autograd.
    """
    B, H, W, _ = x_cl.shape
    dev = x_cl.device
    rec = w_rec is not None
    f32 = lambda *shape: torch.zeros(shape, device=dev, dtype=torch.float32)  # noqa: E731
    out = {"g_x": f32(B, 32, H, W), "g_v_in": f32(B, 32, H, W), "g_w_ff": f32(32, 32, 3, 3), "g_leak": f32(32), "g_thresh": f32(32)}
    if rec:
        out["g_z_in"], out["g_w_rec"] = f32(B, 32, H, W), f32(32, 32, 3, 3)
    gI_hi = torch.empty((B, H, W, 32), device=dev, dtype=torch.bfloat16)
    gI_mid = torch.empty_like(gI_hi)
    w_bwd = split_weights_bwd(w_ff, w_rec)
    partial = torch.empty(L.lib().ef_lif_wgrad_partial_elems(B, H, W, int(rec)), device=dev, dtype=torch.float32) if tc_wgrad else None
    for k in range(steps):
        t = L.LifBwdTcParams()
        t.B, t.H, t.W, t.has_rec, t.hard_reset = B, H, W, int(rec), int(hard_reset)
        t.surrogate, t.act_width = L.SURROGATE_CODES[surrogate], float(act_width)
        t.x_cl, t.z_in_cl, t.v_in, t.v_out = L.ptr(x_cl), L.ptr(z_in_cl), L.ptr(v_in), L.ptr(v_out)
        t.g_out, t.g_v_out, t.g_z_out = L.ptr(g_out), L.ptr(g_v_out), L.ptr(g_z_out)
        t.leak, t.thresh, t.w_bwd = L.ptr(leak), L.ptr(thresh), L.ptr(w_bwd)
        t.gI_hi, t.gI_mid = L.ptr(gI_hi), L.ptr(gI_mid)
        t.g_x, t.g_v_in, t.g_z_in = L.ptr(out["g_x"]), L.ptr(out["g_v_in"]), L.ptr(out.get("g_z_in"))
        t.g_w_ff, t.g_w_rec = L.ptr(out["g_w_ff"]), L.ptr(out.get("g_w_rec"))
        t.g_leak, t.g_thresh = L.ptr(out["g_leak"]), L.ptr(out["g_thresh"])
        if tc_wgrad:
            t.wg_partial = L.ptr(partial)
            t.wg_flags = (L.EF_WG_ACCUMULATE if k > 0 else 0) | (L.EF_WG_FINALIZE if k == steps - 1 else 0)
        L.call("ef_lif_bwd_tc", t)
    out["gI"] = gI_hi.float() + gI_mid.float()
    return out


# ---------------------------------------------------------------------------------------------------------------------
# validation metrics
# ---------------------------------------------------------------------------------------------------------------------
def iwe_metrics(flow_maps, events, pol_mask, *, passes, n_per_pass, flow_scaling, pass_offsets=None):
    """FWL and RSAT of a validation window (loss/flow.py:468-579).  flow_maps [B,Tm,2,H,W].  Returns (FWL [B], RSAT [B])."""
    flow_maps, events, pol_mask = _c(flow_maps.detach().float()), _c(events.float()), _c(pol_mask.float())
    _need_cuda(flow_maps, events, pol_mask)
    B, Tm, _, H, W = flow_maps.shape
    p = L.IweMetricsParams()
    p.B, p.T, p.T_maps, p.H, p.W = B, int(passes), Tm, H, W
    p.n_total, p.n_per_pass, p.flow_scaling = events.shape[1], int(n_per_pass), float(flow_scaling)
    ws = torch.empty(L.lib().ef_iwe_metrics_workspace_elems(B, H, W), device=events.device, dtype=torch.float32)
    out = torch.empty((B, 4), device=events.device, dtype=torch.float32)
    p.events, p.pol_mask, p.flow_maps, p.pass_offsets = L.ptr(events), L.ptr(pol_mask), L.ptr(flow_maps), L.ptr(pass_offsets)
    p.workspace, p.out = L.ptr(ws), L.ptr(out)
    L.call("ef_iwe_metrics", p)
    return out[:, 0], out[:, 1]


def aee(flow, gtflow, event_mask, dt_ratio, flow_scaling):
    """AEE and outlier percentage (loss/flow.py:597-628).  flow/gtflow [B,2,H,W], event_mask [B,H,W], dt_ratio [B]."""
    flow, gtflow, event_mask = _c(flow.detach().float()), _c(gtflow.float()), _c(event_mask.float())
    dt_ratio = _c(dt_ratio.to(flow.device).float().reshape(-1))
    _need_cuda(flow, gtflow, event_mask, dt_ratio)
    B, _, H, W = flow.shape
    if dt_ratio.numel() == 1 and B > 1:
        dt_ratio = dt_ratio.expand(B).contiguous()
    p = L.AeeParams()
    p.B, p.H, p.W, p.flow_scaling = B, H, W, float(flow_scaling)
    ws = torch.empty(2 * B + 1, device=flow.device, dtype=torch.float32)
    out = torch.empty((2, B), device=flow.device, dtype=torch.float32)
    p.flow, p.gtflow, p.event_mask, p.dt_ratio, p.workspace, p.out = (L.ptr(flow), L.ptr(gtflow), L.ptr(event_mask), L.ptr(dt_ratio),
                                                                     L.ptr(ws), L.ptr(out))
    L.call("ef_aee", p)
    return out[0], out[1]


# ---------------------------------------------------------------------------------------------------------------------
# ANN cells (forward only in this version)
# ---------------------------------------------------------------------------------------------------------------------
_ACT_CODES = {None: 0, "relu": 1, "sigmoid": 2, "tanh": 3}


def _plane_ok(t):
    return t is None or (t.stride(-1) == 1 and t.stride(-2) == t.shape[-1] and t.stride(1) == t.shape[-1] * t.shape[-2])


class _ConvAnn(torch.autograd.Function):
    """
    blend(act(conv3x3(cat([x1, x2 * x2_scale]), w) + b + residual)) with gradients for every tensor input, kernels only:
    forward = ONE launch (ef_conv_ann_fwd, stride 1 or 2; with a blend it also stores the activation before the blend),
    backward = ef_ann_gate_bwd (blend + activation derivative + bias gradient), ef_ann_cat_scale (the convolution's input as one
    tensor, only when there is a second input), ef_conv3x3_bwd_s (data and weight gradients), ef_ann_scale_bwd (gate product).
    """

    @staticmethod
    def forward(ctx, x1, x2, x2_scale, weight, bias, residual, blend_h, blend_u, act, stride):
        need_o = blend_h is not None
        out, act_out = _conv_ann_launch(x1, weight, bias, act, x2=x2, x2_scale=x2_scale, residual=residual, blend_h=blend_h, blend_u=blend_u,
                                        stride=stride, want_act_out=need_o)
        ctx.act, ctx.stride = act, stride
        ctx.has = (x2 is not None, x2_scale is not None, bias is not None, residual is not None, blend_h is not None)
        ctx.save_for_backward(x1, x2, x2_scale, weight, blend_h, blend_u, act_out if need_o else out)
        return out

    @staticmethod
    def backward(ctx, g):
        x1, x2, x2_scale, weight, blend_h, blend_u, o = ctx.saved_tensors
        has_x2, has_scale, has_bias, has_res, has_blend = ctx.has
        need = ctx.needs_input_grad  # (x1, x2, x2_scale, weight, bias, residual, blend_h, blend_u, act, stride)
        g = g.contiguous()
        B, Cout, Ho, Wo = g.shape
        dev = g.device
        # (1) blend + activation + bias
        q = L.AnnGateBwdParams()
        q.B, q.C, q.H, q.W, q.act = B, Cout, Ho, Wo, _ACT_CODES[ctx.act]
        g_pre = torch.empty_like(g)
        g_h = torch.empty_like(g) if (has_blend and need[6]) else None
        g_u = torch.empty_like(g) if (has_blend and need[7]) else None
        g_b = torch.zeros(Cout, device=dev, dtype=torch.float32) if (has_bias and need[4]) else None
        q.g_y, q.act_out, q.g_pre, q.g_h, q.g_u, q.g_bias = L.ptr(g), L.ptr(o), L.ptr(g_pre), L.ptr(g_h), L.ptr(g_u), L.ptr(g_b)
        if has_blend:
            bh = blend_h if _plane_ok(blend_h) else blend_h.contiguous()
            bu = blend_u if _plane_ok(blend_u) else blend_u.contiguous()
            q.blend_h, q.blend_u, q.blend_h_bstride, q.blend_u_bstride = bh.data_ptr(), bu.data_ptr(), bh.stride(0), bu.stride(0)
        L.call("ef_ann_gate_bwd", q)
        # (2) the convolution's input as one tensor
        _, C1, H, W = x1.shape
        C2 = x2.shape[1] if has_x2 else 0
        if has_x2:
            xa = x1 if _plane_ok(x1) else x1.contiguous()
            xb = x2 if _plane_ok(x2) else x2.contiguous()
            sc = None if not has_scale else (x2_scale if _plane_ok(x2_scale) else x2_scale.contiguous())
            x = torch.empty((B, C1 + C2, H, W), device=dev, dtype=torch.float32)
            L.LAUNCHES += 1
            L.check(L.lib().ef_ann_cat_scale(xa.data_ptr(), xb.data_ptr(), None if sc is None else sc.data_ptr(), L.ptr(x), B, C1, C2, H, W,
                                             xa.stride(0), xb.stride(0), 0 if sc is None else sc.stride(0), L.stream()), "ef_ann_cat_scale")
        else:
            x = x1.contiguous()
        # (3) convolution gradients
        need_gx = need[0] or (has_x2 and (need[1] or (has_scale and need[2])))
        g_x = torch.empty_like(x) if need_gx else None
        g_w = torch.zeros_like(weight) if need[3] else None
        up = torch.empty((B, Cout, H, W), device=dev, dtype=torch.float32) if ctx.stride == 2 else None
        L.LAUNCHES += 1
        L.check(L.lib().ef_conv3x3_bwd_s(L.ptr(g_pre), L.ptr(x), L.ptr(_c(weight)), L.ptr(g_x), L.ptr(g_w), L.ptr(up), B, C1 + C2, Cout, H, W,
                                         int(ctx.stride), L.stream()), "ef_conv3x3_bwd_s")
        g_x1 = g_x[:, :C1] if (g_x is not None and need[0]) else None
        g_x2 = g_sc = None
        if has_x2 and g_x is not None:
            if has_scale:
                g_x2 = torch.empty((B, C2, H, W), device=dev, dtype=torch.float32) if need[1] else None
                g_sc = torch.empty((B, C2, H, W), device=dev, dtype=torch.float32) if need[2] else None
                if g_x2 is not None or g_sc is not None:
                    L.LAUNCHES += 1
                    L.check(L.lib().ef_ann_scale_bwd(L.ptr(g_x), xb.data_ptr(), sc.data_ptr(), L.ptr(g_x2), L.ptr(g_sc), B, C1, C2, H, W,
                                                     xb.stride(0), sc.stride(0), L.stream()), "ef_ann_scale_bwd")
            elif need[1]:
                g_x2 = g_x[:, C1:]
        g_r = g_pre if (has_res and need[5]) else None
        return g_x1, g_x2, g_sc, g_w, g_b, g_r, g_h, g_u, None, None


def conv_ann(x1, weight, bias, act, *, x2=None, x2_scale=None, residual=None, blend_h=None, blend_u=None, stride=1):
    """
    act(conv3x3(cat([x1, x2 * x2_scale]), weight) + bias + residual), optionally blended h*(1-u) + (.)*u; stride 1 or 2.
    x2 / x2_scale / blend_* may be channel slices of larger NCHW tensors (only the batch stride may be non-dense).
    ONE fused launch; differentiable wrt every tensor argument (_ConvAnn: the backward is kernels only).
    """
    tracked = torch.is_grad_enabled() and any(t is not None and t.requires_grad
                                              for t in (x1, x2, x2_scale, weight, bias, residual, blend_h, blend_u))
    if tracked:
        return _ConvAnn.apply(x1, x2, x2_scale, weight, bias, residual, blend_h, blend_u, act, int(stride))
    return _conv_ann_launch(x1, weight, bias, act, x2=x2, x2_scale=x2_scale, residual=residual, blend_h=blend_h, blend_u=blend_u, stride=stride,
                            inference=True)[0]


def _conv_ann_launch(x1, weight, bias, act, *, x2=None, x2_scale=None, residual=None, blend_h=None, blend_u=None, stride=1, want_act_out=False,
                     inference=False):
    x1 = x1 if _plane_ok(x1) else x1.contiguous()
    tensors = [x2, x2_scale, residual, blend_h, blend_u]
    tensors = [t if _plane_ok(t) else t.contiguous() for t in tensors]
    x2, x2_scale, residual, blend_h, blend_u = tensors
    weight, bias = _c(weight.detach()), (None if bias is None else _c(bias.detach()))
    residual = None if residual is None else residual.contiguous()
    _need_cuda(x1, x2, x2_scale, residual, blend_h, blend_u, weight, bias)
    B, C1, H, W = x1.shape
    C2 = 0 if x2 is None else x2.shape[1]
    Cout = weight.shape[0]
    assert weight.shape[1] == C1 + C2 and weight.shape[2:] == (3, 3), "conv_ann: weight shape does not match the inputs"
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = torch.empty((B, Cout, Ho, Wo), device=x1.device, dtype=torch.float32)
    act_out = torch.empty_like(out) if want_act_out else None
    p = L.ConvAnnParams()
    p.B, p.C1, p.C2, p.Cout, p.H, p.W, p.act = B, C1, C2, Cout, H, W, _ACT_CODES[act]
    p.stride, p.inference = int(stride), int(bool(inference))
    raw = lambda t: None if t is None else t.data_ptr()  # noqa: E731  (slices are not "contiguous"; strides are passed explicitly)
    p.x1, p.x2, p.x2_scale = raw(x1), raw(x2), raw(x2_scale)
    p.x1_bstride = x1.stride(0)
    p.x2_bstride = 0 if x2 is None else x2.stride(0)
    p.x2_scale_bstride = 0 if x2_scale is None else x2_scale.stride(0)
    p.w, p.bias, p.residual = L.ptr(weight), L.ptr(bias), L.ptr(residual)
    p.blend_h, p.blend_u = raw(blend_h), raw(blend_u)
    p.blend_h_bstride = 0 if blend_h is None else blend_h.stride(0)
    p.blend_u_bstride = 0 if blend_u is None else blend_u.stride(0)
    p.out, p.act_out = L.ptr(out), L.ptr(act_out)
    L.call("ef_conv_ann_fwd", p)
    return out, act_out
