"""
Fast path of the spiking FireNet chain (models/model.py:254-265) on the internal formats.

Between the cells, spikes never exist as fp32 NCHW tensors: every cell writes bf16 channel-blocked spikes ("c8") that the
next cell's tcgen05 kernel consumes through TMA; membrane potentials stay fp32 NCHW (the reference's state format).  One
torch.autograd node per MODEL step (instead of ~100 per step in the reference) carries the BPTT: the per-layer state
gradients travel from step t+1 to step t in a side structure (`_Carry`), the autograd graph only orders the steps through
a scalar token and routes the flow / parameter gradients.
"""
import torch

from . import _lib as L
from . import ops

LAYERS = ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b")


class _Carry:
    """dL/d(v, z) of every layer's state, handed from the backward of step t+1 to the backward of step t."""

    def __init__(self):
        self.g_v = [None] * len(LAYERS)
        self.g_z = [None] * len(LAYERS)


class FastState:
    """Per-window internal state of a model on the fast path."""

    def __init__(self, n):
        self.v = [None] * n        # fp32 [B,C,H,W]
        self.z = [None] * n        # bf16 [B,C/8,H,W,8]
        self.token = None          # scalar autograd token ordering the steps of one BPTT window
        self.carry = _Carry()

    def detach(self):
        self.token = None
        self.carry = _Carry()


def eligible(model, x):
    """LIF cells, 32 channels, 3x3, stride 1, CUDA, W % 4 == 0 (TMA stride rule for the fp32 membrane tensor)."""
    if not x.is_cuda or model.residual:
        return False
    for name in LAYERS:
        cell = getattr(model, name)
        if getattr(cell, "neuron", None) != "lif" or cell.hidden_size != 32 or cell.ff.kernel_size != (3, 3) or cell.stride != 1:
            return False
        if name != "head" and cell.input_size != 32:
            return False
    return x.shape[-1] % 4 == 0


def _split_cache(model):
    """bf16 hi/mid/lo weight images of the hidden layers, rebuilt when a weight tensor changed."""
    cache = model.__dict__.setdefault("_w_split_cache", {})
    out = {}
    for name in LAYERS[1:]:
        cell = getattr(model, name)
        rec = cell.rec.weight if cell.recurrent else None
        key = (cell.ff.weight._version, cell.ff.weight.data_ptr(), None if rec is None else rec._version, model.__dict__.get("_w_epoch", 0))
        hit = cache.get(name)
        if hit is None or hit[0] != key:
            hit = (key, ops.split_weights(cell.ff.weight, rec))
            cache[name] = hit
        out[name] = hit[1]
    return out


def invalidate_weights(model):
    """Call after updating parameters outside torch's version tracking (e.g. the fused Adam kernel)."""
    model.__dict__["_w_epoch"] = model.__dict__.get("_w_epoch", 0) + 1


def _params_of(model):
    ps = []
    for name in LAYERS:
        cell = getattr(model, name)
        ps.append(cell.ff.weight)
        if cell.recurrent:
            ps.append(cell.rec.weight)
        ps.append(cell.leak)
        ps.append(cell.thresh)
    ps.append(model.pred.conv2d.weight)
    ps.append(model.pred.conv2d.bias)
    return ps


def _fill_fwd(p, B, Cin, H, W, cell, x_f32, x_c8, v_in, z_in, v_out, leak, thresh):
    p.B, p.Cin, p.C, p.H, p.W = B, Cin, 32, H, W
    p.ksize, p.stride, p.neuron, p.hard_reset = 3, 1, L.EF_LIF, int(cell.hard_reset)
    p.surrogate, p.act_width = L.SURROGATE_CODES[cell.activation], float(cell._act_width_f)
    p.x, p.x_c8 = L.ptr(x_f32), L.ptr(x_c8)
    p.v_in, p.z_in_c8 = L.ptr(v_in), L.ptr(z_in)
    p.w_ff = L.ptr(cell.ff.weight)
    p.w_rec = L.ptr(cell.rec.weight) if cell.recurrent else None
    p.leak, p.thresh = L.ptr(leak), L.ptr(thresh)
    p.v_out = L.ptr(v_out)


class _FireNetStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, token, *params):
        fs = model._fast
        x = x.contiguous()
        B, Cin0, H, W = x.shape
        dev = x.device
        splits = _split_cache(model)
        saved = []
        zs = []
        h = None
        for i, name in enumerate(LAYERS):
            cell = getattr(model, name)
            if not hasattr(cell, "_act_width_f"):
                cell._act_width_f = float(cell.act_width)
            leak, thresh = cell.leak.detach().reshape(-1), cell.thresh.detach().reshape(-1)
            v_in, z_in = fs.v[i], fs.z[i]
            v_out = torch.empty((B, 32, H, W), device=dev, dtype=torch.float32)
            z_out = torch.empty((B, 4, H, W, 8), device=dev, dtype=torch.bfloat16)
            p = L.LifConvParams()
            _fill_fwd(p, B, Cin0 if i == 0 else 32, H, W, cell, x if i == 0 else None, h, v_in, z_in, v_out, leak, thresh)
            p.z_out_c8 = L.ptr(z_out)
            if i > 0:
                p.w_split = L.ptr(splits[name])
            L.call("ef_lif_conv_fwd", p, tag=(p.Cin, 32, cell.recurrent))
            saved.append((x if i == 0 else None, h, v_in, z_in, v_out))
            cap = model.__dict__.get("_capture")
            if cap is not None:  # test hook: what this layer consumed and produced, in the reference's tensor format
                xin = x if i == 0 else ops.unpack_c8(h)
                sin = None if v_in is None else torch.stack([v_in, ops.unpack_c8(z_in)]).cpu()
                zo = ops.unpack_c8(z_out)
                cap[name] = (xin.detach().cpu(), sin, zo.cpu(), torch.stack([v_out, zo]).cpu())
            fs.v[i], fs.z[i] = v_out, z_out
            zs.append(z_out)
            h = z_out
        w, b = model.pred.conv2d.weight.detach(), model.pred.conv2d.bias.detach()
        flow = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
        pp = L.PredParams()
        pp.B, pp.Cin, pp.Cout, pp.H, pp.W = B, 32, 2, H, W
        pp.x_c8, pp.w, pp.b, pp.y = L.ptr(h), L.ptr(w), L.ptr(b), L.ptr(flow)
        L.call("ef_pred_fwd", pp)
        ctx.model, ctx.saved, ctx.flow, ctx.first, ctx.z_last = model, saved, flow, token is None, h
        ctx.carry = fs.carry
        ctx.shapes = (B, Cin0, H, W)
        model._last_spikes = zs
        new_token = torch.zeros((), device=dev, dtype=torch.float32)
        return flow, new_token

    @staticmethod
    def backward(ctx, g_flow, g_token):
        model, carry = ctx.model, ctx.carry
        B, Cin0, H, W = ctx.shapes
        params = _params_of(model)
        dev = ctx.flow.device
        flat = torch.zeros(sum(p.numel() for p in params), device=dev, dtype=torch.float32)
        grads, o = [], 0
        for p in params:
            grads.append(flat[o:o + p.numel()].view(p.shape))
            o += p.numel()
        gi = len(grads) - 2
        # prediction head
        g_h = torch.empty((B, 32, H, W), device=dev, dtype=torch.float32)
        pp = L.PredParams()
        pp.B, pp.Cin, pp.Cout, pp.H, pp.W = B, 32, 2, H, W
        z7 = ctx.z_last
        w, b = model.pred.conv2d.weight.detach(), model.pred.conv2d.bias.detach()
        if g_flow is None:
            g_flow = torch.zeros_like(ctx.flow)
        pp.x_c8, pp.w, pp.b, pp.y, pp.g_y = L.ptr(z7), L.ptr(w), L.ptr(b), L.ptr(ctx.flow), L.ptr(g_flow.contiguous())
        pp.g_x, pp.g_w, pp.g_b = L.ptr(g_h), L.ptr(grads[gi]), L.ptr(grads[gi + 1])
        L.call("ef_pred_bwd", pp)
        # cells, last to first
        for i in reversed(range(len(LAYERS))):
            cell = getattr(model, LAYERS[i])
            x_f32, x_c8, v_in, z_in, v_out = ctx.saved[i]
            gi -= 4 if cell.recurrent else 3
            leak, thresh = cell.leak.detach().reshape(-1), cell.thresh.detach().reshape(-1)
            q = L.LifConvBwdParams()
            _fill_fwd(q.f, B, Cin0 if i == 0 else 32, H, W, cell, x_f32, x_c8, v_in, z_in, v_out, leak, thresh)
            q.g_out, q.g_v_out, q.g_z_out = L.ptr(g_h), L.ptr(carry.g_v[i]), L.ptr(carry.g_z[i])
            scratch = torch.empty((B, 32, H, W), device=dev, dtype=torch.float32)
            q.scratch_gI = L.ptr(scratch)
            g_x = torch.empty((B, 32, H, W), device=dev, dtype=torch.float32) if i > 0 else None
            q.g_x = L.ptr(g_x)
            g_v_in = g_z_in = None
            if not ctx.first and v_in is not None:
                g_v_in = torch.empty((B, 32, H, W), device=dev, dtype=torch.float32)
                q.g_v_in = L.ptr(g_v_in)
                if cell.recurrent:
                    g_z_in = torch.empty((B, 32, H, W), device=dev, dtype=torch.float32)
                    q.g_z_in = L.ptr(g_z_in)
            k = gi
            q.g_w_ff = L.ptr(grads[k])
            k += 1
            if cell.recurrent:
                q.g_w_rec = L.ptr(grads[k])
                k += 1
            q.g_leak, q.g_thresh = L.ptr(grads[k]), L.ptr(grads[k + 1])
            L.call("ef_lif_conv_bwd", q)
            carry.g_v[i], carry.g_z[i] = g_v_in, g_z_in
            g_h = g_x
        out = []
        for p, g in zip(params, grads):
            out.append(g if p.requires_grad else None)
        g_tok = None if ctx.first else torch.zeros((), device=dev, dtype=torch.float32)
        return (None, None, g_tok, *out)


def forward(model, x, log=False):
    """One forward pass of a LIF FireNet on the fast path.  Returns the reference's output dict."""
    fs = model._fast
    if fs is None:
        fs = FastState(len(LAYERS))
        for i, s in enumerate(model._states):  # states set through the reference-format API are converted once
            if s is not None:
                fs.v[i] = s[0].detach().contiguous()
                fs.z[i] = ops.pack_c8(s[1].detach())
        model._fast = fs
    params = _params_of(model)
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    if need_grad:
        token = fs.token
        flow, fs.token = _FireNetStep.apply(model, x, token, *params)
    else:
        with torch.no_grad():
            flow, _ = _FireNetStep.forward(_NoCtx(), model, x, None, *params)
        fs.detach()
    activity = None
    if log:
        names = ["0:input", "1:head", "2:G1", "3:R1a", "4:R1b", "5:G2", "6:R2a", "7:R2b", "8:pred"]
        vals = [x] + list(model._last_spikes) + [flow]
        activity = {n: t.detach().ne(0).float().mean().item() for n, t in zip(names, vals)}
    return {"flow": [flow], "activity": activity}


class _NoCtx:
    """Stand-in for the autograd context when the step runs without gradient tracking."""


def states_of(model):
    """Reference-format view of the internal state: list of stacked [2,B,C,H,W] tensors (fresh tensors = clones)."""
    fs = model._fast
    out = []
    for v, z in zip(fs.v, fs.z):
        out.append(None if v is None else torch.stack([v, ops.unpack_c8(z)]))
    return out
