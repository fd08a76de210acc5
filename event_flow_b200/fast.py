"""
Fast path of the spiking FireNet chain (models/model.py:254-265) on the internal formats.

Between the cells, spikes never exist as fp32 NCHW tensors: every cell writes bf16 channels-last spikes ("cl") that the
next cell's tcgen05 kernel consumes through TMA; membrane potentials stay fp32 NCHW (the reference's state format).

Activations of a BPTT window live in a per-model arena, LAYER-MAJOR and STEP-CONTIGUOUS: for every layer one tensor
[steps, B, ...] of membrane potentials and one of spikes.  That is what makes the backward cheap: the reference (and round 1 of
this package) back-propagates step by step; here the backward of a window is DEFERRED to the window's first step (the last
autograd node to run) and executed layer by layer over the whole window:
  * feed-forward cells: ONE time-fused neuron-backward launch (dL/dv stays in registers over the T steps, every membrane tensor
    is read once), ONE tensor-core data-gradient launch and ONE tensor-core weight-gradient launch over T*B images
    (ef_lif_bwd_window);
  * recurrent cells (G1, G2): pointwise + data gradient step by step (the recurrent path needs dL/dz of step t+1 first), the
    weight gradient of the whole window in one batched launch (ef_lif_bwd_tc + ef_lif_wgrad_tc);
  * prediction head: one launch over T*B images.
One torch.autograd node per MODEL step only orders the steps (scalar token) and collects dL/dflow.
"""
import gc

import torch

from . import _lib as L
from . import ops

LAYERS = ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b")
N_L = len(LAYERS)
DEFAULT_WINDOW_CAP = 12  # steps a bank holds before it has to grow (train_SNN.yml: window_loss / window = 10)


def _capture(fn):
    """
    fn() captured into a CUDA graph.  thread_local error mode: other threads (the NCCL watchdog under data parallelism) keep issuing
    CUDA calls during a capture.  The cyclic garbage collector is paused for the duration: a collection that happens to run inside the
    capture may destroy CUDA objects of earlier models (graphs with their memory pools, tensors), i.e. issue calls that are illegal on a
    capturing thread and invalidate the capture.
    """
    g = torch.cuda.CUDAGraph()
    was_enabled = gc.isenabled()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):  # (its __enter__ runs a full collection first)
        gc.disable()
        try:
            fn()
        finally:
            if was_enabled:
                gc.enable()
    return g


class _Carry:
    """What the steps of one BPTT window hand to the deferred window backward."""

    def __init__(self):
        self.g_flows = {}     # step index -> dL/dflow of that step
        self.n = 0            # steps of the window so far
        self.prev = None      # (v, z) lists: the state the window started from (detached), entries may be None
        self.parity = None    # arena bank of the window
        self.head_tc = False  # the head layer ran on the tensor cores (split input in bank.x_cl)
        self.flow_y = None    # window entry point: the flows [T,B,2,H,W] as computed (tanh outputs the prediction backward needs)
        self.into_sink = False  # the window backward accumulated the parameter gradients straight into the trainer's flat buffer


class _Slot:
    """One model step inside a bank: views of the step's activations, its CUDA graphs and its generation counter."""

    def __init__(self, bank, idx):
        self.gen = 0       # bumped every time a forward step (re)writes this slot: saved activations of older steps are then gone
        self.graphs = {}   # (pointer signature) -> torch.cuda.CUDAGraph of this step's 8 kernels
        self.bind(bank, idx)

    def bind(self, bank, idx):
        self.v = [bank.v[i][idx] for i in range(N_L)]
        self.z = [bank.zs[i][idx + 1] for i in range(N_L)]
        self.flow = bank.flow[idx]
        self.x_in = None if bank.x_in is None else bank.x_in[idx]  # fp32 copy of the model input (legacy head kernels only)
        self.x_cl = None if bank.x_cl is None else bank.x_cl[idx]  # the model input as exact bf16 hi/mid/lo split, channels-last
        self.graphs = {}


class _Bank:
    """
    Activations of one BPTT window: per layer membrane fp32 [cap,B,32,H,W] and spikes bf16 [cap+1,B,H,W,32] (index 0 is reserved for
    the spikes BEFORE the window's first step, so that "previous spikes of steps 0..T-1" is one dense tensor for the batched
    weight gradient of the recurrent cells), the flow maps [cap,B,2,H,W] and the model inputs [cap,B,Cin,H,W].
    """

    def __init__(self, key, cap):
        B, H, W, dev = key
        self.key, self.cap = key, cap
        self.v = [torch.empty((cap, B, 32, H, W), device=dev, dtype=torch.float32) for _ in range(N_L)]
        self.zs = [torch.empty((cap + 1, B, H, W, 32), device=dev, dtype=torch.bfloat16) for _ in range(N_L)]
        self.flow = torch.empty((cap, B, 2, H, W), device=dev, dtype=torch.float32)
        self.x_in = None   # [cap,B,Cin,H,W] fp32, allocated on demand
        self.x_cl = None   # [cap,B,H,W,32] bf16 split input of the head layer (Cin <= 10), allocated on demand
        self.slots = []

    def slot(self, idx):
        while len(self.slots) <= idx:
            self.slots.append(_Slot(self, len(self.slots)))
        return self.slots[idx]

    def need_input(self, cin, f32, split):
        B, H, W, dev = self.key
        changed = False
        if f32 and (self.x_in is None or self.x_in.shape[2] != cin):
            self.x_in, changed = torch.empty((self.cap, B, cin, H, W), device=dev, dtype=torch.float32), True
        if split and self.x_cl is None:
            self.x_cl, changed = torch.empty((self.cap, B, H, W, 32), device=dev, dtype=torch.bfloat16), True
        if changed:
            for i, s in enumerate(self.slots):
                s.bind(self, i)

    def grow(self, n_used):
        """A window longer than the bank: double the capacity, keep the steps recorded so far (graphs are re-captured)."""
        new = _Bank(self.key, 2 * self.cap)
        for i in range(N_L):
            new.v[i][:n_used].copy_(self.v[i][:n_used])
            new.zs[i][:n_used + 1].copy_(self.zs[i][:n_used + 1])
        new.flow[:n_used].copy_(self.flow[:n_used])
        new.need_input(0 if self.x_in is None else self.x_in.shape[2], self.x_in is not None, self.x_cl is not None)
        if self.x_in is not None:
            new.x_in[:n_used].copy_(self.x_in[:n_used])
        if self.x_cl is not None:
            new.x_cl[:n_used].copy_(self.x_cl[:n_used])
        self.cap, self.v, self.zs, self.flow, self.x_in, self.x_cl = new.cap, new.v, new.zs, new.flow, new.x_in, new.x_cl
        for i, s in enumerate(self.slots):
            s.bind(self, i)


class _Arena:
    """
    Activation storage owned by the model and reused window after window (no allocator traffic in steady state, static
    addresses for CUDA-graph replay).  Two banks alternate per BPTT window: window w writes bank w%2, its initial state
    lives in the last used slot of bank (w-1)%2.  CONTRACT: loss.backward() of window w must run before window w+1 COMPLETES
    (window w+1 overwrites, at its own last step, the slot that holds window w's initial state; window w+2 overwrites
    window w's activations) -- train_flow.py:154-171 runs backward right at the window end.  The contract is enforced:
    every slot carries a generation counter, a step's backward raises if a slot it saved from has been rewritten since.
    """

    def __init__(self, B, H, W, dev, cap):
        self.key = (B, H, W, dev)
        self.cap = cap
        self.banks = [None, None]
        self.parity = 0
        self.flat = None   # fp32 [n_params]: parameter gradients of the window being back-propagated
        self.bwd = None    # buffers of the window backward

    def bank(self, parity):
        if self.banks[parity] is None:
            self.banks[parity] = _Bank(self.key, self.cap)
        return self.banks[parity]

    def window_buffers(self, cap):
        """Gradient buffers of the window backward, sized for `cap` steps."""
        if self.bwd is None or self.bwd["cap"] < cap:
            B, H, W, dev = self.key
            f32 = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)  # noqa: E731
            bf = lambda *s: torch.empty(s, device=dev, dtype=torch.bfloat16)  # noqa: E731
            self.bwd = {
                "cap": cap,
                "g_h": [f32(cap, B, 32, H, W), f32(cap, B, 32, H, W)],    # dL/d(spikes) of all steps, ping-pong between neighbouring layers
                "gI_hi": bf(cap, B, H, W, 32), "gI_mid": bf(cap, B, H, W, 32),
                "g_flow": f32(cap, B, 2, H, W),
                "g_v": [f32(B, 32, H, W), f32(B, 32, H, W)], "g_z": [f32(B, 32, H, W), f32(B, 32, H, W)],  # recurrent cells, step to step
                "wg": f32(L.lib().ef_lif_wgrad_partial_elems(cap * B, H, W, 1)),
            }
        return self.bwd


class FastState:
    """Per-window internal state of a model on the fast path."""

    def __init__(self, n):
        self.v = [None] * n        # fp32 [B,C,H,W]
        self.z = [None] * n        # bf16 channels-last [B,H,W,C]
        self.token = None          # scalar autograd token ordering the steps of one BPTT window
        self.carry = _Carry()
        self.step = 0              # index of the next step inside the current window
        self.param_sig = None      # data pointers of the parameters / weight images, part of the graph-cache key
        self.src = None            # (slot, generation) that holds the current state tensors (None: set through the state API)

    def detach(self, arena=None):
        self.token = None
        self.carry = _Carry()
        self.step = 0
        if arena is not None:
            arena.parity ^= 1


def _cells(model):
    """The 7 cells in chain order, looked up once (nn.Module.__getattr__ is slow on the per-step path)."""
    c = model.__dict__.get("_fast_cells")
    if c is None:
        c = model.__dict__["_fast_cells"] = [getattr(model, n) for n in LAYERS]
        for cell in c:
            if "act_width" in cell._buffers:  # spiking cells only (eligible() also inspects the ANN FireNets)
                cell.__dict__["_act_width_f"] = float(cell.act_width)
    return c


def eligible(model, x):
    """LIF cells, 32 channels, 3x3, stride 1, CUDA, W % 4 == 0 (TMA stride rule for the fp32 membrane tensor)."""
    if not x.is_cuda:
        return False
    ok = model.__dict__.get("_fast_eligible")
    if ok is None:
        ok = not model.residual
        for name, cell in zip(LAYERS, _cells(model)):
            if not ok:
                break
            if getattr(cell, "neuron", None) != "lif" or cell.hidden_size != 32 or cell.ff.kernel_size != (3, 3) or cell.stride != 1:
                ok = False
            elif not cell.plain():  # normalisation options / differentiable reset: the cell path
                ok = False
            elif name != "head" and cell.input_size != 32:
                ok = False
        model.__dict__["_fast_eligible"] = ok
    return ok and x.shape[-1] % 4 == 0


def _split_cache(model):
    """
    bf16 hi/mid/lo weight images of the hidden layers.  The buffers are allocated once (stable addresses for CUDA-graph
    replay) and re-filled in place when a weight tensor changed.
    """
    cache = model.__dict__.setdefault("_w_split_cache", {})
    cells = _cells(model)
    # steady state (weights unchanged since the last call): one tuple comparison
    sig = tuple(c.ff.weight._version for c in cells) + tuple(c.rec.weight._version for c in cells[1:] if c.recurrent) + (
        model.__dict__.get("_w_epoch", 0), cells[1].ff.weight.data_ptr())
    last = cache.get("__last__")
    if last is not None and last[0] == sig:
        return last[1]
    out = {}
    head = cells[0]
    if head.input_size <= L.EF_HEAD_MAX_CIN:  # head layer on the tensor cores: weight image for split inputs
        key = (head.ff.weight._version, head.ff.weight.data_ptr(), model.__dict__.get("_w_epoch", 0))
        hit = cache.get("head")
        if hit is None or hit[0] != key:
            same_dev = hit is not None and hit[1].device == head.ff.weight.device
            hit = (key, ops.split_weights_head(head.ff.weight, out=hit[1] if same_dev else None))
            cache["head"] = hit
        out["head"] = hit[1]
    for name, cell in zip(LAYERS[1:], cells[1:]):
        rec = cell.rec.weight if cell.recurrent else None
        key = (cell.ff.weight._version, cell.ff.weight.data_ptr(), None if rec is None else rec._version, model.__dict__.get("_w_epoch", 0))
        hit = cache.get(name)
        if hit is None or hit[0] != key:
            same_dev = hit is not None and hit[1].device == cell.ff.weight.device
            hit = (key, ops.split_weights(cell.ff.weight, rec, out=hit[1] if same_dev else None),
                   ops.split_weights_bwd(cell.ff.weight, rec, out=hit[2] if same_dev else None))
            cache[name] = hit
        out[name] = hit[1]
        out[name + ".bwd"] = hit[2]
    cache["__last__"] = (sig, out)
    return out


def invalidate_weights(model):
    """Call after updating parameters outside torch's version tracking (e.g. the fused Adam kernel)."""
    model.__dict__["_w_epoch"] = model.__dict__.get("_w_epoch", 0) + 1


def invalidate_pointers(model):
    """Call after re-binding parameter storage (p.data = ...): cached graphs / argument structs hold the old addresses."""
    for k in ("_fast_params", "_w_split_cache", "_arena"):
        model.__dict__.pop(k, None)
    if model.__dict__.get("_fast") is not None:
        model._fast.param_sig = None
        model._fast.src = None


def _params_of(model):
    ps = model.__dict__.get("_fast_params")
    if ps is not None:
        return ps
    ps = []
    for cell in _cells(model):
        ps.append(cell.ff.weight)
        if cell.recurrent:
            ps.append(cell.rec.weight)
        ps.append(cell.leak)
        ps.append(cell.thresh)
    ps.append(model.pred.conv2d.weight)
    ps.append(model.pred.conv2d.bias)
    model.__dict__["_fast_params"] = ps
    return ps


def _fill_fwd(p, B, Cin, H, W, cell, x_f32, x_cl, v_in, z_in, v_out, leak, thresh):
    p.B, p.Cin, p.C, p.H, p.W = B, Cin, 32, H, W
    p.ksize, p.stride, p.neuron, p.hard_reset = 3, 1, L.EF_LIF, int(cell.hard_reset)
    p.surrogate, p.act_width = L.SURROGATE_CODES[cell.activation], float(cell._act_width_f)
    p.x, p.x_cl = L.ptr(x_f32), L.ptr(x_cl)
    p.v_in, p.z_in_cl = L.ptr(v_in), L.ptr(z_in)
    p.w_ff = L.ptr(cell.ff.weight)
    p.w_rec = L.ptr(cell.rec.weight) if cell.recurrent else None
    p.leak, p.thresh = L.ptr(leak), L.ptr(thresh)
    p.v_out = L.ptr(v_out)


def _launch_step(model, x, v_in, z_in, slot, splits, B, Cin0, H, W, only_hidden=False):
    """
    The 8 kernels of one model step: head, 6 tensor-core cells, prediction head.  All tensors are caller-provided.
    x: the fp32 network input (legacy CUDA-core head kernel) or None = the head runs on the tensor cores from slot.x_cl.
    """
    h = None
    cells = _cells(model)
    for i, name in enumerate(LAYERS):
        if only_hidden and i == 0:  # measurement replays (bench.py): the head's spikes are already in the slot
            h = slot.z[0]
            continue
        cell = cells[i]
        leak, thresh = cell.leak.detach().reshape(-1), cell.thresh.detach().reshape(-1)
        p = L.LifConvParams()
        if i == 0 and x is None:  # split input [B,H,W,32]: the 32 -> 32 tensor-core kernel with the head's weight image
            _fill_fwd(p, B, 32, H, W, cell, None, slot.x_cl, v_in[i], z_in[i], slot.v[i], leak, thresh)
        else:
            _fill_fwd(p, B, Cin0 if i == 0 else 32, H, W, cell, x if i == 0 else None, h, v_in[i], z_in[i], slot.v[i], leak, thresh)
        p.z_out_cl = L.ptr(slot.z[i])
        if i > 0 or x is None:
            p.w_split = L.ptr(splits[name])
        L.call("ef_lif_conv_fwd", p, tag=(p.Cin, 32, cell.recurrent))
        h = slot.z[i]
    if only_hidden:
        return
    w, b = model.pred.conv2d.weight.detach(), model.pred.conv2d.bias.detach()
    pp = L.PredParams()
    pp.B, pp.Cin, pp.Cout, pp.H, pp.W = B, 32, 2, H, W
    pp.x_cl, pp.w, pp.b, pp.y = L.ptr(h), L.ptr(w), L.ptr(b), L.ptr(slot.flow)
    L.call("ef_pred_fwd", pp)


def capture_window(model, xs, only_hidden=False):
    """
    Measurement aid (bench.py roofline): the kernels of len(xs) - 1 consecutive model steps on a private activation bank,
    captured as ONE CUDA graph (step 0 runs eagerly from the zero state and provides the previous state of step 1).
    With only_hidden the graph holds just the six 32 -> 32 tensor-core cell launches of every step.  Replaying it repeats
    the same computation on the same operands (every launch streams its own ~59 MB, the whole replay far more than L2).
    Returns (graph, kernel launches per replay).
    """
    x0 = xs[0]
    B, Cin0, H, W = x0.shape
    _cells(model)
    splits = _split_cache(model)
    bank = _Bank((B, H, W, x0.device), len(xs))
    head_tc = "head" in splits
    bank.need_input(Cin0, False, head_tc)
    slots = [bank.slot(t) for t in range(len(xs))]
    none = [None] * N_L
    xs = [x.contiguous() for x in xs]
    if head_tc:
        for t, x in enumerate(xs):
            ops.pack_split_cl(x, out=slots[t].x_cl)
        xs = [None] * len(xs)
    with torch.no_grad():
        _launch_step(model, xs[0], none, none, slots[0], splits, B, Cin0, H, W)
        for t in range(1, len(xs)):  # eager pass: fills every slot (the hidden-only replay needs the head spikes in place)
            _launch_step(model, xs[t], slots[t - 1].v, slots[t - 1].z, slots[t], splits, B, Cin0, H, W)
        torch.cuda.synchronize()
        def replayed():
            for t in range(1, len(xs)):
                _launch_step(model, xs[t], slots[t - 1].v, slots[t - 1].z, slots[t], splits, B, Cin0, H, W, only_hidden=only_hidden)

        g = _capture(replayed)
    g._keepalive = (bank, xs, splits)
    return g, (len(xs) - 1) * (N_L - 1 if only_hidden else N_L + 1)


class _FireNetStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, token, *params):
        fs = model._fast
        x = x.contiguous()
        B, Cin0, H, W = x.shape
        dev = x.device
        _cells(model)
        splits = _split_cache(model)
        arena = model.__dict__.get("_arena")
        if arena is None or arena.key != (B, H, W, dev):
            arena = model.__dict__["_arena"] = _Arena(B, H, W, dev, int(model.__dict__.get("_window_cap", DEFAULT_WINDOW_CAP)))
        parity = arena.parity
        bank = arena.bank(parity)
        if fs.step >= bank.cap:
            bank.grow(fs.step)
        # head layer: tensor cores on the exact bf16 split of the input (Cin <= 10); the fp32 copy is only kept for the legacy /
        # step-by-step backward paths
        head_tc = "head" in splits and model.__dict__.get("_tc_head", True)
        keep_f32 = not head_tc or not (model.__dict__.get("_window_backward", True) and model.__dict__.get("_tc_backward", True))
        bank.need_input(Cin0, keep_f32, head_tc)
        idx = fs.step
        slot = bank.slot(idx)
        fs.step += 1
        slot.gen += 1
        v_in, z_in = list(fs.v), list(fs.z)
        ctx.guards = [(slot, slot.gen)] + ([fs.src] if fs.src is not None and fs.src[0] is not slot else [])
        fs.src = (slot, slot.gen)
        carry = fs.carry
        if idx == 0:
            carry.prev, carry.parity = (v_in, z_in), parity
        carry.n = idx + 1
        cap = model.__dict__.get("_capture")
        # the input moves to a fixed address (CUDA-graph replay of the step; the window-wide weight gradient of the head layer)
        if keep_f32:
            slot.x_in.copy_(x)
        if head_tc:
            ops.pack_split_cl(x, out=slot.x_cl)
        x_step = None if head_tc else slot.x_in
        carry.head_tc = head_tc
        use_graph = (model.__dict__.get("_use_graphs", True) and cap is None and L.PROFILE is None
                     and not torch.cuda.is_current_stream_capturing())
        if use_graph:
            if fs.param_sig is None:  # refreshed per sequence (reset_states) and whenever the module is moved (FireNet._apply)
                fs.param_sig = (tuple(p.data_ptr() for p in _params_of(model)), tuple(splits[n].data_ptr() for n in LAYERS if n in splits))
            key = (fs.param_sig, tuple(0 if v is None else v.data_ptr() for v in v_in), (slot.x_cl if head_tc else slot.x_in).data_ptr())
            g = slot.graphs.get(key)
            if g is None:
                _launch_step(model, x_step, v_in, z_in, slot, splits, B, Cin0, H, W)  # eager: results + lazy init
                g = slot.graphs[key] = _capture(lambda: _launch_step(model, x_step, v_in, z_in, slot, splits, B, Cin0, H, W))
            else:
                g.replay()
                L.GRAPH_KERNELS += N_L + 1
        else:
            _launch_step(model, x_step, v_in, z_in, slot, splits, B, Cin0, H, W)
        for i, name in enumerate(LAYERS):
            if cap is not None:  # test hook: what this layer consumed and produced, in the reference's tensor format
                xin = x if i == 0 else ops.unpack_cl(slot.z[i - 1])
                sin = None if v_in[i] is None else torch.stack([v_in[i], ops.unpack_cl(z_in[i])]).cpu()
                zo = ops.unpack_cl(slot.z[i])
                cap[name] = (xin.detach().cpu(), sin, zo.cpu(), torch.stack([slot.v[i], zo]).cpu())
            fs.v[i], fs.z[i] = slot.v[i], slot.z[i]
        flow = slot.flow.clone()  # the caller may keep the flow for as long as it likes; the slot is recycled
        ctx.model, ctx.arena, ctx.carry, ctx.idx, ctx.first = model, arena, carry, idx, token is None
        ctx.shapes = (B, Cin0, H, W)
        model._last_spikes = slot.z
        if isinstance(ctx, _NoCtx):
            return flow, None
        new_token = torch.zeros((), device=dev, dtype=torch.float32)
        return flow, new_token

    @staticmethod
    def backward(ctx, g_flow, g_token):
        model, carry = ctx.model, ctx.carry
        for slot_, gen_ in ctx.guards:
            if slot_.gen != gen_:
                raise RuntimeError(
                    "event_flow_b200 fast path: the activations this backward step needs were overwritten by a later forward pass. "
                    "loss.backward() of a BPTT window must run before the next window completes (see fast._Arena).")
        if g_flow is not None:
            carry.g_flows[ctx.idx] = g_flow
        dev = ctx.arena.key[3]
        if not ctx.first:  # nothing is computed yet: the window's first step (the last node to run) back-propagates the whole window
            return (None, None, torch.zeros((), device=dev, dtype=torch.float32))
        params = _params_of(model)
        carry.into_sink = False
        if model.__dict__.get("_window_backward", True) and model.__dict__.get("_tc_backward", True):
            grads = _window_backward(model, ctx.arena, carry, ctx.shapes)
        else:
            grads = _stepwise_backward(model, ctx.arena, carry, ctx.shapes)
        carry.g_flows = {}
        if carry.into_sink:  # already accumulated into p.grad (views of the trainer's flat buffer)
            return (None, None, None, *[None] * len(params))
        return (None, None, None, *[g.clone() if p.requires_grad else None for p, g in zip(params, grads)])


def _grad_sink_views(model, params):
    """
    Gradient sink: when every parameter's .grad is a dense fp32 view into the flat gradient buffer of a DataParallelTrainer
    (model._grad_sink), the window backward accumulates straight into those views -- every gradient kernel already accumulates (+=),
    which is exactly autograd's `p.grad += g` -- and hands autograd no parameter gradients: this replaces a clone and an add kernel per
    parameter (2 x 25 launches per window).  Returns the views in the parameter order of this path, or None when the layout does not hold
    (gradients set to None, re-bound by autograd, a foreign buffer, ...).
    """
    sink = model.__dict__.get("_grad_sink")
    if sink is None:
        return None
    lo, hi = sink.data_ptr(), sink.data_ptr() + 4 * sink.numel()
    views = []
    for q in params:
        g = q.grad
        if (not q.requires_grad or g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != q.shape
                or g.data_ptr() < lo or g.data_ptr() + 4 * g.numel() > hi):
            return None
        views.append(g)
    return views


def _flat_grads(arena, params, dev):
    n = sum(p.numel() for p in params)
    if arena.flat is None or arena.flat.numel() != n:
        arena.flat = torch.zeros(n, device=dev, dtype=torch.float32)
    else:
        arena.flat.zero_()
    grads, o = [], 0
    for p in params:
        grads.append(arena.flat[o:o + p.numel()].view(p.shape))
        o += p.numel()
    return grads


def _grad_slots(cells, grads):
    """Per layer the views of the flat gradient buffer: (g_w_ff, g_w_rec | None, g_leak, g_thresh); then pred (g_w, g_b)."""
    out, k = [], 0
    for cell in cells:
        g_ff = grads[k]
        k += 1
        g_rec = None
        if cell.recurrent:
            g_rec = grads[k]
            k += 1
        out.append((g_ff, g_rec, grads[k], grads[k + 1]))
        k += 2
    out.append((grads[k], grads[k + 1]))
    return out


def _window_g_flow(carry, Tn, B, H, W, buf):
    """dL/dflow of all steps as one dense [Tn,B,2,H,W] tensor: the loss kernel's gradient slab itself when the steps' gradients are its
    consecutive slices (the usual case), a gathered copy otherwise (flows used elsewhere, missing steps)."""
    gs = [carry.g_flows.get(t) for t in range(Tn)]
    n = B * 2 * H * W
    g0 = gs[0]
    if g0 is not None and all(g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.shape == (B, 2, H, W)
                              and g.data_ptr() == g0.data_ptr() + 4 * n * t for t, g in enumerate(gs)):
        try:
            return g0.as_strided((Tn, B, 2, H, W), (n, 2 * H * W, H * W, W, 1))
        except RuntimeError:
            pass
    out = buf["g_flow"][:Tn]
    for t, g in enumerate(gs):
        if g is None:
            out[t].zero_()
        else:
            out[t].copy_(g)
    return out


def _window_backward(model, arena, carry, shapes):
    """BPTT of one whole window, layer by layer (see the module docstring).  Returns the parameter gradients (views of arena.flat)."""
    B, Cin0, H, W = shapes
    dev = arena.key[3]
    bank = arena.banks[carry.parity]
    Tn = carry.n
    cells, params, splits = _cells(model), _params_of(model), _split_cache(model)
    grads = _grad_sink_views(model, params)
    carry.into_sink = grads is not None
    if grads is None:
        grads = _flat_grads(arena, params, dev)
    gs = _grad_slots(cells, grads)
    buf = arena.window_buffers(bank.cap)
    v0, z0 = carry.prev
    # prediction head, all steps at once
    g_flow = _window_g_flow(carry, Tn, B, H, W, buf)
    cur = 0
    pp = L.PredParams()
    pp.B, pp.Cin, pp.Cout, pp.H, pp.W = Tn * B, 32, 2, H, W
    w, b = model.pred.conv2d.weight.detach(), model.pred.conv2d.bias.detach()
    pp.x_cl, pp.w, pp.b, pp.y, pp.g_y = L.ptr(bank.zs[N_L - 1][1:Tn + 1]), L.ptr(w), L.ptr(b), L.ptr(bank.flow[:Tn]), L.ptr(g_flow)
    pp.g_x, pp.g_w, pp.g_b = L.ptr(buf["g_h"][cur][:Tn]), L.ptr(gs[N_L][0]), L.ptr(gs[N_L][1])
    L.call("ef_pred_bwd", pp)
    for i in reversed(range(N_L)):
        cell = cells[i]
        g_ff, g_rec, g_leak, g_thresh = gs[i]
        leak, thresh = cell.leak.detach().reshape(-1), cell.thresh.detach().reshape(-1)
        surr, width, hard = L.SURROGATE_CODES[cell.activation], float(cell._act_width_f), int(cell.hard_reset)
        g_out, g_x = buf["g_h"][cur], buf["g_h"][cur ^ 1]
        if not cell.recurrent:  # head and feed-forward cells: three launches for the whole window
            q = L.LifBwdWindowParams()
            q.B, q.T, q.H, q.W, q.hard_reset, q.surrogate, q.act_width = B, Tn, H, W, hard, surr, width
            q.z_cl, q.z_prev_cl = L.ptr(bank.zs[i][1:Tn + 1]), L.ptr(z0[i])
            q.v, q.v_prev, q.g_out = L.ptr(bank.v[i][:Tn]), L.ptr(v0[i]), L.ptr(g_out[:Tn])
            q.leak, q.thresh = L.ptr(leak), L.ptr(thresh)
            q.g_w_ff, q.g_leak, q.g_thresh = L.ptr(g_ff), L.ptr(g_leak), L.ptr(g_thresh)
            if i == 0 and carry.head_tc:  # split-input head: tensor-core weight gradient, no data gradient
                q.Cin, q.x_cl = Cin0, L.ptr(bank.x_cl[:Tn])
                q.gI_hi, q.gI_mid, q.wg_partial = L.ptr(buf["gI_hi"][:Tn]), L.ptr(buf["gI_mid"][:Tn]), L.ptr(buf["wg"])
            elif i == 0:
                q.Cin, q.x_f32, q.gI_f32 = Cin0, L.ptr(bank.x_in[:Tn]), L.ptr(g_x[:Tn])  # (the other ping-pong buffer is free: nothing below the head)
            else:
                q.x_cl, q.w_bwd = L.ptr(bank.zs[i - 1][1:Tn + 1]), L.ptr(splits[LAYERS[i] + ".bwd"])
                q.gI_hi, q.gI_mid = L.ptr(buf["gI_hi"][:Tn]), L.ptr(buf["gI_mid"][:Tn])
                q.g_x, q.wg_partial = L.ptr(g_x[:Tn]), L.ptr(buf["wg"])
            L.call("ef_lif_bwd_window", q)
        else:  # recurrent cells: dL/dz of step t+1 reaches step t through the recurrent convolution, so pointwise + data gradient go step by step
            if z0[i] is not None:
                bank.zs[i][0].copy_(z0[i])  # previous spikes of steps 0..Tn-1 become ONE dense tensor for the batched weight gradient
            else:
                bank.zs[i][0].zero_()
            g_v_next = g_z_next = None
            for t in reversed(range(Tn)):
                par = t & 1
                has_prev = t > 0 or v0[i] is not None
                q = L.LifBwdTcParams()
                q.B, q.H, q.W, q.has_rec, q.hard_reset, q.surrogate, q.act_width = B, H, W, 1, hard, surr, width
                q.x_cl = L.ptr(bank.zs[i - 1][t + 1])
                q.z_in_cl = L.ptr(bank.zs[i][t]) if has_prev else None
                q.v_in = L.ptr(bank.v[i][t - 1] if t > 0 else v0[i])
                q.v_out = L.ptr(bank.v[i][t])
                q.g_out, q.g_v_out, q.g_z_out = L.ptr(g_out[t]), L.ptr(g_v_next), L.ptr(g_z_next)
                q.leak, q.thresh, q.w_bwd = L.ptr(leak), L.ptr(thresh), L.ptr(splits[LAYERS[i] + ".bwd"])
                q.gI_hi, q.gI_mid = L.ptr(buf["gI_hi"][t]), L.ptr(buf["gI_mid"][t])
                q.g_x = L.ptr(g_x[t])
                g_v_next = g_z_next = None
                if t > 0:  # (the state before step 0 is detached: nobody consumes its gradient)
                    g_v_next, g_z_next = buf["g_v"][par], buf["g_z"][par]
                    q.g_v_in, q.g_z_in = L.ptr(g_v_next), L.ptr(g_z_next)
                q.g_leak, q.g_thresh = L.ptr(g_leak), L.ptr(g_thresh)
                L.call("ef_lif_bwd_tc", q)  # no weight-gradient pointers: pointwise + data gradient only
            L.LAUNCHES += 1
            L.check(L.lib().ef_lif_wgrad_tc(L.ptr(bank.zs[i - 1][1:Tn + 1]), L.ptr(bank.zs[i][:Tn]), L.ptr(buf["gI_hi"][:Tn]), L.ptr(buf["gI_mid"][:Tn]),
                                            1, Tn * B, H, W, L.ptr(buf["wg"]), L.EF_WG_FINALIZE, L.ptr(g_ff), L.ptr(g_rec), L.stream()),
                    "ef_lif_wgrad_tc")
        cur ^= 1
    return grads


def _stepwise_backward(model, arena, carry, shapes):
    """
    Fallback / A-B reference (model._window_backward = False or model._tc_backward = False): the same BPTT step by step, one cell-step
    per library call, state gradients handed from step t+1 to step t in fp32 buffers.  With _tc_backward = False every cell runs on the
    generic CUDA-core backward (ef_lif_conv_bwd).
    """
    B, Cin0, H, W = shapes
    dev = arena.key[3]
    bank = arena.banks[carry.parity]
    Tn = carry.n
    cells, params, splits = _cells(model), _params_of(model), _split_cache(model)
    grads = _flat_grads(arena, params, dev)
    gs = _grad_slots(cells, grads)
    tc = model.__dict__.get("_tc_backward", True)
    mk = lambda: torch.empty((B, 32, H, W), device=dev, dtype=torch.float32)  # noqa: E731
    g_hh = [mk(), mk()]
    g_vb = [[mk() for _ in range(N_L)] for _ in range(2)]
    g_zb = [[mk() for _ in range(N_L)] for _ in range(2)]
    scratch = mk()
    gI_hi = torch.empty((B, H, W, 32), device=dev, dtype=torch.bfloat16)
    gI_mid = torch.empty_like(gI_hi)
    v0, z0 = carry.prev
    g_v, g_z = [None] * N_L, [None] * N_L
    w, b = model.pred.conv2d.weight.detach(), model.pred.conv2d.bias.detach()
    for t in reversed(range(Tn)):
        par = t & 1
        g_flow = carry.g_flows.get(t)
        g_flow = torch.zeros((B, 2, H, W), device=dev) if g_flow is None else g_flow.contiguous()
        g_h = g_hh[0]
        pp = L.PredParams()
        pp.B, pp.Cin, pp.Cout, pp.H, pp.W = B, 32, 2, H, W
        pp.x_cl, pp.w, pp.b, pp.y, pp.g_y = L.ptr(bank.zs[N_L - 1][t + 1]), L.ptr(w), L.ptr(b), L.ptr(bank.flow[t]), L.ptr(g_flow)
        pp.g_x, pp.g_w, pp.g_b = L.ptr(g_h), L.ptr(gs[N_L][0]), L.ptr(gs[N_L][1])
        L.call("ef_pred_bwd", pp)
        for i in reversed(range(N_L)):
            cell = cells[i]
            g_ff, g_rec, g_leak, g_thresh = gs[i]
            leak, thresh = cell.leak.detach().reshape(-1), cell.thresh.detach().reshape(-1)
            v_in = bank.v[i][t - 1] if t > 0 else v0[i]
            z_in = bank.zs[i][t] if t > 0 else z0[i]
            v_out = bank.v[i][t]
            x_cl = bank.zs[i - 1][t + 1] if i > 0 else None
            g_x = (g_hh[1] if g_h is g_hh[0] else g_hh[0]) if i > 0 else None
            g_v_in = g_z_in = None
            if t > 0:
                g_v_in = g_vb[par][i]
                if cell.recurrent:
                    g_z_in = g_zb[par][i]
            if tc:
                q = L.LifBwdTcParams()
                q.B, q.H, q.W, q.has_rec, q.hard_reset = B, H, W, int(cell.recurrent), int(cell.hard_reset)
                q.surrogate, q.act_width = L.SURROGATE_CODES[cell.activation], float(cell._act_width_f)
                q.z_in_cl, q.v_in, q.v_out = L.ptr(z_in), L.ptr(v_in), L.ptr(v_out)
                q.g_out, q.g_v_out, q.g_z_out = L.ptr(g_h), L.ptr(g_v[i]), L.ptr(g_z[i])
                q.leak, q.thresh = L.ptr(leak), L.ptr(thresh)
                q.g_v_in, q.g_w_ff, q.g_leak, q.g_thresh = L.ptr(g_v_in), L.ptr(g_ff), L.ptr(g_leak), L.ptr(g_thresh)
                if i == 0:
                    q.Cin, q.x_f32, q.gI_f32 = Cin0, L.ptr(bank.x_in[t]), L.ptr(scratch)
                else:
                    q.x_cl, q.w_bwd = L.ptr(x_cl), L.ptr(splits[LAYERS[i] + ".bwd"])
                    q.gI_hi, q.gI_mid = L.ptr(gI_hi), L.ptr(gI_mid)
                    q.g_x, q.g_z_in, q.g_w_rec = L.ptr(g_x), L.ptr(g_z_in), L.ptr(g_rec)
                L.call("ef_lif_bwd_tc", q)  # wg_partial NULL: CUDA-core weight gradient, added to g_w_* at once
            else:
                q = L.LifConvBwdParams()
                _fill_fwd(q.f, B, Cin0 if i == 0 else 32, H, W, cell, bank.x_in[t] if i == 0 else None, x_cl, v_in, z_in, v_out, leak, thresh)
                q.g_out, q.g_v_out, q.g_z_out = L.ptr(g_h), L.ptr(g_v[i]), L.ptr(g_z[i])
                q.scratch_gI = L.ptr(scratch)
                q.g_x, q.g_v_in, q.g_z_in = L.ptr(g_x), L.ptr(g_v_in), L.ptr(g_z_in)
                q.g_w_ff, q.g_w_rec, q.g_leak, q.g_thresh = L.ptr(g_ff), L.ptr(g_rec), L.ptr(g_leak), L.ptr(g_thresh)
                L.call("ef_lif_conv_bwd", q)
            g_v[i], g_z[i] = g_v_in, g_z_in
            g_h = g_x
    return grads


def _launch_window(model, bank, T, v0, z0, splits, B, H, W, save_all_v, only=None):
    """
    The launches of a whole window of T steps, LAYER-MAJOR: every feed-forward cell runs its T steps in ONE time-fused launch
    (ef_lif_conv_fwd_window: the state of a tile stays in registers over the window), the two recurrent cells run step by step (their
    recurrent convolution needs the neighbours' spikes of the previous step), the prediction head runs once over T*B images.
    Inputs: bank.x_cl[:T] (split network inputs); outputs: bank.v / bank.zs / bank.flow of steps 0..T-1.
    """
    cells = _cells(model)
    for i, name in enumerate(LAYERS):
        if only is not None and name not in only:  # measurement replays (capture_window_fused): a subset of the layers
            continue
        cell = cells[i]
        leak, thresh = cell.leak.detach().reshape(-1), cell.thresh.detach().reshape(-1)
        x_all = bank.x_cl[:T] if i == 0 else bank.zs[i - 1][1:T + 1]
        if not cell.recurrent:
            q = L.LifConvWindowParams()
            q.B, q.T, q.H, q.W, q.hard_reset, q.save_all_v = B, T, H, W, int(cell.hard_reset), int(save_all_v)
            q.x_cl, q.v_in, q.z_in_cl = L.ptr(x_all), L.ptr(v0[i]), L.ptr(z0[i])
            q.leak, q.thresh, q.w_split = L.ptr(leak), L.ptr(thresh), L.ptr(splits[name])
            q.v_out = L.ptr(bank.v[i][:T] if save_all_v else bank.v[i][T - 1])
            q.z_out_cl = L.ptr(bank.zs[i][1:T + 1])
            L.call("ef_lif_conv_fwd_window", q)
        else:
            for t in range(T):
                p = L.LifConvParams()
                _fill_fwd(p, B, 32, H, W, cell, None, x_all[t], v0[i] if t == 0 else bank.v[i][t - 1], z0[i] if t == 0 else bank.zs[i][t],
                          bank.v[i][t], leak, thresh)
                p.z_out_cl, p.w_split = L.ptr(bank.zs[i][t + 1]), L.ptr(splits[name])
                L.call("ef_lif_conv_fwd", p, tag=(32, 32, True))
    if only is not None and "pred" not in only:
        return
    w, b = model.pred.conv2d.weight.detach(), model.pred.conv2d.bias.detach()
    pp = L.PredParams()
    pp.B, pp.Cin, pp.Cout, pp.H, pp.W = T * B, 32, 2, H, W
    pp.x_cl, pp.w, pp.b, pp.y = L.ptr(bank.zs[N_L - 1][1:T + 1]), L.ptr(w), L.ptr(b), L.ptr(bank.flow[:T])
    L.call("ef_pred_fwd", pp)


def capture_window_fused(model, xs0, xs, only=None, save_all_v=False):
    """
    Measurement aid (bench.py roofline): the launches of one time-fused window (xs [T,B,Cin,H,W]) on a private activation bank, captured
    as ONE CUDA graph.  The window starts from the state a first eager window over xs0 leaves (a non-zero state, like every window but
    the first of a sequence).  `only`: subset of LAYERS + ("pred",) to capture (the eager passes always run everything, so the inputs
    of the selected layers are in place).  Returns (graph, kernel launches per replay, cell-steps per replay).
    """
    T, B, Cin0, H, W = xs.shape
    _cells(model)
    splits = _split_cache(model)
    banks = [_Bank((B, H, W, xs.device), T), _Bank((B, H, W, xs.device), T)]
    none = [None] * N_L
    with torch.no_grad():
        for bank, x in zip(banks, (xs0, xs)):
            bank.need_input(Cin0, False, True)
            ops.pack_split_cl(x.reshape(T * B, Cin0, H, W).contiguous(), out=bank.x_cl[:T].view(T * B, H, W, 32))
        _launch_window(model, banks[0], T, none, none, splits, B, H, W, True)
        v0 = [banks[0].v[i][T - 1] for i in range(N_L)]
        z0 = [banks[0].zs[i][T] for i in range(N_L)]
        _launch_window(model, banks[1], T, v0, z0, splits, B, H, W, save_all_v)
        torch.cuda.synchronize()
        g = _capture(lambda: _launch_window(model, banks[1], T, v0, z0, splits, B, H, W, save_all_v, only=only))
    g._keepalive = (banks, splits)
    names = LAYERS + ("pred",) if only is None else only
    cells = dict(zip(LAYERS, _cells(model)))
    launches = sum(1 if n == "pred" or not cells[n].recurrent else T for n in names)
    return g, launches, T * sum(1 for n in names if n != "pred")


WINDOW_LAUNCHES = lambda T: 5 + 2 * T + 1  # noqa: E731  kernels of one window forward: 5 fused cells + 2 recurrent cells x T + prediction


class _FireNetWindow(torch.autograd.Function):
    """T model steps at once (model.forward_window).  One autograd node per WINDOW; its backward is the deferred window backward."""

    @staticmethod
    def forward(ctx, model, xs, *params):
        fs = model._fast
        T, B, Cin0, H, W = xs.shape
        dev = xs.device
        _cells(model)
        splits = _split_cache(model)
        if "head" not in splits:
            raise L.EventFlowError(f"forward_window: the head layer needs at most {L.EF_HEAD_MAX_CIN} input channels (got {Cin0})")
        if fs.step != 0:
            raise RuntimeError("forward_window must start at a window boundary (after reset_states() / detach_states())")
        arena = model.__dict__.get("_arena")
        if arena is None or arena.key != (B, H, W, dev):
            arena = model.__dict__["_arena"] = _Arena(B, H, W, dev, max(T, int(model.__dict__.get("_window_cap", DEFAULT_WINDOW_CAP))))
        parity = arena.parity
        bank = arena.bank(parity)
        while T > bank.cap:
            bank.grow(0)
        bank.need_input(Cin0, False, True)
        need_grad = not isinstance(ctx, _NoCtx)
        v0, z0 = list(fs.v), list(fs.z)
        guards = [] if fs.src is None else [fs.src]
        for t in range(T):
            slot = bank.slot(t)
            slot.gen += 1
            guards.append((slot, slot.gen))
        ops.pack_split_cl(xs.reshape(T * B, Cin0, H, W), out=bank.x_cl[:T].view(T * B, H, W, 32))
        use_graph = model.__dict__.get("_use_graphs", True) and L.PROFILE is None and not torch.cuda.is_current_stream_capturing()
        if use_graph:
            if fs.param_sig is None:
                fs.param_sig = (tuple(p.data_ptr() for p in _params_of(model)), tuple(splits[n].data_ptr() for n in LAYERS if n in splits))
            key = ("window", fs.param_sig, tuple(0 if v is None else v.data_ptr() for v in v0), T, need_grad, bank.x_cl.data_ptr())
            graphs = bank.__dict__.setdefault("window_graphs", {})
            g = graphs.get(key)
            if g is None:
                _launch_window(model, bank, T, v0, z0, splits, B, H, W, need_grad)
                g = graphs[key] = _capture(lambda: _launch_window(model, bank, T, v0, z0, splits, B, H, W, need_grad))
            else:
                g.replay()
                L.GRAPH_KERNELS += WINDOW_LAUNCHES(T)
        else:
            _launch_window(model, bank, T, v0, z0, splits, B, H, W, need_grad)
        last = bank.slot(T - 1)
        for i in range(N_L):
            fs.v[i], fs.z[i] = last.v[i], last.z[i]
        fs.src = (last, last.gen)
        fs.step = T
        carry = fs.carry
        carry.prev, carry.parity, carry.n, carry.head_tc = (v0, z0), parity, T, True
        flows = bank.flow[:T].clone()  # the caller may keep the flows for as long as it likes; the bank is recycled
        ctx.model, ctx.arena, ctx.carry, ctx.guards, ctx.shapes = model, arena, carry, guards, (B, Cin0, H, W)
        model._last_spikes = last.z
        return tuple(flows.unbind(0))

    @staticmethod
    def backward(ctx, *g_flows):
        model, carry = ctx.model, ctx.carry
        for slot_, gen_ in ctx.guards:
            if slot_.gen != gen_:
                raise RuntimeError(
                    "event_flow_b200 fast path: the activations this backward needs were overwritten by a later forward pass. "
                    "loss.backward() of a BPTT window must run before the next window completes (see fast._Arena).")
        carry.g_flows = {t: g for t, g in enumerate(g_flows) if g is not None}
        params = _params_of(model)
        grads = _window_backward(model, ctx.arena, carry, ctx.shapes)
        carry.g_flows = {}
        if carry.into_sink:  # already accumulated into p.grad (views of the trainer's flat buffer)
            return (None, None, *[None] * len(params))
        return (None, None, *[g.clone() if p.requires_grad else None for p, g in zip(params, grads)])


def forward_window(model, xs):
    """
    T forward passes of a LIF FireNet at once, layer-major and time-fused (model.forward_window).  xs [T,B,Cin,H,W].
    Returns the list of the T flow maps [B,2,H,W]; numerically the same steps as T calls of forward().
    """
    fs = model._fast
    if fs is None:
        fs = FastState(len(LAYERS))
        for i, s in enumerate(model._states):
            if s is not None:
                fs.v[i] = s[0].detach().contiguous()
                fs.z[i] = ops.pack_cl(s[1].detach())
        model._fast = fs
    params = _params_of(model)
    xs = xs.contiguous()
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return list(_FireNetWindow.apply(model, xs, *params))
    with torch.no_grad():
        flows = list(_FireNetWindow.forward(_NoCtx(), model, xs, *params))
    fs.detach(model.__dict__.get("_arena"))
    return flows


def forward(model, x, log=False):
    """One forward pass of a LIF FireNet on the fast path.  Returns the reference's output dict."""
    fs = model._fast
    if fs is None:
        fs = FastState(len(LAYERS))
        for i, s in enumerate(model._states):  # states set through the reference-format API are converted once
            if s is not None:
                fs.v[i] = s[0].detach().contiguous()
                fs.z[i] = ops.pack_cl(s[1].detach())
        model._fast = fs
    params = _params_of(model)
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    if need_grad:
        token = fs.token
        # the parameters are autograd inputs of the window's FIRST step only (its backward runs the window's whole BPTT and hands
        # over the gradients); later steps are chained through the token, which keeps their apply() cheap
        if token is None:
            flow, fs.token = _FireNetStep.apply(model, x, None, *params)
        else:
            flow, fs.token = _FireNetStep.apply(model, x, token)
    else:
        with torch.no_grad():
            flow, _ = _FireNetStep.forward(_NoCtx(), model, x, None, *params)
        fs.detach(model.__dict__.get("_arena"))
    activity = None
    if log:
        names = ["0:input", "1:head", "2:G1", "3:R1a", "4:R1b", "5:G2", "6:R2a", "7:R2b", "8:pred"]
        vals = [x] + list(model._last_spikes) + [flow]
        activity = {n: t.detach().ne(0).float().mean().item() for n, t in zip(names, vals)}
    return {"flow": [flow], "activity": activity}


class _NoCtx:
    """Stand-in for the autograd context when the step runs without gradient tracking."""


def detach(model):
    """End of a BPTT window (model.detach_states()): cut the chain, switch the activation bank."""
    model._fast.detach(model.__dict__.get("_arena"))


def states_of(model):
    """Reference-format view of the internal state: list of stacked [2,B,C,H,W] tensors (fresh tensors = clones)."""
    fs = model._fast
    out = []
    for v, z in zip(fs.v, fs.z):
        out.append(None if v is None else torch.stack([v, ops.unpack_cl(z)]))
    return out
