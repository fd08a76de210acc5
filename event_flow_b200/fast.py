"""
Fast path of the spiking FireNet chain (models/model.py:254-265) on the internal formats.

Between the cells, spikes never exist as fp32 NCHW tensors: every cell writes bf16 channels-last spikes ("cl") that the
next cell's tcgen05 kernel consumes through TMA; membrane potentials stay fp32 NCHW (the reference's state format).  One
torch.autograd node per MODEL step (instead of ~100 per step in the reference) carries the BPTT: the per-layer state
gradients travel from step t+1 to step t in a side structure (`_Carry`), the autograd graph only orders the steps through
a scalar token and routes the flow / parameter gradients.
"""
import torch

from . import _lib as L
from . import ops

LAYERS = ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b")


N_L = len(LAYERS)


class _Carry:
    """
    BPTT side channel of one window: dL/d(v, z) of every layer's state, handed from the backward of step t+1 to the
    backward of step t, and the flat parameter-gradient buffer the kernels accumulate into over the whole window.
    """

    def __init__(self):
        self.g_v = [None] * N_L
        self.g_z = [None] * N_L
        self.flat = None      # fp32 [n_params]: += by every step's backward kernels, handed to autograd by the window's first step
        self.sweep = 0        # number of backward steps executed in the current sweep (0 = next call starts a new sweep)


class _Slot:
    """Activations of one model step: membrane fp32 [B,32,H,W] and spikes bf16 [B,H,W,32] of the 7 layers, one allocation."""

    def __init__(self, B, H, W, dev):
        nv, nz = B * 32 * H * W * 4, B * 32 * H * W * 2
        self.slab = torch.empty(N_L * (nv + nz), device=dev, dtype=torch.uint8)
        self.v, self.z = [], []
        o = 0
        for _ in range(N_L):
            self.v.append(self.slab[o:o + nv].view(torch.float32).view(B, 32, H, W))
            o += nv
        for _ in range(N_L):
            self.z.append(self.slab[o:o + nz].view(torch.bfloat16).view(B, H, W, 32))
            o += nz
        self.flow = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
        self.gen = 0       # bumped every time a forward step (re)writes this slot: saved activations of older steps are then gone
        self.x_in = None   # static copy of the model input (graph replay reads a fixed address)
        self.graphs = {}   # (pointer signature) -> torch.cuda.CUDAGraph of this step's 8 kernels
        self.bwd_calls = {}  # (pointer signature, sweep position) -> prepared argument structs (+ CUDA graph) of this step's backward
        self.g_flow_in = None  # static copy of the incoming flow gradient (graph replay reads a fixed address)


class _Arena:
    """
    Activation storage owned by the model and reused window after window (no allocator traffic in steady state, static
    addresses for CUDA-graph replay).  Two banks alternate per BPTT window: window w writes bank w%2, its initial state
    lives in the last slot of bank (w-1)%2.  CONTRACT: loss.backward() of window w must run before window w+1 COMPLETES
    (window w+1 overwrites, at its own last step, the slot that holds window w's initial state; window w+2 overwrites
    window w's activations) -- train_flow.py:154-171 runs backward right at the window end.  The contract is enforced:
    every slot carries a generation counter, a step's backward raises if a slot it saved from has been rewritten since.
    """

    def __init__(self, B, H, W, dev):
        self.key = (B, H, W, dev)
        self.banks = ([], [])
        self.parity = 0
        self.bwd = None
        self.flat = None   # fp32 [n_params]: parameter gradients of the window being back-propagated (static address)

    def slot(self, parity, idx):
        bank = self.banks[parity]
        while len(bank) <= idx:
            bank.append(_Slot(*self.key))
        return bank[idx]

    def bwd_buffers(self):
        if self.bwd is None:
            B, H, W, dev = self.key
            mk = lambda: torch.empty((B, 32, H, W), device=dev, dtype=torch.float32)  # noqa: E731
            mkcl = lambda: torch.empty((B, H, W, 32), device=dev, dtype=torch.bfloat16)  # noqa: E731
            self.bwd = {"g_h": [mk(), mk()], "scratch": mk(), "g_v": [[mk() for _ in range(N_L)] for _ in range(2)],
                        "g_z": [[mk() for _ in range(N_L)] for _ in range(2)], "gI_hi": mkcl(), "gI_mid": mkcl(),
                        # per-CTA partial sums of the tensor-core weight gradient, one buffer per hidden cell (kept over a sweep)
                        "wg": [None] + [torch.empty(L.lib().ef_lif_wgrad_partial_elems(B, H, W, 1), device=dev, dtype=torch.float32)
                                        for _ in range(N_L - 1)]}
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
    sig = tuple(c.ff.weight._version for c in cells[1:]) + tuple(c.rec.weight._version for c in cells[1:] if c.recurrent) + (
        model.__dict__.get("_w_epoch", 0), cells[1].ff.weight.data_ptr())
    last = cache.get("__last__")
    if last is not None and last[0] == sig:
        return last[1]
    out = {}
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
    """The 8 kernels of one model step: head, 6 tensor-core cells, prediction head.  All tensors are caller-provided."""
    h = None
    cells = _cells(model)
    for i, name in enumerate(LAYERS):
        if only_hidden and i == 0:  # measurement replays (bench.py): the head's spikes are already in the slot
            h = slot.z[0]
            continue
        cell = cells[i]
        leak, thresh = cell.leak.detach().reshape(-1), cell.thresh.detach().reshape(-1)
        p = L.LifConvParams()
        _fill_fwd(p, B, Cin0 if i == 0 else 32, H, W, cell, x if i == 0 else None, h, v_in[i], z_in[i], slot.v[i], leak, thresh)
        p.z_out_cl = L.ptr(slot.z[i])
        if i > 0:
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
    Measurement aid (bench.py roofline): the kernels of len(xs) - 1 consecutive model steps on private activation slots,
    captured as ONE CUDA graph (step 0 runs eagerly from the zero state and provides the previous state of step 1).
    With only_hidden the graph holds just the six 32 -> 32 tensor-core cell launches of every step.  Replaying it repeats
    the same computation on the same operands (every launch streams its own ~59 MB, the whole replay far more than L2).
    Returns (graph, kernel launches per replay).
    """
    x0 = xs[0]
    B, Cin0, H, W = x0.shape
    for name in LAYERS:
        cell = getattr(model, name)
        if not hasattr(cell, "_act_width_f"):
            cell._act_width_f = float(cell.act_width)
    splits = _split_cache(model)
    slots = [_Slot(B, H, W, x0.device) for _ in xs]
    none = [None] * N_L
    xs = [x.contiguous() for x in xs]
    with torch.no_grad():
        _launch_step(model, xs[0], none, none, slots[0], splits, B, Cin0, H, W)
        for t in range(1, len(xs)):  # eager pass: fills every slot (the hidden-only replay needs the head spikes in place)
            _launch_step(model, xs[t], slots[t - 1].v, slots[t - 1].z, slots[t], splits, B, Cin0, H, W)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for t in range(1, len(xs)):
                _launch_step(model, xs[t], slots[t - 1].v, slots[t - 1].z, slots[t], splits, B, Cin0, H, W, only_hidden=only_hidden)
    g._keepalive = (slots, xs, splits)
    return g, (len(xs) - 1) * (N_L - 1 if only_hidden else N_L + 1)


class _FireNetStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, token, *params):
        fs = model._fast
        x = x.contiguous()
        B, Cin0, H, W = x.shape
        dev = x.device
        for cell in _cells(model):
            if "_act_width_f" not in cell.__dict__:
                cell._act_width_f = float(cell.act_width)
        splits = _split_cache(model)
        arena = model.__dict__.get("_arena")
        if arena is None or arena.key != (B, H, W, dev):
            arena = model.__dict__["_arena"] = _Arena(B, H, W, dev)
        slot = arena.slot(arena.parity, fs.step)
        fs.step += 1
        slot.gen += 1
        v_in, z_in = list(fs.v), list(fs.z)
        ctx.guards = [(slot, slot.gen)] + ([fs.src] if fs.src is not None and fs.src[0] is not slot else [])
        fs.src = (slot, slot.gen)
        cap = model.__dict__.get("_capture")
        use_graph = (model.__dict__.get("_use_graphs", True) and cap is None and L.PROFILE is None
                     and not torch.cuda.is_current_stream_capturing())
        if use_graph:
            # CUDA-graph replay of the step: the input is copied to a fixed address, everything else already is static
            if slot.x_in is None or slot.x_in.shape != x.shape:
                slot.x_in, slot.graphs = torch.empty_like(x), {}
            slot.x_in.copy_(x)
            if fs.param_sig is None:  # refreshed per sequence (reset_states) and whenever the module is moved (FireNet._apply)
                fs.param_sig = (tuple(p.data_ptr() for p in _params_of(model)), tuple(splits[n].data_ptr() for n in LAYERS[1:]))
            key = (fs.param_sig, tuple(0 if v is None else v.data_ptr() for v in v_in))
            ctx.splits = splits
            ctx.fkey = key
            g = slot.graphs.get(key)
            if g is None:
                _launch_step(model, slot.x_in, v_in, z_in, slot, splits, B, Cin0, H, W)  # eager: results + lazy init
                g = torch.cuda.CUDAGraph()
                # thread_local: other threads (the NCCL watchdog under data parallelism) keep issuing CUDA calls during a capture
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    _launch_step(model, slot.x_in, v_in, z_in, slot, splits, B, Cin0, H, W)
                slot.graphs[key] = g
            else:
                g.replay()
                L.GRAPH_KERNELS += N_L + 1
            x_used = slot.x_in
        else:
            _launch_step(model, x, v_in, z_in, slot, splits, B, Cin0, H, W)
            x_used = x
            ctx.splits = splits
            ctx.fkey = None
        saved = []
        for i, name in enumerate(LAYERS):
            saved.append((x_used if i == 0 else None, slot.z[i - 1] if i > 0 else None, v_in[i], z_in[i], slot.v[i]))
            if cap is not None:  # test hook: what this layer consumed and produced, in the reference's tensor format
                xin = x if i == 0 else ops.unpack_cl(slot.z[i - 1])
                sin = None if v_in[i] is None else torch.stack([v_in[i], ops.unpack_cl(z_in[i])]).cpu()
                zo = ops.unpack_cl(slot.z[i])
                cap[name] = (xin.detach().cpu(), sin, zo.cpu(), torch.stack([slot.v[i], zo]).cpu())
            fs.v[i], fs.z[i] = slot.v[i], slot.z[i]
        flow = slot.flow.clone()  # the caller may keep the flow for as long as it likes; the slot is recycled
        ctx.model, ctx.saved, ctx.flow, ctx.first, ctx.z_last = model, saved, slot.flow, token is None, slot.z[N_L - 1]
        ctx.carry, ctx.arena, ctx.slot = fs.carry, arena, slot
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
                    "loss.backward() of a BPTT window must run before the next window completes (see fast._Arena); "
                    "use model._use_graphs / per-cell API for other schedules.")
        B, Cin0, H, W = ctx.shapes
        params = _params_of(model)
        dev = ctx.flow.device
        buf = ctx.arena.bwd_buffers()
        arena = ctx.arena
        sweep_first = carry.sweep == 0
        if sweep_first:  # first backward call of this sweep = last step of the window
            n = sum(p.numel() for p in params)
            if arena.flat is None or arena.flat.numel() != n:
                arena.flat = torch.zeros(n, device=dev, dtype=torch.float32)
            else:
                arena.flat.zero_()
            carry.flat = arena.flat
            carry.g_v, carry.g_z = [None] * N_L, [None] * N_L
        par = carry.sweep & 1  # ping-pong of the state-gradient buffers between consecutive steps
        wg_flags = (L.EF_WG_ACCUMULATE if carry.sweep > 0 else 0) | (L.EF_WG_FINALIZE if ctx.first else 0)
        carry.sweep += 1
        if g_flow is None:
            g_flow = torch.zeros_like(ctx.flow)
        g_flow = g_flow.contiguous()
        # Steady state: the argument structs of this step's ~8 library calls depend only on static addresses (arena slots,
        # ping-pong buffers, the flat gradient buffer) -- they are built once per (slot, sweep position) and replayed; only the
        # incoming flow gradient lives at a fresh address.
        ckey = None if ctx.fkey is None else (ctx.fkey, par, sweep_first, ctx.first, model.__dict__.get("_tc_backward", True),
                                              model.__dict__.get("_tc_wgrad", True))
        hit = ctx.slot.bwd_calls.get(ckey) if ckey is not None else None
        if hit is not None:
            calls, g_v_next, g_z_next, graph, n_kernels = hit
            slot = ctx.slot
            if slot.g_flow_in is None:
                slot.g_flow_in = torch.empty_like(ctx.flow)
            slot.g_flow_in.copy_(g_flow)  # fixed address for the replay, like the model input of the forward graph
            calls[0][1].g_y = L.ptr(slot.g_flow_in)
            if graph is None and model.__dict__.get("_use_graphs", True) and L.PROFILE is None and not torch.cuda.is_current_stream_capturing():
                graph = torch.cuda.CUDAGraph()  # second visit of this (slot, sweep position): capture the ~25 kernels once
                n0 = L.lib().ef_launch_count()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    for name, st in calls:
                        L.call(name, st)
                n_kernels = L.lib().ef_launch_count() - n0  # kernels recorded into the graph (counted by the library itself)
                slot.bwd_calls[ckey] = (calls, g_v_next, g_z_next, graph, n_kernels)
            if graph is not None:
                graph.replay()
                L.GRAPH_KERNELS += n_kernels
            else:
                for name, st in calls:
                    L.call(name, st)
            carry.g_v, carry.g_z = list(g_v_next), list(g_z_next)
            if not ctx.first:
                return (None, None, torch.zeros((), device=dev, dtype=torch.float32))
            carry.sweep = 0
            grads, o = [], 0
            for p in params:
                grads.append(carry.flat[o:o + p.numel()].view(p.shape))
                o += p.numel()
            return (None, None, None, *[g.clone() if p.requires_grad else None for p, g in zip(params, grads)])
        calls = []

        def emit(name, st):
            calls.append((name, st))
            L.call(name, st)

        grads, o = [], 0
        for p in params:
            grads.append(carry.flat[o:o + p.numel()].view(p.shape))
            o += p.numel()
        gi = len(grads) - 2
        # prediction head
        g_h = buf["g_h"][0]
        pp = L.PredParams()
        pp.B, pp.Cin, pp.Cout, pp.H, pp.W = B, 32, 2, H, W
        z7 = ctx.z_last
        w, b = model.pred.conv2d.weight.detach(), model.pred.conv2d.bias.detach()
        pp.x_cl, pp.w, pp.b, pp.y, pp.g_y = L.ptr(z7), L.ptr(w), L.ptr(b), L.ptr(ctx.flow), L.ptr(g_flow)
        pp.g_x, pp.g_w, pp.g_b = L.ptr(g_h), L.ptr(grads[gi]), L.ptr(grads[gi + 1])
        emit("ef_pred_bwd", pp)
        # cells, last to first
        for i in reversed(range(len(LAYERS))):
            cell = getattr(model, LAYERS[i])
            x_f32, x_cl, v_in, z_in, v_out = ctx.saved[i]
            gi -= 4 if cell.recurrent else 3
            leak, thresh = cell.leak.detach().reshape(-1), cell.thresh.detach().reshape(-1)
            g_x = (buf["g_h"][1] if g_h is buf["g_h"][0] else buf["g_h"][0]) if i > 0 else None
            g_v_in = g_z_in = None
            if not ctx.first and v_in is not None:
                g_v_in = buf["g_v"][par][i]
                if cell.recurrent:
                    g_z_in = buf["g_z"][par][i]
            k = gi
            g_w_ff = grads[k]
            k += 1
            g_w_rec = None
            if cell.recurrent:
                g_w_rec = grads[k]
                k += 1
            g_leak, g_thresh = grads[k], grads[k + 1]
            if i == 0 and Cin0 <= 8 and model.__dict__.get("_tc_backward", True):
                # head layer on the fast formats: pointwise (fp32 g_I) + register-tiled CUDA-core weight gradient, no data gradient
                t = L.LifBwdTcParams()
                t.B, t.H, t.W, t.has_rec, t.hard_reset = B, H, W, 0, int(cell.hard_reset)
                t.surrogate, t.act_width = L.SURROGATE_CODES[cell.activation], float(cell._act_width_f)
                t.z_in_cl, t.v_in, t.v_out = L.ptr(z_in), L.ptr(v_in), L.ptr(v_out)
                t.g_out, t.g_v_out, t.g_z_out = L.ptr(g_h), L.ptr(carry.g_v[i]), L.ptr(carry.g_z[i])
                t.leak, t.thresh = L.ptr(leak), L.ptr(thresh)
                t.g_v_in, t.g_w_ff, t.g_leak, t.g_thresh = L.ptr(g_v_in), L.ptr(g_w_ff), L.ptr(g_leak), L.ptr(g_thresh)
                t.Cin, t.x_f32, t.gI_f32 = Cin0, L.ptr(x_f32), L.ptr(buf["scratch"])
                emit("ef_lif_bwd_tc", t)
            elif i > 0 and model.__dict__.get("_tc_backward", True):
                # 32 -> 32 cells: pointwise + tensor-core data gradient + channels-last weight gradient
                t = L.LifBwdTcParams()
                t.B, t.H, t.W, t.has_rec, t.hard_reset = B, H, W, int(cell.recurrent), int(cell.hard_reset)
                t.surrogate, t.act_width = L.SURROGATE_CODES[cell.activation], float(cell._act_width_f)
                t.x_cl, t.z_in_cl, t.v_in, t.v_out = L.ptr(x_cl), L.ptr(z_in), L.ptr(v_in), L.ptr(v_out)
                t.g_out, t.g_v_out, t.g_z_out = L.ptr(g_h), L.ptr(carry.g_v[i]), L.ptr(carry.g_z[i])
                t.leak, t.thresh, t.w_bwd = L.ptr(leak), L.ptr(thresh), L.ptr(ctx.splits[LAYERS[i] + ".bwd"])
                t.gI_hi, t.gI_mid = L.ptr(buf["gI_hi"]), L.ptr(buf["gI_mid"])
                t.g_x, t.g_v_in, t.g_z_in = L.ptr(g_x), L.ptr(g_v_in), L.ptr(g_z_in)
                t.g_w_ff, t.g_w_rec, t.g_leak, t.g_thresh = L.ptr(g_w_ff), L.ptr(g_w_rec), L.ptr(g_leak), L.ptr(g_thresh)
                if model.__dict__.get("_tc_wgrad", True):
                    t.wg_partial, t.wg_flags = L.ptr(buf["wg"][i]), wg_flags
                emit("ef_lif_bwd_tc", t)
            else:
                q = L.LifConvBwdParams()
                _fill_fwd(q.f, B, Cin0 if i == 0 else 32, H, W, cell, x_f32, x_cl, v_in, z_in, v_out, leak, thresh)
                q.g_out, q.g_v_out, q.g_z_out = L.ptr(g_h), L.ptr(carry.g_v[i]), L.ptr(carry.g_z[i])
                q.scratch_gI = L.ptr(buf["scratch"])
                q.g_x, q.g_v_in, q.g_z_in = L.ptr(g_x), L.ptr(g_v_in), L.ptr(g_z_in)
                q.g_w_ff, q.g_w_rec, q.g_leak, q.g_thresh = L.ptr(g_w_ff), L.ptr(g_w_rec), L.ptr(g_leak), L.ptr(g_thresh)
                emit("ef_lif_conv_bwd", q)
            carry.g_v[i], carry.g_z[i] = g_v_in, g_z_in
            g_h = g_x
        if ckey is not None:
            ctx.slot.bwd_calls[ckey] = (calls, list(carry.g_v), list(carry.g_z), None, 0)
        if not ctx.first:  # parameter gradients keep accumulating in carry.flat; the window's first step hands them over
            return (None, None, torch.zeros((), device=dev, dtype=torch.float32))
        carry.sweep = 0
        out = [g.clone() if p.requires_grad else None for p, g in zip(params, grads)]
        return (None, None, None, *out)


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
        # the parameters are autograd inputs of the window's FIRST step only (its backward hands over the gradients accumulated
        # over the whole window); later steps are chained through the token, which keeps their apply() cheap
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
