"""
CUDA-graph replay of a model's per-step forward pass under torch.no_grad() -- the evaluation loop of eval_flow.py:103-160 calls
`model(voxel, cnt)` once per time step with tensors of one shape, and for the cell-by-cell models (ANN FireNet at batch 1, the PLIF /
ALIF / XLIF FireNets, the ANN U-Nets) such a step is a few dozen small launches whose cost is the host's, not the GPU's.  The model's
own forward code is what gets captured; nothing is re-implemented here.

How a step becomes a graph.  A *situation* is (input shape, nested shapes of the recurrent states, identity + version of every
parameter, the weight-image epoch of ops.py, train/eval flag).  The first step in a situation runs eagerly, the second is captured,
later ones replay: the input is copied into the graph's static input, the recurrent states live in static tensors the graph updates in
place at its end (new state -> static state), outputs are returned as clones.  States the caller (or an eager step, or another graph)
put into the model are copied into the static tensors first; before any eager step -- log=True, grad mode -- static tensors are
replaced by clones, so nothing autograd saves is ever overwritten by a replay.  A parameter update or a new shape is a new situation;
a model with forward hooks is never replayed (the hooks would not run).
"""
import torch

from . import fast, ops

MAX_GRAPHS = 4  # per model: (first step of a sequence, later steps) x a couple of input shapes


def _flat(states, out):
    """Leaves of a nested state structure (list / tuple / tensor / None) in a fixed order."""
    if torch.is_tensor(states):
        out.append(states)
    elif isinstance(states, (list, tuple)):
        for s in states:
            _flat(s, out)
    return out


def _shape_of(states):
    if torch.is_tensor(states):
        return (tuple(states.shape), states.dtype, states.stride())
    if isinstance(states, (list, tuple)):
        return (type(states).__name__,) + tuple(_shape_of(s) for s in states)
    return None if states is None else type(states).__name__


def _rebuild(like, leaves):
    """The structure of `like` with its tensor leaves taken from the iterator `leaves`."""
    if torch.is_tensor(like):
        return next(leaves)
    if isinstance(like, list):
        return [_rebuild(s, leaves) for s in like]
    if isinstance(like, tuple):
        return tuple(_rebuild(s, leaves) for s in like)
    return like


def _situation(model, x, states):
    """None if the step must not be replayed at all (forward hooks: they would not fire)."""
    known = model.__dict__.get("_graph_params")
    if known is None:
        known = model.__dict__["_graph_params"] = (list(model.parameters()), list(model.modules()))
    params, modules = known
    for m in modules:
        if m._forward_hooks or m._forward_pre_hooks:
            return None
    return (tuple(x.shape), x.dtype, x.device, x.stride(), _shape_of(states), model.training, ops.WEIGHT_EPOCH,
            tuple((p.data_ptr(), p._version) for p in params))


class _StepGraph:
    def __init__(self, x, holder, attr, eager):
        self.holder, self.attr = holder, attr
        self.x = x.clone()
        before = getattr(holder, attr)
        self.static_in = [t.clone() for t in _flat(before, [])]
        setattr(holder, attr, _rebuild(before, iter(self.static_in)))
        box = {}

        def body():
            # (runs ONCE, under capture; written so that running it again -- the CPU stand-in of tests/test_graphed_cpu.py replays by
            # re-running it -- reuses the same static tensors, as a replay of the real graph does)
            out = eager(self.x)
            after = getattr(holder, attr)
            new = _flat(after, [])
            if "keep" not in box:
                # a state that existed before the step is updated in place; one the step created (first step of a sequence) gets its
                # own static tensor.  (Same leaf order before and after: cells that have a state keep it, the others turn None into one.)
                alias = _shape_of(before) == _shape_of(after)
                box["keep"] = list(self.static_in) if alias else [torch.empty_like(t) for t in new]
                box["flows"], box["after"] = list(out["flow"]), after
            else:
                for dst, f in zip(box["flows"], out["flow"]):
                    dst.copy_(f)
            for dst, t in zip(box["keep"], new):
                dst.copy_(t)

        try:
            self.graph = fast._capture(body)
        finally:
            setattr(holder, attr, before)
        self.flows = box["flows"]
        self.static_out = box["keep"]
        self.after = box["after"]

    def replay(self, x):
        self.x.copy_(x)
        for dst, src in zip(self.static_in, _flat(getattr(self.holder, self.attr), [])):
            if src is not dst:
                dst.copy_(src)
        self.graph.replay()
        setattr(self.holder, self.attr, _rebuild(self.after, iter(self.static_out)))
        return {"flow": [f.clone() for f in self.flows], "activity": None}

    def owns(self, t):
        return any(t is s for s in self.static_out) or any(t is s for s in self.static_in)


def usable(model, x, log=False):
    return (not log and torch.is_tensor(x) and x.is_cuda and not torch.is_grad_enabled() and not model.__dict__.get("_graph_off", False)
            and not torch.cuda.is_current_stream_capturing())


def leave(model, holder, attr):
    """Before an eager step: states that are static tensors of a graph are replaced by clones (a later replay overwrites the originals)."""
    graphs = model.__dict__.get("_step_graphs")
    if not graphs:
        return
    states = getattr(holder, attr)
    leaves = _flat(states, [])
    owned = [any(isinstance(g, _StepGraph) and g.owns(t) for g in graphs.values()) for t in leaves]
    if any(owned):
        setattr(holder, attr, _rebuild(states, iter([t.clone() if o else t for t, o in zip(leaves, owned)])))


def step(model, x, holder, attr, eager):
    """
    One no-grad forward step of `model` on `x`; `getattr(holder, attr)` is the model's list of recurrent states and `eager(x)` its own
    forward code (returns the output dict, leaves the new states in the holder).
    """
    graphs = model.__dict__.setdefault("_step_graphs", {})
    key = _situation(model, x, getattr(holder, attr))
    if key is None:
        leave(model, holder, attr)
        return eager(x)
    g = graphs.get(key)
    if g is None:  # first time in this situation: eager, remember that it happened
        while len(graphs) >= MAX_GRAPHS:
            graphs.pop(next(iter(graphs)))
        graphs[key] = "seen"
        leave(model, holder, attr)
        return eager(x)
    if g == "seen":
        try:
            g = graphs[key] = _StepGraph(x, holder, attr, eager)
        except Exception as exc:  # something in this model's step cannot be captured (a host read, ...): stay on the launch-by-launch path
            model.__dict__["_graph_off"], model.__dict__["_graph_error"] = True, repr(exc)
            graphs.clear()
            return eager(x)
    return g.replay(x)

