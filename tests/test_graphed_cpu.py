"""
Host logic of event_flow_b200/graphed.py without a GPU: the CUDA-graph capture is replaced by a stand-in that re-runs the captured
Python body on replay() (test infrastructure only), so situations, static state tensors, caller-set states, eager steps in between,
weight updates, hooks and the capture-failure fallback are exercised against a twin that steps eagerly.
"""
import copy

import pytest
import torch

from event_flow_b200 import fast, graphed


class Toy(torch.nn.Module):
    """Two recurrent states (one of them a tuple, like ConvLSTM's) and a list of two outputs."""

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor([0.5, -0.25, 0.125]))
        self.lin = torch.nn.Linear(3, 3)
        self._states = [None, None]
        self.calls = 0

    def eager(self, x):
        self.calls += 1
        h = self._states[0]
        h = torch.tanh(self.lin(x) + (0 if h is None else h * self.a))
        c = self._states[1]
        c = (h, h * 2) if c is None else (c[0] * 0.5 + h, c[1] - h)
        self._states = [h, c]  # (a new list, as some of the models do; others assign entries)
        return {"flow": [h + c[0], c[1]], "activity": None}

    def forward(self, x):
        if not torch.is_grad_enabled() and not self.__dict__.get("_graph_off"):
            return graphed.step(self, x, self, "_states", self.eager)
        graphed.leave(self, self, "_states")
        return self.eager(x)


class FakeGraph:
    def __init__(self, fn):
        self.fn = fn

    def replay(self):
        self.fn()


@pytest.fixture
def fake_capture(monkeypatch):
    made = []

    def capture(fn):
        fn()  # (a real capture records without executing; executing here only disturbs tensors the replay protocol rewrites anyway)
        made.append(1)
        return FakeGraph(fn)

    monkeypatch.setattr(fast, "_capture", capture)
    return made


def pair():
    torch.manual_seed(0)
    a = Toy()
    b = copy.deepcopy(a)
    b.__dict__["_graph_off"] = True
    return a, b


def same(x, y):
    fx, fy = graphed._flat(x, []), graphed._flat(y, [])
    assert len(fx) == len(fy) and all(torch.equal(p, q) for p, q in zip(fx, fy))


def test_replayed_steps_equal_eager_steps(fake_capture):
    a, b = pair()
    outs = []
    with torch.no_grad():
        for k in range(14):
            if k == 6:
                a._states, b._states = [None, None], [None, None]  # a new sequence
            if k == 10:
                a._states = copy.deepcopy(b._states)  # states handed in by the caller
            x = torch.randn(2, 3, generator=torch.Generator().manual_seed(k))
            oa, ob = a(x)["flow"], b(x)["flow"]
            same(oa, ob)
            same(a._states, b._states)
            outs.append((oa, [t.clone() for t in oa]))
    assert len(fake_capture) == 2  # first-step situation and later-step situation, each captured once
    for live, kept in outs:  # handed-out outputs are clones: later replays never touch them
        same(live, kept)


def test_eager_steps_in_between_never_see_static_tensors(fake_capture):
    a, b = pair()
    with torch.no_grad():
        for k in range(5):
            x = torch.randn(2, 3, generator=torch.Generator().manual_seed(k))
            same(a(x)["flow"], b(x)["flow"])
    g = [v for v in a.__dict__["_step_graphs"].values() if isinstance(v, graphed._StepGraph)][-1]
    assert all(g.owns(t) for t in graphed._flat(a._states, []))
    x = torch.randn(2, 3, generator=torch.Generator().manual_seed(99))
    la, lb = a(x)["flow"][0].sum(), b(x)["flow"][0].sum()  # grad mode: eager, on clones of the static states
    assert not any(g.owns(t) for t in graphed._flat(a._states, []))
    with torch.no_grad():
        sa, sb = ([t.detach().clone() for t in graphed._flat(m._states, [])] for m in (a, b))
        for k in range(3):
            x2 = torch.randn(2, 3, generator=torch.Generator().manual_seed(50 + k))
            same(a(x2)["flow"], b(x2)["flow"])
    la.backward(), lb.backward()
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.equal(p.grad, q.grad)
    same(sa, sb)


def test_weight_updates_hooks_and_capture_failure(fake_capture, monkeypatch):
    a, b = pair()

    def run(k0, n):
        with torch.no_grad():
            for k in range(k0, k0 + n):
                x = torch.randn(2, 3, generator=torch.Generator().manual_seed(k))
                same(a(x)["flow"], b(x)["flow"])

    run(0, 4)
    n_graphs = len(fake_capture)
    with torch.no_grad():
        for m in (a, b):
            m.a.mul_(1.5)  # an in-place update is a new situation: one eager step, then a new capture
    calls = a.calls
    run(10, 3)
    assert len(fake_capture) == n_graphs + 1 and a.calls >= calls + 2
    h = a.lin.register_forward_hook(lambda m, i, o: None)
    before = len(fake_capture)
    run(20, 3)
    assert len(fake_capture) == before  # a model with forward hooks is never captured or replayed
    h.remove()

    def broken(fn):
        raise RuntimeError("operation not permitted when stream is capturing")

    monkeypatch.setattr(fast, "_capture", broken)
    with torch.no_grad():
        for m in (a, b):
            m.a.mul_(0.5)
    run(30, 4)  # the capture fails: the model stays on the launch-by-launch path, results unchanged
    assert a.__dict__.get("_graph_off") and "capturing" in a.__dict__.get("_graph_error", "")
