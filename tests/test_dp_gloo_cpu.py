"""
CPU, world_size 2 over gloo: the host side of the data-parallel training step -- the code under test is
event_flow_b200.parallel.DataParallelTrainer and event_flow_b200.train.train_windows themselves.  The two optimiser kernels
(ef_grad_sqnorm, ef_clip_adam) cannot run without a GPU, so a torch stand-in with their documented semantics
(include/eventflow.h) is injected through the trainer's `kernels` hook; everything else -- flat parameter / gradient buffers,
gradient views, the ONE all-reduce(SUM) per step (no division by the world size: the loss sums over the batch,
loss/flow.py:226,259), clip after the reduce, step counting, zeroing -- is the product code.  The GPU twin of this test
(NCCL, real kernels, LIFFireNet) is tests/test_gpu_dp.py.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class TorchKernels:
    """ef_grad_sqnorm / ef_clip_adam restated with torch ops (test stand-in, not product code)."""

    @staticmethod
    def grad_sqnorm(flat_grad, n, sqnorm):
        sqnorm += (flat_grad.double() ** 2).sum().float()

    @staticmethod
    def clip_adam(flat_param, flat_grad, m, v, n, sqnorm, clip, lr, beta1, beta2, eps, step):
        scale = min(1.0, clip / (float(sqnorm.sqrt()) + 1e-6)) if clip > 0 else 1.0
        g = flat_grad * scale
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        mhat, vhat = m / (1 - beta1 ** step), v / (1 - beta2 ** step)
        flat_param.sub_(lr * mhat / (vhat.sqrt() + eps))


def make_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(2, 6, 3, padding=1), torch.nn.Tanh(), torch.nn.Conv2d(6, 2, 1))


def batch(step):
    g = torch.Generator().manual_seed(100 + step)
    return torch.randn(4, 2, 10, 12, generator=g)


def run_steps(model, trainer, shard, n_steps, set_to_none=False):
    for it in range(n_steps):
        x = batch(it)[shard]
        loss = (model(x) ** 2).sum()  # a SUM over the batch, like the reference loss
        if set_to_none and it == 1:   # autograd re-binds .grad to fresh tensors: the trainer must fold them back into its flat buffer
            for p in model.parameters():
                p.grad = None
        loss.backward()
        trainer.step()
    return trainer


def worker(rank, world, port, out):
    from event_flow_b200.parallel import DataParallelTrainer

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = make_model()
    tr = DataParallelTrainer(model, lr=1e-2, clip_grad=5.0, kernels=TorchKernels)
    assert tr.world_size == world
    # parameters and gradients are views of the flat buffers
    o = 0
    for p in tr.params:
        assert p.data.data_ptr() == tr.flat_param[o:o + p.numel()].data_ptr() and p.grad.data_ptr() == tr.flat_grad[o:o + p.numel()].data_ptr()
        o += p.numel()
    run_steps(model, tr, slice(rank * 2, rank * 2 + 2), 3, set_to_none=True)
    gathered = [torch.empty_like(tr.flat_param) for _ in range(world)]
    dist.all_gather(gathered, tr.flat_param)
    if rank == 0:
        out["param"], out["same"] = tr.flat_param.clone(), all(torch.equal(g, tr.flat_param) for g in gathered)
        out["norm"], out["steps"], out["grad_zero"] = tr.grad_norm().item(), tr.step_count, bool(tr.flat_grad.abs().max() == 0)
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_trainer_equals_one_process_on_the_whole_batch():
    from event_flow_b200.parallel import DataParallelTrainer

    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(2, free_port(), out), nprocs=2, join=True)
    model = make_model()
    tr = run_steps(model, DataParallelTrainer(model, lr=1e-2, clip_grad=5.0, kernels=TorchKernels), slice(0, 4), 3)
    assert out["same"], "ranks diverged"
    assert out["steps"] == 3 and out["grad_zero"]
    torch.testing.assert_close(out["param"], tr.flat_param, rtol=1e-5, atol=1e-7)
    assert abs(out["norm"] - tr.grad_norm().item()) <= 1e-5 * tr.grad_norm().item()  # the clip saw the REDUCED gradient
    # and the flat-buffer Adam is torch's Adam + clip_grad_norm_
    ref = make_model()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-2)
    for it in range(3):
        (ref(batch(it)) ** 2).sum().backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 5.0)
        opt.step()
        opt.zero_grad()
    flat_ref = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
    torch.testing.assert_close(tr.flat_param, flat_ref, rtol=1e-4, atol=1e-6)


def test_trainer_without_kernels_refuses_cpu_models():
    import pytest

    from event_flow_b200._lib import EventFlowError
    from event_flow_b200.parallel import DataParallelTrainer

    with pytest.raises(EventFlowError):
        DataParallelTrainer(make_model())


def test_train_windows_follows_the_reference_loop():
    """train_flow.py:97-171: new_seq resets loss / states / grads, the loss fires at window_loss events, then step, detach, reset."""
    from event_flow_b200.train import train_windows

    log = []

    class Loader:
        def __init__(self):
            self.new_seq = False

        def __iter__(self):
            for i in range(9):
                self.new_seq = i in (0, 5)  # a second recording starts in the middle of a loss window
                yield {"event_voxel": torch.zeros(1), "event_cnt": torch.zeros(1), "event_list": torch.zeros(1, 100, 4),
                       "event_list_pol_mask": torch.zeros(1, 100, 2), "event_mask": torch.zeros(1)}

    class Model:
        def train(self):
            log.append("train")

        def reset_states(self):
            log.append("reset_states")

        def detach_states(self):
            log.append("detach_states")

        def __call__(self, vox, cnt):
            log.append("forward")
            return {"flow": [torch.zeros(1)]}

    class Loss:
        num_events = 0

        def reset(self):
            log.append("loss.reset")
            self.num_events = 0

        def event_flow_association(self, flow, ev, pm, mask):
            self.num_events += ev.shape[1]

        def __call__(self):
            log.append("loss()")
            return torch.zeros((), requires_grad=True) + 1.0

    class Trainer:
        def zero_grad(self):
            log.append("zero_grad")

        def step(self):
            log.append("step")

    losses = train_windows(Model(), Loss(), Trainer(), Loader(), window_loss=300, n_windows=2)
    assert losses == [1.0, 1.0]
    assert log == ["train",
                   "loss.reset", "reset_states", "zero_grad", "forward", "forward", "forward", "loss()", "step", "detach_states", "loss.reset",
                   "forward", "forward",                                        # 200 events of the next window ...
                   "loss.reset", "reset_states", "zero_grad", "forward",        # ... dropped by new_seq (train_flow.py:100-105)
                   "forward", "forward", "loss()", "step", "detach_states", "loss.reset"]


def test_train_windows_staged_runs_whole_windows_and_drops_partial_ones():
    """staged=True: the items of a loss window are staged and handed to model.forward_window; new_seq drops the staged steps."""
    from event_flow_b200.train import train_windows

    log = []

    class Loader:
        def __init__(self):
            self.new_seq = False

        def __iter__(self):
            for i in range(9):
                self.new_seq = i in (0, 5)
                yield {"event_voxel": torch.full((1,), float(i)), "event_cnt": torch.zeros(1), "event_list": torch.zeros(1, 100, 4),
                       "event_list_pol_mask": torch.zeros(1, 100, 2), "event_mask": torch.zeros(1)}

    class Model:
        def train(self):
            pass

        def reset_states(self):
            log.append("reset_states")

        def detach_states(self):
            log.append("detach_states")

        def forward_window(self, vox, cnt):
            log.append(("forward_window", vox.flatten().tolist()))
            return [{"flow": [torch.zeros(1)]} for _ in range(vox.shape[0])]

    class Loss:
        num_events = 0

        def reset(self):
            self.num_events = 0

        def event_flow_association(self, flow, ev, pm, mask):
            self.num_events += ev.shape[1]

        def __call__(self):
            log.append(("loss()", self.num_events))
            return torch.zeros((), requires_grad=True) + 1.0

    class Trainer:
        def zero_grad(self):
            pass

        def step(self):
            log.append("step")

    losses = train_windows(Model(), Loss(), Trainer(), Loader(), window_loss=300, n_windows=2, staged=True)
    assert losses == [1.0, 1.0]
    assert log == ["reset_states", ("forward_window", [0.0, 1.0, 2.0]), ("loss()", 300), "step", "detach_states",
                   "reset_states",  # items 3, 4 were staged and are dropped by the new recording
                   ("forward_window", [5.0, 6.0, 7.0]), ("loss()", 300), "step", "detach_states"]


def test_synthetic_stream_shards_are_disjoint_and_reproducible():
    from event_flow_b200.train import SyntheticEventStream

    a = SyntheticEventStream(2, 50, (16, 16), 2, "cpu", rank=0).host_events(3)
    b = SyntheticEventStream(2, 50, (16, 16), 2, "cpu", rank=1).host_events(3)
    a2 = SyntheticEventStream(2, 50, (16, 16), 2, "cpu", rank=0).host_events(3)
    assert torch.equal(a, a2) and not torch.equal(a, b)
    assert a.shape == (2, 50, 4) and a[:, :, 0].min() == 0 and a[:, :, 0].max() == 1 and set(a[:, :, 3].unique().tolist()) <= {-1.0, 1.0}


def test_rebinding_the_gradient_buffer_keeps_contents_views_and_sink():
    """
    The fused peer-memory step moves the flat gradient into a buffer the peers can map (parallel._PeerStep -> _bind_grad): the gradient
    accumulated so far, the p.grad views, the offsets of _view_of and the model's gradient sink must all follow.
    """
    from event_flow_b200.parallel import DataParallelTrainer

    torch.manual_seed(1)
    model = make_model()
    tr = DataParallelTrainer(model, lr=1e-2, clip_grad=5.0, kernels=TorchKernels)
    x = batch(0)
    (model(x) ** 2).sum().backward()
    before = tr.flat_grad.clone()
    assert before.abs().sum() > 0
    shared = torch.full((tr.n,), 7.0)
    tr._bind_grad(shared)
    assert tr.flat_grad is shared and torch.equal(shared, before)
    assert model.__dict__["_grad_sink"] is shared
    o = 0
    for p in tr.params:
        assert p.grad.data_ptr() == shared[o:o + p.numel()].data_ptr() and p.grad.shape == p.shape
        assert tr._view_of(p).data_ptr() == p.grad.data_ptr()
        o += p.numel()
    (model(x) ** 2).sum().backward()  # autograd accumulates into the new buffer
    assert torch.allclose(shared, 2 * before)
    tr.step()
    assert shared.abs().max().item() == 0.0


def test_graph_replay_bookkeeping_on_nested_states():
    """Pure host logic of event_flow_b200/graphed.py: leaf order, structure signatures and rebuilding of nested recurrent states."""
    from event_flow_b200 import graphed

    a, b, c = torch.zeros(2, 3), torch.ones(4), torch.zeros(1)
    states = [None, a, (b, c), [None, a]]
    leaves = graphed._flat(states, [])
    assert [t is u for t, u in zip(leaves, (a, b, c, a))] == [True] * 4
    sig = graphed._shape_of(states)
    assert sig == graphed._shape_of([None, a.clone(), (b.clone(), c.clone()), [None, a.clone()]])
    assert sig != graphed._shape_of([None, a, (b, c), [a, None]]) and sig != graphed._shape_of([None, a.t(), (b, c), [None, a]])
    fresh = [t + 1 for t in leaves]
    rebuilt = graphed._rebuild(states, iter(fresh))
    assert rebuilt[0] is None and rebuilt[1] is fresh[0] and type(rebuilt[2]) is tuple and rebuilt[2][1] is fresh[2] and rebuilt[3][1] is fresh[3]
    m = torch.nn.Linear(2, 2)
    assert not graphed.usable(m, torch.zeros(1, 2))  # CPU tensors never take the graph path
    holder = type("H", (), {})()
    holder.states = states
    graphed.leave(m, holder, "states")  # no graphs: a no-op
    assert holder.states is states


def test_launcher_builds_every_model_class():
    """tools/train_dp.py --model <any class of the reference's import list>: the per-family neuron / activation settings construct the model."""
    import event_flow_b200.models.model as M
    from tools.train_dp import model_config

    names = ["FireNet", "RNNFireNet", "LeakyFireNet", "FireFlowNet", "LeakyFireFlowNet", "E2VID", "EVFlowNet", "RecEVFlowNet", "LeakyRecEVFlowNet",
             "RNNRecEVFlowNet", "LIFFireNet", "PLIFFireNet", "ALIFFireNet", "XLIFFireNet", "LIFFireFlowNet", "SpikingRecEVFlowNet", "PLIFRecEVFlowNet",
             "ALIFRecEVFlowNet", "XLIFRecEVFlowNet"]
    for name in names:
        model = getattr(M, name)(dict(model_config(name, 5)))
        assert sum(p.numel() for p in model.parameters()) > 0, name
