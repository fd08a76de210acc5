"""
Data-parallel training step for the hot path: batch shards over the ranks (one process per GPU), replicated
parameters, ONE all-reduce(SUM) of the flat fp32 gradient buffer per loss window over NCCL (NVLink / NVSwitch), then
gradient-norm clipping and Adam identically on every rank in two fused kernels (ef_grad_sqnorm + ef_clip_adam).
Replaces, for the DP case, train_flow.py:154-163 (loss.backward() is still the caller's; clip_grad_norm_ + optimizer.step()
+ zero_grad() are this class).  SUM, not MEAN: the reference loss is a sum over the batch (loss/flow.py:226,259).
"""
import os

import torch
import torch.distributed as dist

from . import _lib as L


class _LibKernels:
    """The two optimiser kernels of libeventflow.so (ef_grad_sqnorm, ef_clip_adam) on the current CUDA stream."""

    @staticmethod
    def grad_sqnorm(flat_grad, n, sqnorm):
        L.check(L.lib().ef_grad_sqnorm(L.ptr(flat_grad), n, L.ptr(sqnorm), L.stream()), "ef_grad_sqnorm")
        L.LAUNCHES += 1

    @staticmethod
    def clip_adam(flat_param, flat_grad, m, v, n, sqnorm, clip, lr, beta1, beta2, eps, step):
        L.check(L.lib().ef_clip_adam(L.ptr(flat_param), L.ptr(flat_grad), L.ptr(m), L.ptr(v), n, L.ptr(sqnorm), clip, lr, beta1, beta2, eps,
                                     step, L.stream()), "ef_clip_adam")
        L.LAUNCHES += 1


class _IpcBuffer:
    """A cudaMalloc'ed buffer of libeventflow.so (ef_ipc_alloc) that torch wraps without owning: __cuda_array_interface__ of raw bytes."""

    def __init__(self, nbytes):
        import ctypes as C

        self.nbytes, ptr, handle = nbytes, C.c_void_p(), C.create_string_buffer(64)
        L.check(L.lib().ef_ipc_alloc(nbytes, C.byref(ptr), handle), "ef_ipc_alloc")
        self.ptr, self.handle = int(ptr.value), bytes(handle.raw)

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}


class _PeerStep:
    """
    The optimiser step of all ranks as one kernel each over NVLink peer memory (csrc/dp_step.cu).  Every rank allocates its flat gradient
    buffer plus a few signal words as ONE cudaMalloc of the library (ef_ipc_alloc), the 64-byte CUDA IPC handles go round with
    all_gather_object, and each rank maps its peers' buffers with lazy peer access (ef_ipc_open).  The trainer's gradient views are re-bound
    to the shared buffer; then `ef_dp_step` = wait for all gradients -> sum them in rank order -> clip -> Adam -> zero the own gradient.
    Construction ends with a self test on copies of the optimiser state against NCCL + the two-kernel path; any failure raises, and the
    trainer keeps using NCCL.
    """

    SIGNAL_BYTES = 256

    def __init__(self, tr):
        import ctypes as C

        self.tr = tr
        self.world, self.rank = tr.world_size, dist.get_rank(tr.group)
        if self.world > L.EF_DP_MAX_RANKS:
            raise RuntimeError(f"at most {L.EF_DP_MAX_RANKS} ranks")
        dev = tr.flat_grad.device
        # one GPU per rank: kernels of two processes on ONE device are time-sliced, a rank spinning for its peer would stall it
        where = [None] * self.world
        dist.all_gather_object(where, (os.uname().nodename, str(getattr(torch.cuda.get_device_properties(dev), "uuid", dev.index))), group=tr.group)
        if len(set(where)) != self.world or len({w[0] for w in where}) != 1:
            raise RuntimeError("ranks share a device or span several nodes")
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.grid = int(L.lib().ef_dp_step_grid(tr.n))
        self.scratch = torch.zeros(self.grid, dtype=torch.float32, device=dev)
        self.signal_offset = (tr.n * 4 + 255) // 256 * 256
        with torch.cuda.device(dev):
            err = None
            try:
                self.buffer = _IpcBuffer(self.signal_offset + self.SIGNAL_BYTES)
            except Exception as exc:
                err, self.buffer = repr(exc), None
            handles = [None] * self.world
            dist.all_gather_object(handles, (err, self.buffer.handle if self.buffer else None), group=tr.group)
            bad = [h[0] for h in handles if h[0]]
            if bad:  # every rank leaves together: nobody waits in a later collective
                raise RuntimeError(bad[0])
            self.grad_ptrs, self.signal_ptrs, err = [0] * self.world, [0] * self.world, None
            for r, (_, h) in enumerate(handles):
                base = self.buffer.ptr
                if r != self.rank:
                    q = C.c_void_p()
                    try:
                        L.check(L.lib().ef_ipc_open(h, C.byref(q)), "ef_ipc_open")
                    except Exception as exc:
                        err = repr(exc)
                        break
                    base = int(q.value)
                self.grad_ptrs[r], self.signal_ptrs[r] = base, base + self.signal_offset
            errs = [None] * self.world
            dist.all_gather_object(errs, err, group=tr.group)
            if any(errs):
                raise RuntimeError(next(e for e in errs if e))
        raw = torch.as_tensor(self.buffer, device=dev)
        self.own_grad = raw[:tr.n * 4].view(torch.float32)
        self.epoch = 0
        self._self_test()
        tr._bind_grad(self.own_grad)

    def close(self):
        """
        Hand the gradient views back to an ordinary torch buffer and release the shared memory: the peers' mappings first, then -- once
        every rank has closed its mappings of it -- this rank's own allocation.  Collective: every rank calls it.
        """
        tr, dev = self.tr, self.own_grad.device
        torch.cuda.synchronize(dev)
        dist.barrier(group=tr.group)  # nobody is still inside a step that reads a peer's buffer
        tr._bind_grad(torch.empty(tr.n, device=dev, dtype=torch.float32))  # (copies the current gradient over)
        tr.fused = None
        torch.cuda.synchronize(dev)
        with torch.cuda.device(dev):
            for r, ptr in enumerate(self.grad_ptrs):
                if r != self.rank:
                    L.check(L.lib().ef_ipc_close(ptr), "ef_ipc_close")
            dist.barrier(group=tr.group)
            self.own_grad = None
            L.check(L.lib().ef_ipc_free(self.buffer.ptr), "ef_ipc_free")
        self.buffer = None

    def _launch(self, param, m, v, step, graceful):
        tr = self.tr
        self.epoch += 1
        p = L.DpStepParams()
        p.world, p.rank, p.n, p.step = self.world, self.rank, tr.n, int(step)
        p.clip, p.lr, p.beta1, p.beta2, p.eps = float(tr.clip) if tr.clip else 0.0, tr.lr, tr.betas[0], tr.betas[1], tr.eps
        p.epoch, p.epoch_launches, p.graceful, p.grid_expected = self.epoch, self.epoch, int(graceful), self.grid
        # how long a rank waits for late peers inside the kernel: the self test gives up quickly, training waits like NCCL's watchdog would
        p.timeout_ms = 5000 if graceful else int(float(os.environ.get("EF_DP_TIMEOUT_S", "600")) * 1000)
        for r in range(self.world):
            p.grads[r], p.signals[r] = self.grad_ptrs[r], self.signal_ptrs[r]
        p.param, p.m, p.v = L.ptr(param), L.ptr(m), L.ptr(v)
        p.sqnorm, p.scratch, p.status = L.ptr(tr.sqnorm), L.ptr(self.scratch), L.ptr(self.status)
        L.call("ef_dp_step", p)

    def step(self, param, m, v, step):
        self._launch(param, m, v, step, graceful=False)

    def _self_test(self):
        """One fused step on COPIES of the state with rank-dependent gradients, against all_reduce + ef_grad_sqnorm + ef_clip_adam."""
        tr = self.tr
        gen = torch.Generator(device="cpu").manual_seed(1234 + self.rank)
        g = torch.randn(tr.n, generator=gen).to(self.own_grad.device)
        self.own_grad.copy_(g)
        pa, ma, va = tr.flat_param.clone(), torch.rand_like(tr.m) * 0.1, torch.rand_like(tr.v) * 0.01
        pb, mb, vb = pa.clone(), ma.clone(), va.clone()
        ref = g.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM, group=tr.group)
        sq = torch.zeros(1, device=g.device)
        _LibKernels.grad_sqnorm(ref, tr.n, sq)
        _LibKernels.clip_adam(pb, ref, mb, vb, tr.n, sq, float(tr.clip) if tr.clip else 0.0, tr.lr, tr.betas[0], tr.betas[1], tr.eps, 3)
        self._launch(pa, ma, va, 3, graceful=True)
        torch.cuda.synchronize()
        ok = int(self.status.item()) == 0
        ok = ok and (pa - pb).abs().max().item() <= 1e-6 * pb.abs().max().item() + 1e-9
        ok = ok and abs(tr.sqnorm.item() - sq.item()) <= 1e-5 * sq.item() and self.own_grad.abs().max().item() == 0.0
        flag = torch.tensor([1 if ok else 0], device=g.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=tr.group)  # all ranks take the same path
        self.own_grad.zero_()
        tr.sqnorm.zero_()
        if int(flag.item()) != 1:
            raise RuntimeError(f"ef_dp_step self test failed on some rank (status {int(self.status.item())})")


class DataParallelTrainer:
    """
    :param kernels: object with grad_sqnorm / clip_adam (default: the CUDA library).  Tests inject a stand-in so that the
                    host logic -- flat buffers, gradient views, all-reduce(SUM), step counting -- runs under gloo on CPU;
                    without it the model must live on a CUDA device (there is no CPU path in the product).
    """

    def __init__(self, model, lr=2e-4, clip_grad=100.0, betas=(0.9, 0.999), eps=1e-8, process_group=None, kernels=None, fused=None):
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("model has no trainable parameters")
        dev = self.params[0].device
        if kernels is None and dev.type != "cuda":
            raise L.EventFlowError("DataParallelTrainer needs the model on a CUDA device (no CPU path)")
        self.kernels = kernels if kernels is not None else _LibKernels
        n = sum(p.numel() for p in self.params)
        self.flat_param = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.sqnorm = torch.zeros(1, device=dev, dtype=torch.float32)
        o = 0
        for p in self.params:  # parameters and their .grad become views of the flat buffers
            k = p.numel()
            self.flat_param[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[o:o + k].view(p.shape)
            o += k
        self.model = model
        self._bind_grad(self.flat_grad)
        self.n, self.lr, self.clip, self.betas, self.eps = n, lr, clip_grad, betas, eps
        self.group = process_group
        self.step_count = 0
        self.fused_error = None
        # the parameters moved into flat_param: drop everything of the model's fast path that is keyed on the old pointers
        # (cached step graphs, prepared backward calls, weight images)
        from . import fast

        fast.invalidate_pointers(model)
        # more than one rank on CUDA: the whole step as ONE kernel over NVLink peer memory (ef_dp_step), NCCL as the fallback
        self.fused = None
        if fused is not False and kernels is None and self.world_size > 1 and os.environ.get("EF_DP_FUSED", "1") != "0":
            try:
                self.fused = _PeerStep(self)
            except Exception as exc:  # IPC not available (allocator mode, container limits, ...): the NCCL path is always there
                self.fused, self.fused_error = None, repr(exc)

    def close(self):
        """Release what the fused peer-memory step shares between the ranks (collective; the trainer keeps working over NCCL)."""
        if self.fused is not None:
            self.fused.close()

    def _bind_grad(self, flat_grad):
        """Make `flat_grad` (n fp32 values, current contents kept) the gradient buffer: every p.grad becomes a view of it."""
        if flat_grad is not self.flat_grad:
            flat_grad.copy_(self.flat_grad)
        self.flat_grad, o = flat_grad, 0
        self.__dict__.pop("_offsets", None)
        for p in self.params:
            k = p.numel()
            p.grad = flat_grad[o:o + k].view(p.shape)
            o += k
        # the fast path may accumulate its parameter gradients straight into flat_grad (fast._grad_sink_views checks the layout every window)
        self.model.__dict__["_grad_sink"] = flat_grad

    @property
    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def zero_grad(self):
        self.flat_grad.zero_()

    def grad_norm(self):
        """Global L2 norm of the (already reduced) gradient as a 0-d tensor -- what clip_grad_norm_ returns."""
        return self.sqnorm.sqrt()[0]

    def step(self):
        """all-reduce(SUM) -> clip -> Adam -> zero grads.  Call after loss.backward()."""
        if self.fused is not None:
            self._fold_rebound_grads()
            self.step_count += 1
            self.fused.step(self.flat_param, self.m, self.v, self.step_count)
            self._after_update()
            return
        self.reduce_gradients()
        self.apply_gradients()

    def _fold_rebound_grads(self):
        for p in self.params:  # autograd may have re-bound .grad (e.g. first backward after set_to_none); fold it back
            if p.grad is not None and p.grad.data_ptr() != self._view_of(p).data_ptr():
                self._view_of(p).copy_(p.grad)
                p.grad = self._view_of(p)

    def _after_update(self):
        from . import fast, ops

        fast.invalidate_weights(self.model)  # the kernel wrote the parameters behind torch's version counters
        ops.invalidate_weight_images()

    def reduce_gradients(self):
        """First half of step(): fold re-bound .grad tensors back into the flat buffer, ONE all-reduce(SUM) over the ranks."""
        self._fold_rebound_grads()
        if self.world_size > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)

    def apply_gradients(self):
        """Second half of step(): global-norm clip + Adam on the flat buffers (two kernels), then zero the gradients."""
        self.step_count += 1
        self.sqnorm.zero_()
        self.kernels.grad_sqnorm(self.flat_grad, self.n, self.sqnorm)
        self.kernels.clip_adam(self.flat_param, self.flat_grad, self.m, self.v, self.n, self.sqnorm, float(self.clip) if self.clip else 0.0,
                               self.lr, self.betas[0], self.betas[1], self.eps, self.step_count)
        self.flat_grad.zero_()
        self._after_update()

    def _view_of(self, p):
        if not hasattr(self, "_offsets"):
            self._offsets, o = {}, 0
            for q in self.params:
                self._offsets[id(q)] = o
                o += q.numel()
        o = self._offsets[id(p)]
        return self.flat_grad[o:o + p.numel()].view(p.shape)
