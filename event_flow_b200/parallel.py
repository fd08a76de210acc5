"""
Data-parallel training step for the hot path: batch shards over the ranks (one process per GPU), replicated
parameters, ONE all-reduce(SUM) of the flat fp32 gradient buffer per loss window over NCCL (NVLink / NVSwitch), then
gradient-norm clipping and Adam identically on every rank in two fused kernels (ef_grad_sqnorm + ef_clip_adam).
Replaces, for the DP case, train_flow.py:154-163 (loss.backward() is still the caller's; clip_grad_norm_ + optimizer.step()
+ zero_grad() are this class).  SUM, not MEAN: the reference loss is a sum over the batch (loss/flow.py:226,259).
"""
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


class DataParallelTrainer:
    """
    :param kernels: object with grad_sqnorm / clip_adam (default: the CUDA library).  Tests inject a stand-in so that the
                    host logic -- flat buffers, gradient views, all-reduce(SUM), step counting -- runs under gloo on CPU;
                    without it the model must live on a CUDA device (there is no CPU path in the product).
    """

    def __init__(self, model, lr=2e-4, clip_grad=100.0, betas=(0.9, 0.999), eps=1e-8, process_group=None, kernels=None):
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
            p.grad = self.flat_grad[o:o + k].view(p.shape)
            o += k
        self.n, self.lr, self.clip, self.betas, self.eps = n, lr, clip_grad, betas, eps
        self.group = process_group
        self.step_count = 0
        self.model = model
        # the parameters moved into flat_param: drop everything of the model's fast path that is keyed on the old pointers
        # (cached step graphs, prepared backward calls, weight images)
        from . import fast

        fast.invalidate_pointers(model)
        # the fast path may accumulate its parameter gradients straight into flat_grad (fast._grad_sink_views checks the layout every window)
        model.__dict__["_grad_sink"] = self.flat_grad

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
        self.reduce_gradients()
        self.apply_gradients()

    def reduce_gradients(self):
        """First half of step(): fold re-bound .grad tensors back into the flat buffer, ONE all-reduce(SUM) over the ranks."""
        for p in self.params:  # autograd may have re-bound .grad (e.g. first backward after set_to_none); fold it back
            if p.grad is not None and p.grad.data_ptr() != self._view_of(p).data_ptr():
                self._view_of(p).copy_(p.grad)
                p.grad = self._view_of(p)
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
        from . import fast, ops

        fast.invalidate_weights(self.model)  # the kernel wrote the parameters behind torch's version counters
        ops.invalidate_weight_images()

    def _view_of(self, p):
        if not hasattr(self, "_offsets"):
            self._offsets, o = {}, 0
            for q in self.params:
                self._offsets[id(q)] = o
                o += q.numel()
        o = self._offsets[id(p)]
        return self.flat_grad[o:o + p.numel()].view(p.shape)
