"""
Spike functions with the names, signatures and defaults of models/spiking_util.py:96-109.  On the model path the Heaviside
forward and the surrogate backward are fused into the conv + neuron kernels (csrc/common.cuh: neuron_update, surrogate_grad);
the cells only read `.name` of these objects.  Called stand-alone -- `arctanspike(x, thresh, width)` like the reference's
cells do -- they run two small kernels of libeventflow.so (ef_spike_fwd / ef_spike_bwd) with the same maths.
"""
import torch

from .. import _lib as L


class _SpikeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, thresh, kind, width):
        if not x.is_cuda:
            raise L.EventFlowError("event_flow_b200 has no CPU path: tensors must live on a CUDA device (got %s)" % x.device)
        x = x.detach().float().contiguous()
        th = torch.as_tensor(thresh, dtype=torch.float32, device=x.device).detach()
        ctx.th_shape = th.shape
        if th.numel() == 1:
            mode, C, hw, th = 0, 1, 1, th.reshape(1).contiguous()
        elif x.dim() == 4 and th.numel() == x.shape[1] and tuple(th.shape[-3:]) in ((x.shape[1], 1, 1),):
            mode, C, hw, th = 1, x.shape[1], x.shape[2] * x.shape[3], th.reshape(-1).contiguous()
        else:
            mode, C, hw, th = 2, 1, 1, th.expand_as(x).contiguous()
        z = torch.empty_like(x)
        L.LAUNCHES += 1
        L.check(L.lib().ef_spike_fwd(L.ptr(x), L.ptr(th), mode, C, hw, x.numel(), L.ptr(z), L.stream()), "ef_spike_fwd")
        ctx.save_for_backward(x, th)
        ctx.meta = (mode, C, hw, kind, float(width))
        return z

    @staticmethod
    def backward(ctx, g):
        x, th = ctx.saved_tensors
        mode, C, hw, kind, width = ctx.meta
        g = g.float().contiguous()
        g_x = torch.empty_like(x)
        L.LAUNCHES += 1
        L.check(L.lib().ef_spike_bwd(L.ptr(x), L.ptr(th), mode, C, hw, x.numel(), L.ptr(g), kind, width, L.ptr(g_x), L.stream()), "ef_spike_bwd")
        g_th = None
        if ctx.needs_input_grad[1]:  # d(x - thresh)/d thresh = -1, summed over the broadcast dimensions
            g_th = (-g_x).sum_to_size(ctx.th_shape) if len(ctx.th_shape) else (-g_x).sum()
        return g_x, g_th, None, None


class _SpikeFunction:
    """Callable with the reference's signature fn(x, thresh=1.0, width=<default>)."""

    def __init__(self, name, default_width):
        self.name, self.default_width = name, default_width
        self.__name__ = name

    def __call__(self, x, thresh=1.0, width=None):
        width = self.default_width if width is None else float(width)
        if not torch.is_tensor(thresh):
            thresh = torch.tensor(float(thresh))
        return _SpikeFn.apply(x, thresh, L.SURROGATE_CODES[self.name], width)

    def __repr__(self):
        return f"<spike fn {self.name}>"


superspike = _SpikeFunction("superspike", 10.0)
mgspike = _SpikeFunction("mgspike", 0.5)
trianglespike = _SpikeFunction("trianglespike", 1.0)
arctanspike = _SpikeFunction("arctanspike", 10.0)
