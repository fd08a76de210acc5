"""
Spiking conv cells with the constructor signatures, parameter names (state_dict keys), initialisers and forward contract
of models/spiking_submodules.py (ConvLIF :24, ConvPLIF :129, ConvALIF :230, ConvXLIF :337, ConvLIFRecurrent :438,
ConvPLIFRecurrent :554, ConvALIFRecurrent :660, ConvXLIFRecurrent :771).  forward() is ONE fused CUDA kernel
(ef_lif_conv_fwd) instead of conv2d + ~15 pointwise ops; backward is ef_lif_conv_bwd.
"""
import math

import torch
import torch.nn as nn

from .. import ops
from . import spiking_util as spiking


class _SpikingConvCell(nn.Module):
    neuron = None  # "lif" | "plif" | "alif" | "xlif"
    recurrent = False

    def _build(self, input_size, hidden_size, kernel_size, stride, activation, act_width, leak_group, thresh_group, learn_leak,
               learn_thresh, hard_reset, detach, norm):
        padding = kernel_size // 2
        self.input_size, self.hidden_size, self.stride = input_size, hidden_size, stride
        # parameter creation order follows the reference so that the same torch seed gives the same initial values
        self.ff = nn.Conv2d(input_size, hidden_size, kernel_size, stride=stride, padding=padding, bias=False)
        if self.recurrent:
            self.rec = nn.Conv2d(hidden_size, hidden_size, kernel_size, padding=padding, bias=False)
        for group, learn in ((leak_group, learn_leak), (thresh_group, learn_thresh)):
            for name, (mu, sd) in group:
                value = torch.randn(hidden_size, 1, 1) * sd + mu
                if learn:
                    setattr(self, name, nn.Parameter(value))
                else:
                    self.register_buffer(name, value)
        nn.init.uniform_(self.ff.weight, -math.sqrt(1 / input_size), math.sqrt(1 / input_size))
        if self.recurrent:
            nn.init.uniform_(self.rec.weight, -math.sqrt(1 / hidden_size), math.sqrt(1 / hidden_size))
        assert isinstance(activation, str), "Spiking neurons need a valid activation, see models/spiking_util.py for choices"
        self.spike_fn = getattr(spiking, activation)
        self.activation = activation
        self.register_buffer("act_width", torch.tensor(act_width))
        self.hard_reset = hard_reset
        self.detach = detach
        # normalisation options of the reference (spiking_submodules.py:86-94): weight normalisation of the feed-forward convolution
        # (parameters ff.weight_g / ff.weight_v) or a group norm of the INPUT; anything else means none
        names = ("norm_ff", "norm_rec") if self.recurrent else ("norm",)  # attribute names of the reference (:86-94, :501-514)
        if self.neuron == "lif":
            for n in names:
                setattr(self, n, None)
        if self.neuron != "lif":
            pass  # the reference's PLIF / ALIF / XLIF cells accept `norm` and ignore it (:153-227, 262-334, 361-435 have no norm code)
        elif norm == "weight":
            self.ff = nn.utils.weight_norm(self.ff)
            if self.recurrent:
                self.rec = nn.utils.weight_norm(self.rec)
        elif norm == "group":
            setattr(self, names[0], nn.GroupNorm(min(1, input_size // 4), input_size))
            if self.recurrent:
                self.norm_rec = nn.GroupNorm(min(1, hidden_size // 4), hidden_size)

    def __getattr__(self, name):
        # Cells unpickled from a checkpoint the REFERENCE wrote (utils/utils.py:19-20 pickles whole modules; the aliased module
        # paths make them resolve to these classes) carry the reference's attributes only: derive the two this package adds.
        if name == "stride":
            return self.ff.stride[0]
        if name == "activation":
            fn = self.spike_fn
            return getattr(fn, "name", getattr(fn, "__name__", "arctanspike"))
        return super().__getattr__(name)

    def _width(self):
        """float(act_width) without a device-to-host copy per call: the buffer's value is read once per version."""
        w = self.act_width
        hit = self.__dict__.get("_width_cache")
        if hit is None or hit[0] != (w._version, w.data_ptr()):
            hit = self.__dict__["_width_cache"] = ((w._version, w.data_ptr()), float(w))
        return hit[1]

    @staticmethod
    def _kernel_of(conv):
        """A convolution's kernel; with weight normalisation g * v / ||v|| (what nn.utils.weight_norm's hook computes when the conv is
        called -- it never is here: the convolution runs inside the fused kernel)."""
        if "weight_g" in conv._parameters:
            return torch._weight_norm(conv.weight_v, conv.weight_g, 0)
        return conv.weight

    def plain(self):
        """No normalisation, detached reset: what the fused fast paths (fast.py, fast_unet.py) implement."""
        return (getattr(self, "detach", True) and "weight_g" not in self.ff._parameters
                and all(self._modules.get(n) is None for n in ("norm", "norm_ff", "norm_rec")))

    def forward(self, input_, prev_state, residual=0):
        chan = {n: getattr(self, n) for n in ops.param_names(self.neuron)}
        x_kind = self.__dict__.get("_x_kind")  # set by the model that owns the cell when it knows what the cell's input is
        norm_in = self._modules.get("norm_ff" if self.recurrent else "norm")
        if norm_in is not None:  # group norm of the input (:97-98, :518-519)
            input_, x_kind = norm_in(input_), None
        norm_rec = self._modules.get("norm_rec")
        if norm_rec is not None:  # the recurrent cells normalise the previous spikes -- for the recurrent current AND the reset term (:528-529)
            if prev_state is None:
                h, w = (input_.shape[2] - 1) // self.stride + 1, (input_.shape[3] - 1) // self.stride + 1
                prev_state = input_.new_zeros((2 if self.neuron == "lif" else 3, input_.shape[0], self.hidden_size, h, w))
            parts = list(prev_state.unbind(0))
            parts[1] = norm_rec(parts[1])
            prev_state = torch.stack(parts)
        return ops.cell_step(
            self.neuron,
            input_,
            prev_state,
            self._kernel_of(self.ff),
            self._kernel_of(self.rec) if self.recurrent else None,
            chan,
            hard_reset=self.hard_reset,
            surrogate=self.activation,
            width=self._width(),
            stride=self.stride,
            residual=residual,
            x_kind=x_kind,
            detach=getattr(self, "detach", True),
        )


class ConvLIF(_SpikingConvCell):
    """models/spiking_submodules.py:24-126."""

    neuron = "lif"

    def __init__(self, input_size, hidden_size, kernel_size, stride=1, activation="arctanspike", act_width=10.0, leak=(-4.0, 0.1),
                 thresh=(0.8, 0.0), learn_leak=True, learn_thresh=True, hard_reset=True, detach=True, norm=None):
        super().__init__()
        self._build(input_size, hidden_size, kernel_size, stride, activation, act_width, [("leak", leak)], [("thresh", thresh)],
                    learn_leak, learn_thresh, hard_reset, detach, norm)


class ConvPLIF(_SpikingConvCell):
    """models/spiking_submodules.py:129-227."""

    neuron = "plif"

    def __init__(self, input_size, hidden_size, kernel_size, stride=1, activation="arctanspike", act_width=10.0, leak_v=(-4.0, 0.1),
                 leak_pt=(-4.0, 0.1), add_pt=(-2.0, 0.1), thresh=(0.8, 0.0), learn_leak=True, learn_thresh=True, hard_reset=True,
                 detach=True, norm=None):
        super().__init__()
        self._build(input_size, hidden_size, kernel_size, stride, activation, act_width,
                    [("leak_v", leak_v), ("leak_pt", leak_pt), ("add_pt", add_pt)], [("thresh", thresh)], learn_leak, learn_thresh,
                    hard_reset, detach, norm)


class ConvALIF(_SpikingConvCell):
    """models/spiking_submodules.py:230-334."""

    neuron = "alif"

    def __init__(self, input_size, hidden_size, kernel_size, stride=1, activation="arctanspike", act_width=10.0, leak_v=(-4.0, 0.1),
                 leak_t=(-4.0, 0.1), t0=(0.01, 0.0), t1=(1.8, 0.0), learn_leak=True, learn_thresh=False, hard_reset=False,
                 detach=True, norm=None):
        super().__init__()
        self._build(input_size, hidden_size, kernel_size, stride, activation, act_width, [("leak_v", leak_v), ("leak_t", leak_t)],
                    [("t0", t0), ("t1", t1)], learn_leak, learn_thresh, hard_reset, detach, norm)


class ConvXLIF(_SpikingConvCell):
    """models/spiking_submodules.py:337-435."""

    neuron = "xlif"

    def __init__(self, input_size, hidden_size, kernel_size, stride=1, activation="arctanspike", act_width=10.0, leak_v=(-4.0, 0.1),
                 leak_pt=(-4.0, 0.1), t0=(0.01, 0.0), t1=(1.8, 0.0), learn_leak=True, learn_thresh=False, hard_reset=False,
                 detach=True, norm=None):
        super().__init__()
        self._build(input_size, hidden_size, kernel_size, stride, activation, act_width, [("leak_v", leak_v), ("leak_pt", leak_pt)],
                    [("t0", t0), ("t1", t1)], learn_leak, learn_thresh, hard_reset, detach, norm)


class ConvLIFRecurrent(_SpikingConvCell):
    """models/spiking_submodules.py:438-551."""

    neuron, recurrent = "lif", True

    def __init__(self, input_size, hidden_size, kernel_size, activation="arctanspike", act_width=10.0, leak=(-4.0, 0.1),
                 thresh=(0.8, 0.0), learn_leak=True, learn_thresh=True, hard_reset=True, detach=True, norm=None):
        super().__init__()
        self._build(input_size, hidden_size, kernel_size, 1, activation, act_width, [("leak", leak)], [("thresh", thresh)],
                    learn_leak, learn_thresh, hard_reset, detach, norm)

    def forward(self, input_, prev_state):
        return super().forward(input_, prev_state)


class ConvPLIFRecurrent(_SpikingConvCell):
    """models/spiking_submodules.py:554-657."""

    neuron, recurrent = "plif", True

    def __init__(self, input_size, hidden_size, kernel_size, activation="arctanspike", act_width=10.0, leak_v=(-4.0, 0.1),
                 leak_pt=(-4.0, 0.1), add_pt=(-2.0, 0.1), thresh=(0.8, 0.0), learn_leak=True, learn_thresh=True, hard_reset=True,
                 detach=True, norm=None):
        super().__init__()
        self._build(input_size, hidden_size, kernel_size, 1, activation, act_width,
                    [("leak_v", leak_v), ("leak_pt", leak_pt), ("add_pt", add_pt)], [("thresh", thresh)], learn_leak, learn_thresh,
                    hard_reset, detach, norm)


class ConvALIFRecurrent(_SpikingConvCell):
    """models/spiking_submodules.py:660-768."""

    neuron, recurrent = "alif", True

    def __init__(self, input_size, hidden_size, kernel_size, activation="arctanspike", act_width=10.0, leak_v=(-4.0, 0.1),
                 leak_t=(-4.0, 0.1), t0=(0.01, 0.0), t1=(1.8, 0.0), learn_leak=True, learn_thresh=False, hard_reset=False,
                 detach=True, norm=None):
        super().__init__()
        self._build(input_size, hidden_size, kernel_size, 1, activation, act_width, [("leak_v", leak_v), ("leak_t", leak_t)],
                    [("t0", t0), ("t1", t1)], learn_leak, learn_thresh, hard_reset, detach, norm)

    def forward(self, input_, prev_state):
        return super().forward(input_, prev_state)


class ConvXLIFRecurrent(_SpikingConvCell):
    """models/spiking_submodules.py:771-875."""

    neuron, recurrent = "xlif", True

    def __init__(self, input_size, hidden_size, kernel_size, activation="arctanspike", act_width=10.0, leak_v=(-4.0, 0.1),
                 leak_pt=(-4.0, 0.1), t0=(0.01, 0.0), t1=(1.8, 0.0), learn_leak=True, learn_thresh=False, hard_reset=False,
                 detach=True, norm=None):
        super().__init__()
        self._build(input_size, hidden_size, kernel_size, 1, activation, act_width, [("leak_v", leak_v), ("leak_pt", leak_pt)],
                    [("t0", t0), ("t1", t1)], learn_leak, learn_thresh, hard_reset, detach, norm)

    def forward(self, input_, prev_state):
        return super().forward(input_, prev_state)


_FF_BLOCKS = {"lif": ConvLIF, "alif": ConvALIF, "plif": ConvPLIF, "xlif": ConvXLIF}
_REC_BLOCKS = {"lif": ConvLIFRecurrent, "alif": ConvALIFRecurrent, "plif": ConvPLIFRecurrent, "xlif": ConvXLIFRecurrent}


class SpikingRecurrentConvLayer(nn.Module):
    """
    Convolution followed by a recurrent convolutional block, both spiking (models/spiking_submodules.py:878-930): the
    encoder stage of the spiking U-Net.  Two fused conv+neuron kernels; state = stack([ff_state, rec_state]).
    """

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, recurrent_block_type="lif", activation_ff="arctanspike",
                 activation_rec="arctanspike", **kwargs):
        super().__init__()
        assert recurrent_block_type in ["lif", "alif", "plif", "xlif"]
        kwargs.pop("spiking_feedforward_block_type", None)
        self.conv = _FF_BLOCKS[recurrent_block_type](in_channels, out_channels, kernel_size, stride, activation_ff, **kwargs)
        self.recurrent_block = _REC_BLOCKS[recurrent_block_type](out_channels, out_channels, kernel_size, activation=activation_rec, **kwargs)

    def forward(self, x, prev_state):
        if prev_state is None:
            prev_state = [None, None]
        ff, rec = prev_state  # unbind op, removes dimension
        x1, ff = self.conv(x, ff)
        x2, rec = self.recurrent_block(x1, rec)
        return x2, torch.stack([ff, rec])


class SpikingResidualBlock(nn.Module):
    """Spiking residual block (models/spiking_submodules.py:933-975); the residual is added inside the second cell's kernel."""

    def __init__(self, in_channels, out_channels, stride=1, spiking_feedforward_block_type="lif", activation="arctanspike", **kwargs):
        super().__init__()
        assert spiking_feedforward_block_type in ["lif", "alif", "plif", "xlif"]
        block = _FF_BLOCKS[spiking_feedforward_block_type]
        self.conv1 = block(in_channels, out_channels, kernel_size=3, stride=stride, activation=activation, **kwargs)
        self.conv2 = block(out_channels, out_channels, kernel_size=3, stride=1, activation=activation, **kwargs)

    def forward(self, x, prev_state):
        if prev_state is None:
            prev_state = [None, None]
        conv1, conv2 = prev_state  # unbind op, removes dimension
        residual = x
        x1, conv1 = self.conv1(x, conv1)
        x2, conv2 = self.conv2(x1, conv2, residual=residual)  # add res inside
        return x2, torch.stack([conv1, conv2])


class SpikingUpsampleConvLayer(nn.Module):
    """Bilinear x2 upsampling + spiking conv cell (models/spiking_submodules.py:978-1013): the decoder stage of the spiking U-Net."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, spiking_feedforward_block_type="lif", activation="arctanspike",
                 **kwargs):
        super().__init__()
        assert spiking_feedforward_block_type in ["lif", "alif", "plif", "xlif"]
        self.conv2d = _FF_BLOCKS[spiking_feedforward_block_type](in_channels, out_channels, kernel_size, stride=stride, activation=activation,
                                                                 **kwargs)

    def forward(self, x, prev_state):
        x_up = ops.upsample_bilinear2x(x)
        x1, state = self.conv2d(x_up, prev_state)
        return x1, state


class SpikingTransposedConvLayer(nn.Module):
    """models/spiking_submodules.py:1016-1062 (use_upsample_conv=False); no shipped config selects it."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("event_flow_b200: SpikingTransposedConvLayer (use_upsample_conv=False) is not on the CUDA path")
