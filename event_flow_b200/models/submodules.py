"""
ANN layers of models/submodules.py used by the FireNet family: ConvLayer (:12-61), ConvLayer_ (:64-83), ConvGRU (:377-418),
with the reference's constructor signatures, parameter names and initialisers.  Forward passes are CUDA kernels
(ef_conv_ann_fwd / ef_pred_fwd); under autograd the convolution gradients come from ef_conv3x3_bwd (ops._ConvAnn).
"""
import torch
import torch.nn as nn

from .. import ops


_TORCH_ACT = {None: None, "relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}


def _make_norm(norm, channels, momentum=0.1):
    """The reference's normalisation options of the ANN layers (models/submodules.py:45-49): "BN" / "IN" torch modules, else none."""
    if norm == "BN":
        return nn.BatchNorm2d(channels, momentum=momentum)
    if norm == "IN":
        return nn.InstanceNorm2d(channels, track_running_stats=True)
    return None


def _conv_norm_act(x, conv, norm_layer, act, residual=None, stride=1):
    """act(norm(conv3x3(x) + b) + residual).  Without a norm layer everything is ONE fused launch (ef_conv_ann_fwd); with one the kernel
    stops after the bias and the normalisation module / residual / activation follow on its output (models/submodules.py:52-61)."""
    if norm_layer is None:
        return ops.conv_ann(x, conv.weight, conv.bias, act, residual=residual, stride=stride)
    out = norm_layer(ops.conv_ann(x, conv.weight, conv.bias, None, stride=stride))
    if residual is not None:
        out = out + residual
    return out if act is None else _TORCH_ACT[act](out)


class ConvLayer(nn.Module):
    """Convolutional layer: conv + bias [+ BN / IN] + activation (default ReLU) (models/submodules.py:12-61)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, activation="relu", norm=None, BN_momentum=0.1, w_scale=None):
        super().__init__()
        if kernel_size not in (1, 3) or stride not in (1, 2) or (kernel_size == 1 and stride != 1):
            raise NotImplementedError(f"event_flow_b200 ConvLayer: kernel_size={kernel_size}, stride={stride} not on the CUDA path yet")
        if activation not in (None, "relu", "sigmoid", "tanh"):
            raise NotImplementedError(f"event_flow_b200 ConvLayer: activation={activation!r} not on the CUDA path")
        padding = kernel_size // 2
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=norm != "BN")
        if w_scale is not None:
            nn.init.uniform_(self.conv2d.weight, -w_scale, w_scale)
            nn.init.zeros_(self.conv2d.bias)
        self.kernel_size = kernel_size
        self.stride = stride
        self.activation = activation
        self.norm = norm
        layer = _make_norm(norm, out_channels, BN_momentum)
        if layer is not None:
            self.norm_layer = layer

    def __getattr__(self, name):
        # layers unpickled from a checkpoint the REFERENCE wrote keep only its attributes: derive the ones this package adds
        if name in ("kernel_size", "stride"):
            return getattr(self.conv2d, name)[0]
        return super().__getattr__(name)

    def _act_name(self):
        act = self.activation  # a name here; a torch function (torch.relu, torch.tanh, ...) in objects the reference pickled
        return getattr(act, "__name__", act) if callable(act) else act

    def forward(self, x):
        norm_layer = self._modules.get("norm_layer")
        act = self._act_name()
        if self.kernel_size == 1:
            if norm_layer is None and act == "tanh":
                return ops.pred_head(x, self.conv2d.weight, self.conv2d.bias)  # the prediction layers: 1x1 conv + bias + tanh in one kernel
            # other 1x1 layers (normalised predictions, no shipped config): the 1x1 weight as the centre tap of a 3x3 kernel
            w3 = torch.nn.functional.pad(self.conv2d.weight, (1, 1, 1, 1))
            out = ops.conv_ann(x, w3, self.conv2d.bias, None if norm_layer is not None else act)
            if norm_layer is None:
                return out
            out = norm_layer(out)
            return out if act is None else _TORCH_ACT[act](out)
        return _conv_norm_act(x, self.conv2d, norm_layer, act, stride=self.stride)  # (stride 2: the encoders of the ANN U-Nets)


class TransposedConvLayer(nn.Module):
    """
    Transposed convolutional layer (x2 resolution) of the ANN decoders with `use_upsample_conv=False` (models/submodules.py:86-137):
    ConvTranspose2d(kernel 3, stride 2, padding 1, output_padding 1) [+ BN / IN] + activation.  Computed as a stride-1 convolution
    (ef_conv_ann_fwd / ef_conv3x3_bwd) over the zero-inserted input with the flipped, transposed kernel:
    out[o] = sum_i x[i] w[o + 1 - 2 i]  =  sum_j z[j] w'[j - o + 1],  z[2 i] = x[i],  w'[co, ci, ky, kx] = w[ci, co, 2 - ky, 2 - kx].
    """

    def __init__(self, in_channels, out_channels, kernel_size, activation="relu", norm=None):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("event_flow_b200 TransposedConvLayer: kernel_size 3 only")
        if activation not in (None, "relu", "sigmoid", "tanh"):
            raise NotImplementedError(f"event_flow_b200 TransposedConvLayer: activation={activation!r} not on the CUDA path")
        self.transposed_conv2d = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride=2, padding=kernel_size // 2,
                                                    output_padding=1, bias=norm != "BN")
        self.activation = activation
        self.norm = norm
        layer = _make_norm(norm, out_channels)
        if layer is not None:
            self.norm_layer = layer

    def forward(self, x):
        B, C, H, W = x.shape
        z = x.new_zeros((B, C, 2 * H, 2 * W))
        z[:, :, ::2, ::2] = x
        conv = self.transposed_conv2d
        w = conv.weight.flip(2, 3).transpose(0, 1).contiguous()  # [Cout, Cin, 3, 3] of the equivalent correlation
        act = getattr(self.activation, "__name__", self.activation) if callable(self.activation) else self.activation
        norm_layer = self._modules.get("norm_layer")
        if norm_layer is None:
            return ops.conv_ann(z, w, conv.bias, act)
        out = norm_layer(ops.conv_ann(z, w, conv.bias, None))
        return out if act is None else _TORCH_ACT[act](out)


class ConvLayer_(ConvLayer):
    """Clone of ConvLayer that acts like it has state, and allows residual (models/submodules.py:64-83)."""

    def forward(self, x, prev_state, residual=0):
        if prev_state is None:
            prev_state = torch.tensor(0)  # not used
        res = residual if torch.is_tensor(residual) else None
        out = ops.conv_ann(x, self.conv2d.weight, self.conv2d.bias, self._act_name(), residual=res)
        return out, prev_state


class RecurrentConvLayer(nn.Module):
    """Convolution followed by a recurrent block (models/submodules.py:188-235): ConvLSTM, ConvGRU or ConvRecurrent."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, recurrent_block_type="convlstm", activation_ff="relu",
                 activation_rec=None, norm=None, BN_momentum=0.1):
        super().__init__()
        assert recurrent_block_type in ["convlstm", "convgru", "convrnn"]
        self.recurrent_block_type = recurrent_block_type
        block = {"convlstm": ConvLSTM, "convgru": ConvGRU, "convrnn": ConvRecurrent}[recurrent_block_type]
        self.conv = ConvLayer(in_channels, out_channels, kernel_size, stride, activation_ff, norm, BN_momentum=BN_momentum)
        self.recurrent_block = block(input_size=out_channels, hidden_size=out_channels, kernel_size=3, activation=activation_rec)

    def forward(self, x, prev_state):
        x = self.conv(x)
        x, state = self.recurrent_block(x, prev_state)
        if isinstance(self.recurrent_block, ConvLSTM):
            state = (x, state)
        return x, state


class ResidualBlock(nn.Module):
    """He et al. residual block (models/submodules.py:238-312): conv [+ norm] + act, conv [+ norm] + residual + act.  Two launches
    without normalisation."""

    def __init__(self, in_channels, out_channels, stride=1, activation="relu", downsample=None, norm=None, BN_momentum=0.1):
        super().__init__()
        if stride != 1:
            raise NotImplementedError("event_flow_b200 ResidualBlock: stride 1 only (no model of the reference builds another)")
        if activation not in (None, "relu", "sigmoid", "tanh"):
            raise NotImplementedError(f"event_flow_b200 ResidualBlock: activation={activation!r} not on the CUDA path")
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=norm != "BN")
        self.activation = activation
        self.norm = norm
        if norm in ("BN", "IN"):
            self.bn1 = _make_norm(norm, out_channels, BN_momentum)
            self.bn2 = _make_norm(norm, out_channels, BN_momentum)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=norm != "BN")
        self.downsample = downsample

    def forward(self, x):
        act = getattr(self.activation, "__name__", self.activation) if callable(self.activation) else self.activation
        residual = self.downsample(x) if self.downsample else x
        out1 = _conv_norm_act(x, self.conv1, self._modules.get("bn1"), act)
        out2 = _conv_norm_act(out1, self.conv2, self._modules.get("bn2"), act, residual=residual)
        return out2, out1


class UpsampleConvLayer(nn.Module):
    """Bilinear x2 upsampling + conv [+ norm] + activation (models/submodules.py:140-185): the decoder stage of the ANN U-Net."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, activation="relu", norm=None):
        super().__init__()
        if stride != 1 or kernel_size != 3:
            raise NotImplementedError("event_flow_b200 UpsampleConvLayer: kernel_size 3, stride 1")
        if activation not in (None, "relu", "sigmoid", "tanh"):
            raise NotImplementedError(f"event_flow_b200 UpsampleConvLayer: activation={activation!r} not on the CUDA path")
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride, kernel_size // 2, bias=norm != "BN")
        self.activation = activation
        self.norm = norm
        layer = _make_norm(norm, out_channels)
        if layer is not None:
            self.norm_layer = layer

    def forward(self, x):
        act = getattr(self.activation, "__name__", self.activation) if callable(self.activation) else self.activation
        return _conv_norm_act(ops.upsample_bilinear2x(x), self.conv2d, self._modules.get("norm_layer"), act)


class ConvGRU(nn.Module):
    """Convolutional GRU cell (models/submodules.py:377-418): two fused launches instead of 3 convs + 2 cats + 8 pointwise ops."""

    def __init__(self, input_size, hidden_size, kernel_size, activation=None):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("event_flow_b200 ConvGRU: kernel_size 3 only")
        padding = kernel_size // 2
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.reset_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=padding)
        self.update_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=padding)
        self.out_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=padding)
        assert activation is None, "ConvGRU activation cannot be set (just for compatibility)"
        nn.init.orthogonal_(self.reset_gate.weight)
        nn.init.orthogonal_(self.update_gate.weight)
        nn.init.orthogonal_(self.out_gate.weight)
        nn.init.constant_(self.reset_gate.bias, 0.0)
        nn.init.constant_(self.update_gate.bias, 0.0)
        nn.init.constant_(self.out_gate.bias, 0.0)

    def forward(self, input_, prev_state):
        if prev_state is None:
            prev_state = torch.zeros([input_.shape[0], self.hidden_size] + list(input_.shape[2:]), dtype=input_.dtype, device=input_.device)
        C = self.hidden_size
        w_ur = torch.cat([self.update_gate.weight, self.reset_gate.weight], dim=0)
        b_ur = torch.cat([self.update_gate.bias, self.reset_gate.bias], dim=0)
        ur = ops.conv_ann(input_, w_ur, b_ur, "sigmoid", x2=prev_state)  # [B, 2C, H, W]: update, reset
        new_state = ops.conv_ann(input_, self.out_gate.weight, self.out_gate.bias, "tanh", x2=prev_state, x2_scale=ur[:, C:],
                                 blend_h=prev_state, blend_u=ur[:, :C])
        return new_state, new_state


# ---- the rest of the cell zoo: convolutions on ef_conv_ann_fwd / ef_conv3x3_bwd, the leak / gate arithmetic as
# ---- elementwise tensor ops around them (first version: these models carry little benchmark weight) --------------------------
_ACTS = {None: None, "relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}


def _zeros_like_out(x, channels, stride=1):
    h, w = (x.shape[2] - 1) // stride + 1, (x.shape[3] - 1) // stride + 1
    return torch.zeros((x.shape[0], channels, h, w), dtype=x.dtype, device=x.device)


class ConvLSTM(nn.Module):
    """Convolutional LSTM cell (models/submodules.py:314-374): one conv launch for the four gates of cat([x, h])."""

    def __init__(self, input_size, hidden_size, kernel_size, activation=None):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("event_flow_b200 ConvLSTM: kernel_size 3 only")
        self.input_size = input_size
        self.hidden_size = hidden_size
        assert activation is None, "ConvLSTM activation cannot be set (just for compatibility)"
        self.zero_tensors = {}
        self.Gates = nn.Conv2d(input_size + hidden_size, 4 * hidden_size, kernel_size, padding=kernel_size // 2)

    def forward(self, input_, prev_state=None):
        if prev_state is None:
            z = _zeros_like_out(input_, self.hidden_size)
            prev_state = (z, z.clone())
        prev_hidden, prev_cell = prev_state
        gates = ops.conv_ann(input_, self.Gates.weight, self.Gates.bias, None, x2=prev_hidden)
        in_gate, remember_gate, out_gate, cell_gate = gates.chunk(4, 1)
        in_gate, remember_gate, out_gate = torch.sigmoid(in_gate), torch.sigmoid(remember_gate), torch.sigmoid(out_gate)
        cell_gate = torch.tanh(cell_gate)
        cell = (remember_gate * prev_cell) + (in_gate * cell_gate)
        hidden = out_gate * torch.tanh(cell)
        return hidden, cell


class ConvRecurrent(nn.Module):
    """Convolutional recurrent cell (models/submodules.py:421-451): tanh(ff(x) + rec(h)) as ONE conv over cat([x, h]), then out conv + ReLU."""

    def __init__(self, input_size, hidden_size, kernel_size, activation=None):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("event_flow_b200 ConvRecurrent: kernel_size 3 only")
        padding = kernel_size // 2
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.ff = nn.Conv2d(input_size, hidden_size, kernel_size, padding=padding)
        self.rec = nn.Conv2d(input_size, hidden_size, kernel_size, padding=padding)
        self.out = nn.Conv2d(input_size, hidden_size, kernel_size, padding=padding)
        assert activation is None, "ConvRecurrent activation cannot be set (just for compatibility)"

    def forward(self, input_, prev_state):
        if prev_state is None:
            prev_state = _zeros_like_out(input_, self.hidden_size)
        w = torch.cat([self.ff.weight, self.rec.weight], dim=1)
        state = ops.conv_ann(input_, w, self.ff.bias + self.rec.bias, "tanh", x2=prev_state)
        out = ops.conv_ann(state, self.out.weight, self.out.bias, "relu")
        return out, state


class ConvLeakyRecurrent(nn.Module):
    """Leaky convolutional recurrent cell (models/submodules.py:454-499)."""

    def __init__(self, input_size, hidden_size, kernel_size, activation=None, leak=(-4.0, 0.1), learn_leak=True, norm=None):
        super().__init__()
        if kernel_size != 3 or norm is not None:
            raise NotImplementedError("event_flow_b200 ConvLeakyRecurrent: kernel_size 3, no norm")
        padding = kernel_size // 2
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.ff = nn.Conv2d(input_size, hidden_size, kernel_size, padding=padding)
        self.rec = nn.Conv2d(input_size, hidden_size, kernel_size, padding=padding)
        self.out = nn.Conv2d(input_size, hidden_size, kernel_size, padding=padding)
        if learn_leak:
            self.leak = nn.Parameter(torch.randn(hidden_size, 1, 1) * leak[1] + leak[0])
        else:
            self.register_buffer("leak", torch.randn(hidden_size, 1, 1) * leak[1] + leak[0])
        assert activation is None, "ConvLeakyRecurrent activation cannot be set (just for compatibility)"

    def forward(self, input_, prev_state):
        if prev_state is None:
            prev_state = _zeros_like_out(input_, self.hidden_size)
        w = torch.cat([self.ff.weight, self.rec.weight], dim=1)
        pre = ops.conv_ann(input_, w, self.ff.bias + self.rec.bias, None, x2=prev_state)  # ff(x) + rec(h)
        leak = torch.sigmoid(self.leak)
        state = torch.tanh(prev_state * leak + (1 - leak) * pre)
        out = ops.conv_ann(state, self.out.weight, self.out.bias, "relu")
        return out, state


class ConvLeaky(nn.Module):
    """Leaky stateful convolutional cell (models/submodules.py:502-554)."""

    def __init__(self, input_size, hidden_size, kernel_size, stride=1, activation="relu", leak=(-4.0, 0.1), learn_leak=True, norm=None):
        super().__init__()
        if kernel_size != 3 or stride not in (1, 2) or norm is not None or activation not in _ACTS:
            raise NotImplementedError(f"event_flow_b200 ConvLeaky: kernel_size 3, stride 1|2, relu/tanh/sigmoid/None (got {kernel_size}, {stride}, {activation!r})")
        padding = kernel_size // 2
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.stride = stride
        self.ff = nn.Conv2d(input_size, hidden_size, kernel_size, stride=stride, padding=padding)
        if learn_leak:
            self.leak = nn.Parameter(torch.randn(hidden_size, 1, 1) * leak[1] + leak[0])
        else:
            self.register_buffer("leak", torch.randn(hidden_size, 1, 1) * leak[1] + leak[0])
        self.activation = _ACTS[activation]

    def forward(self, input_, prev_state, residual=0):
        ff = ops.conv_ann(input_, self.ff.weight, self.ff.bias, None, stride=self.stride)
        if prev_state is None:
            prev_state = torch.zeros_like(ff)
        leak = torch.sigmoid(self.leak)
        state = prev_state * leak + (1 - leak) * (ff + residual)
        out = self.activation(state) if self.activation is not None else state
        return out, state


class LeakyResidualBlock(nn.Module):
    """models/submodules.py:557-592."""

    def __init__(self, in_channels, out_channels, stride=1, feedforward_block_type="convleaky", activation="relu", **kwargs):
        super().__init__()
        assert feedforward_block_type in ["convleaky"]
        self.conv1 = ConvLeaky(in_channels, out_channels, kernel_size=3, stride=stride, activation=activation, **kwargs)
        self.conv2 = ConvLeaky(out_channels, out_channels, kernel_size=3, stride=1, activation=activation, **kwargs)

    def forward(self, x, prev_state):
        if prev_state is None:
            prev_state = [None, None]
        conv1, conv2 = prev_state
        residual = x
        x1, conv1 = self.conv1(x, conv1)
        x2, conv2 = self.conv2(x1, conv2, residual=residual)
        return x2, torch.stack([conv1, conv2])


class LeakyUpsampleConvLayer(nn.Module):
    """models/submodules.py:595-623."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, feedforward_block_type="convleaky", activation="relu", **kwargs):
        super().__init__()
        assert feedforward_block_type in ["convleaky"]
        self.conv2d = ConvLeaky(in_channels, out_channels, kernel_size, stride=stride, activation=activation, **kwargs)

    def forward(self, x, prev_state):
        return self.conv2d(ops.upsample_bilinear2x(x), prev_state)


class LeakyTransposedConvLayer(nn.Module):
    """models/submodules.py:626-641 (raises in the reference as well)."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError


class LeakyRecurrentConvLayer(nn.Module):
    """models/submodules.py:644-686."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=2, recurrent_block_type="convleaky", activation_ff="relu",
                 activation_rec=None, **kwargs):
        super().__init__()
        assert recurrent_block_type in ["convleaky"]
        self.conv = ConvLeaky(in_channels, out_channels, kernel_size, stride, activation_ff, **kwargs)
        self.recurrent_block = ConvLeakyRecurrent(out_channels, out_channels, kernel_size, activation=activation_rec, **kwargs)

    def forward(self, x, prev_state):
        if prev_state is None:
            prev_state = [None, None]
        ff, rec = prev_state
        x1, ff = self.conv(x, ff)
        x2, rec = self.recurrent_block(x1, rec)
        return x2, torch.stack([ff, rec])
