"""
ANN layers of models/submodules.py used by the FireNet family: ConvLayer (:12-61), ConvLayer_ (:64-83), ConvGRU (:377-418),
with the reference's constructor signatures, parameter names and initialisers.  Forward passes are CUDA kernels
(ef_conv_ann_fwd / ef_pred_fwd); under autograd the convolution gradients come from ef_conv3x3_bwd (ops._ConvAnn).
"""
import torch
import torch.nn as nn

from .. import ops


class ConvLayer(nn.Module):
    """Convolutional layer: conv + bias + activation (default ReLU), no downsampling, no batch norm."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, activation="relu", norm=None, BN_momentum=0.1, w_scale=None):
        super().__init__()
        if norm is not None:
            raise NotImplementedError("event_flow_b200 ConvLayer: norm=%r is not on the CUDA path (no shipped FireNet config uses it)" % (norm,))
        if kernel_size not in (1, 3) or stride not in (1, 2) or (kernel_size == 1 and stride != 1):
            raise NotImplementedError(f"event_flow_b200 ConvLayer: kernel_size={kernel_size}, stride={stride} not on the CUDA path yet")
        if kernel_size == 1 and activation != "tanh":
            raise NotImplementedError("event_flow_b200 ConvLayer: the 1x1 layer is built for the tanh prediction head only")
        if kernel_size == 3 and activation not in (None, "relu", "sigmoid", "tanh"):
            raise NotImplementedError(f"event_flow_b200 ConvLayer: activation={activation!r} not on the CUDA path")
        padding = kernel_size // 2
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=True)
        if w_scale is not None:
            nn.init.uniform_(self.conv2d.weight, -w_scale, w_scale)
            nn.init.zeros_(self.conv2d.bias)
        self.kernel_size = kernel_size
        self.stride = stride
        self.activation = activation
        self.norm = norm

    def forward(self, x):
        if self.kernel_size == 1:
            return ops.pred_head(x, self.conv2d.weight, self.conv2d.bias)
        out = ops.conv_ann(x, self.conv2d.weight, self.conv2d.bias, self.activation)
        if self.stride == 2:
            # a stride-2 3x3 conv with padding 1 is the stride-1 result at the even pixels (same window, same summation order);
            # first version of the U-Net encoders: 4x the minimal FLOPs on these four layers
            out = out[:, :, ::2, ::2].contiguous()
        return out


class ConvLayer_(ConvLayer):
    """Clone of ConvLayer that acts like it has state, and allows residual (models/submodules.py:64-83)."""

    def forward(self, x, prev_state, residual=0):
        if prev_state is None:
            prev_state = torch.tensor(0)  # not used
        res = residual if torch.is_tensor(residual) else None
        out = ops.conv_ann(x, self.conv2d.weight, self.conv2d.bias, self.activation, residual=res)
        return out, prev_state


class RecurrentConvLayer(nn.Module):
    """Convolution followed by a recurrent block (models/submodules.py:188-235); ConvGRU blocks only on the CUDA path."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, recurrent_block_type="convlstm", activation_ff="relu",
                 activation_rec=None, norm=None, BN_momentum=0.1):
        super().__init__()
        assert recurrent_block_type in ["convlstm", "convgru", "convrnn"]
        if recurrent_block_type != "convgru":
            raise NotImplementedError(f"event_flow_b200 RecurrentConvLayer: recurrent_block_type={recurrent_block_type!r} is not on the CUDA path "
                                      "(ConvGRU is)")
        self.recurrent_block_type = recurrent_block_type
        self.conv = ConvLayer(in_channels, out_channels, kernel_size, stride, activation_ff, norm, BN_momentum=BN_momentum)
        self.recurrent_block = ConvGRU(input_size=out_channels, hidden_size=out_channels, kernel_size=3, activation=activation_rec)

    def forward(self, x, prev_state):
        x = self.conv(x)
        x, state = self.recurrent_block(x, prev_state)
        return x, state


class ResidualBlock(nn.Module):
    """He et al. residual block (models/submodules.py:238-312): conv+act, conv + residual + act.  Two launches."""

    def __init__(self, in_channels, out_channels, stride=1, activation="relu", downsample=None, norm=None, BN_momentum=0.1):
        super().__init__()
        if norm is not None or downsample is not None or stride != 1 or in_channels != out_channels:
            raise NotImplementedError("event_flow_b200 ResidualBlock: norm / downsample / stride are not on the CUDA path (no shipped config uses them)")
        if activation not in (None, "relu", "sigmoid", "tanh"):
            raise NotImplementedError(f"event_flow_b200 ResidualBlock: activation={activation!r} not on the CUDA path")
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=True)
        self.activation = activation
        self.norm = norm
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=True)
        self.downsample = downsample

    def forward(self, x):
        out1 = ops.conv_ann(x, self.conv1.weight, self.conv1.bias, self.activation)
        out2 = ops.conv_ann(out1, self.conv2.weight, self.conv2.bias, self.activation, residual=x)
        return out2, out1


class UpsampleConvLayer(nn.Module):
    """Bilinear x2 upsampling + conv + activation (models/submodules.py:140-185): the decoder stage of the ANN U-Net."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, activation="relu", norm=None):
        super().__init__()
        if norm is not None or stride != 1 or kernel_size != 3:
            raise NotImplementedError("event_flow_b200 UpsampleConvLayer: kernel_size 3, stride 1, no norm")
        if activation not in (None, "relu", "sigmoid", "tanh"):
            raise NotImplementedError(f"event_flow_b200 UpsampleConvLayer: activation={activation!r} not on the CUDA path")
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride, kernel_size // 2, bias=True)
        self.activation = activation
        self.norm = norm

    def forward(self, x):
        return ops.conv_ann(ops.upsample_bilinear2x(x), self.conv2d.weight, self.conv2d.bias, self.activation)


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
