"""
ANN layers of models/submodules.py used by the FireNet family: ConvLayer (:12-61), ConvLayer_ (:64-83), ConvGRU (:377-418),
with the reference's constructor signatures, parameter names and initialisers.  Forward passes are CUDA kernels
(ef_conv_ann_fwd / ef_pred_fwd); the backward of the 3x3 ANN cells is not built in this version (it raises).
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
        if stride != 1 or kernel_size not in (1, 3):
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
        self.activation = activation
        self.norm = norm

    def forward(self, x):
        if self.kernel_size == 1:
            return ops.pred_head(x, self.conv2d.weight, self.conv2d.bias)
        return ops.conv_ann(x, self.conv2d.weight, self.conv2d.bias, self.activation)


class ConvLayer_(ConvLayer):
    """Clone of ConvLayer that acts like it has state, and allows residual (models/submodules.py:64-83)."""

    def forward(self, x, prev_state, residual=0):
        if prev_state is None:
            prev_state = torch.tensor(0)  # not used
        res = residual if torch.is_tensor(residual) else None
        out = ops.conv_ann(x, self.conv2d.weight, self.conv2d.bias, self.activation, residual=res)
        return out, prev_state


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
