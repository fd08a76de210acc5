"""
ANN layers of models/submodules.py.  Round 1 covers what the spiking FireNet family needs: ConvLayer as the 1x1 tanh
prediction head (models/submodules.py:12-61 as built at models/model.py:197-199).  Other configurations raise.
"""
import torch.nn as nn

from .. import ops


class ConvLayer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, activation="relu", norm=None, BN_momentum=0.1, w_scale=None):
        super().__init__()
        if kernel_size != 1 or stride != 1 or activation != "tanh" or norm is not None:
            raise NotImplementedError(
                "event_flow_b200 ConvLayer: only the 1x1 tanh prediction head is on the CUDA path in this version "
                f"(got kernel_size={kernel_size}, stride={stride}, activation={activation}, norm={norm})"
            )
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride, kernel_size // 2, bias=True)
        if w_scale is not None:
            nn.init.uniform_(self.conv2d.weight, -w_scale, w_scale)
            nn.init.zeros_(self.conv2d.bias)
        self.norm = norm

    def forward(self, x):
        return ops.pred_head(x, self.conv2d.weight, self.conv2d.bias)
