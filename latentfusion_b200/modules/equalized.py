"""Equalised-learning-rate convolutions backed by the fused lfb200 conv kernel.

API + checkpoint mirror of reference ``latentfusion/modules/equalized.py`` (Equalized :39-75,
EqualizedConv2d/3d :77-92): weights are stored N(0,1) under ``<name>.module.weight`` with a separate
``<name>.bias``; the He constant sqrt(2/fan_in) is applied at run time.  Unlike the reference, which
runs conv -> mul_ -> add as three passes (and the Block's LeakyReLU/PixelNorm as four more), a layer
here can be asked to fuse its trailing activation + PixelNorm into the same kernel
(``forward(x, act=..., norm=...)``); called plainly it is exactly ``conv(x)*he + bias``.
"""
import math

import torch
from torch import nn

from .. import ops


class _ConvParams(nn.Module):
    """Holds ``weight`` so that state_dict keys read ``<layer>.module.weight`` like nn.ConvNd did."""

    def __init__(self, in_channels, out_channels, kernel_size, ndim, padding):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = (kernel_size,) * ndim
        self.padding = (padding,) * ndim
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *self.kernel_size))
        self.bias = None

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, padding={self.padding}"


class Equalized(nn.Module):
    ndim = 0

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding: int = 0,
                 padding_mode='zeros', equalized=True, lr_scale=1.0, bias=True):
        super().__init__()
        if stride != 1 or padding_mode != 'zeros':
            raise ValueError("lfb200 convolutions support stride=1 and zero padding only "
                             "(all the reconstruct->render path uses)")
        if kernel_size not in (1, 3) or padding != kernel_size // 2:
            raise ValueError(f"lfb200 convolutions support k=1/pad=0 and k=3/pad=1, got k={kernel_size} pad={padding}")
        if not equalized:
            raise ValueError("only the equalised form is used on the path")
        if lr_scale != 1.0:
            # reference equalized.py:44-47 rescales the stored weights by 1/lr_scale and the run-time constant by
            # lr_scale; nothing on the path passes it, and the fused kernels recompute sqrt(2/fan_in) themselves
            raise ValueError("lfb200 convolutions support lr_scale=1.0 only")
        self.module = _ConvParams(in_channels, out_channels, kernel_size, self.ndim, padding)
        self.equalized = equalized
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter('bias', None)
        self.module.weight.data.normal_(0, 1)
        self.module.weight.data /= lr_scale
        self.weight = self.get_he_constant() * lr_scale      # python float, not persisted (as in the reference)
        self.precision = None                                  # None -> ops default

    def get_he_constant(self):
        return math.sqrt(2.0 / math.prod(self.module.weight.shape[1:]))

    def forward(self, x, act=False, slope=0.2, norm=False, emit_split=False):
        return ops.eq_conv(x, self.module.weight, self.bias, act=act, slope=slope, norm=norm,
                           precision=self.precision, emit_split=emit_split)


class EqualizedConv2d(Equalized):
    ndim = 2


class EqualizedConv3d(Equalized):
    ndim = 3
