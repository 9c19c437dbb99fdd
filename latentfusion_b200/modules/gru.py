"""Convolutional GRU cell over latent volumes.  API/checkpoint mirror of reference
``latentfusion/modules/gru.py`` (ConvGRUCell :7-46); gate math runs in two fused kernels."""
import torch
from torch import nn

from . import EqualizedConv3d
from .. import ops


class ConvGRUCell(nn.Module):

    def __init__(self, in_channels, hidden_channels, kernel_size, bias=True, conv_module=EqualizedConv3d):
        super().__init__()
        self.input_dim, self.hidden_dim = in_channels, hidden_channels
        self.kernel_size, self.padding, self.bias = kernel_size, kernel_size // 2, bias
        cin = in_channels + hidden_channels

        def gate():
            return conv_module(in_channels=cin, out_channels=hidden_channels, kernel_size=kernel_size,
                               padding=self.padding, bias=bias)
        self.update_gate, self.reset_gate, self.out_gate = gate(), gate(), gate()

    def forward(self, x, h_cur):
        x_in = torch.cat([x, h_cur], dim=1)
        update, h_reset = ops.gru_gates1(self.update_gate(x_in), self.reset_gate(x_in), h_cur)
        x_out = self.out_gate(torch.cat([x, h_reset], dim=1))
        return ops.gru_gates2(h_cur, update, x_out)
