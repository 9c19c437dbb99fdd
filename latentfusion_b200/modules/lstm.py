"""Convolutional LSTM cell.  API/checkpoint mirror of reference ``latentfusion/modules/lstm.py``
(ConvLSTMCell :7-56).  The 4-gate convolution is the fused lfb200 kernel; the gate non-linearities run in one
elementwise kernel each way (``lf_lstm_gates_fwd`` / ``_bwd``) on CUDA tensors."""
import torch
from torch import nn

from . import EqualizedConv3d
from .. import ops


class ConvLSTMCell(nn.Module):

    def __init__(self, in_channels, hidden_channels, kernel_size, bias=True, conv_module=EqualizedConv3d):
        super().__init__()
        self.in_channels, self.hidden_channels = in_channels, hidden_channels
        self.kernel_size, self.padding, self.bias = kernel_size, kernel_size // 2, bias
        self.conv = conv_module(in_channels=in_channels + hidden_channels, out_channels=4 * hidden_channels,
                                kernel_size=kernel_size, padding=self.padding, bias=bias)

    def forward(self, input_tensor, cur_state):
        h_cur, c_cur = cur_state
        gates = self.conv(torch.cat([input_tensor, h_cur], dim=1))
        if gates.is_cuda:
            return ops.lstm_gates(gates, c_cur)
        gi, gf, go, gg = torch.split(gates, self.hidden_channels, dim=1)
        c_next = torch.sigmoid(gf) * c_cur + torch.sigmoid(gi) * torch.tanh(gg)
        return torch.sigmoid(go) * torch.tanh(c_next), c_next
