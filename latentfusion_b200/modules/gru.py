"""Convolutional GRU cell over latent volumes.  API/checkpoint mirror of reference
``latentfusion/modules/gru.py`` (ConvGRUCell :7-46); gate math runs in two fused kernels.

The reference concatenates ``[x, h]`` (2C+3 channels at the GRU fuser) and convolves once per gate.  A
convolution is linear in its input channels, so here each gate is evaluated as ``conv_x(x) + conv_h(h)`` over
the two channel groups (x zero-padded to a multiple of 4 channels): both halves then have shapes the tcgen05
kernel takes (Cin % 4 == 0, packed weights resident in shared memory) and the ``torch.cat`` copy disappears.
``precision=0`` keeps the single concatenated exact-fp32 convolution.

The GRU fuser's input is ``[cube_i, voxel coordinates]``: the coordinate channels are the same constant field at
every step, so their contribution ``conv(coords, W[:, C:C+3])`` is evaluated ONCE per forward per gate
(``extra_terms``) and added at each step; the per-step convolutions then see exactly C input channels (no 35 -> 48
channel padding, and the tensor-core weight-gradient kernel applies: Cin_pad in {16, 32}).
"""
import math

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
        self._parts_cache = {}

    # -- channel-group evaluation -----------------------------------------------------------------
    def _weight_parts(self, name, gate, cx, cxp):
        """(weight[:, :cx] zero-padded to cxp input channels, weight[:, cx:]) — cached while the weights are frozen."""
        w = gate.module.weight
        track = torch.is_grad_enabled() and w.requires_grad
        key = (w._version, w.data_ptr(), cxp)
        if not track:
            hit = self._parts_cache.get(name)
            if hit is not None and hit[0] == key:
                return hit[1], hit[2]
        wx = w[:, :cx]
        if cxp != cx:
            wx = torch.cat([wx, wx.new_zeros(w.shape[0], cxp - cx, *w.shape[2:])], dim=1)
        wx, wh = wx.contiguous(), w[:, cx:].contiguous()
        if not track:
            wx, wh = wx.detach(), wh.detach()
            self._parts_cache[name] = (key, wx, wh)
        return wx, wh

    def _gate(self, name, gate, xp, cx, h):
        wx, wh = self._weight_parts(name, gate, cx, xp.shape[1])
        fan_in = int(math.prod(gate.module.weight.shape[1:]))
        prec = gate.precision
        return (ops.eq_conv(xp, wx, gate.bias, precision=prec, fan_in=fan_in)
                + ops.eq_conv(h, wh, None, precision=prec, fan_in=fan_in))

    # -- constant extra input channels (the fuser's coordinate field) ------------------------------
    def splits_inputs(self, like):
        """True when forward() runs the channel-group evaluation for tensors like `like` (tensor-core precisions)"""
        prec = self.update_gate.precision
        return like.is_cuda and (ops.get_default_precision() if prec is None else prec) != ops.PRECISION_FP32

    def extra_terms(self, extra, main_channels):
        """per-gate conv(extra, W[:, main:main+e]) (no bias) for an input laid out as [main | extra | hidden]"""
        e = extra.shape[1]
        ep = (e + 3) // 4 * 4
        xe = ops.empty_cl((extra.shape[0], ep, *extra.shape[2:]), extra.device)
        xe[:, :e] = extra
        xe[:, e:] = 0
        out = []
        for gate in (self.update_gate, self.reset_gate, self.out_gate):
            w = gate.module.weight[:, main_channels:main_channels + e]
            if ep != e:
                w = torch.cat([w, w.new_zeros(w.shape[0], ep - e, *w.shape[2:])], dim=1)
            fan_in = int(math.prod(gate.module.weight.shape[1:]))
            out.append(ops.eq_conv(xe, w.contiguous(), None, precision=gate.precision, fan_in=fan_in))
        return tuple(out)

    def _gate3(self, gate, x, h, term):
        """conv over [x | extra | h] with the extra part precomputed: x and h both have hidden-size channel groups"""
        fan_in = int(math.prod(gate.module.weight.shape[1:]))
        wx, wh = self._split_weights(gate, x.shape[1], h.shape[1])
        return (ops.eq_conv(x, wx, gate.bias, precision=gate.precision, fan_in=fan_in)
                + ops.eq_conv(h, wh, None, precision=gate.precision, fan_in=fan_in) + term)

    def _split_weights(self, gate, cx, ch):
        """(W[:, :cx], W[:, -ch:]) contiguous; cached while the weights are frozen (keeps their packed-weight entries)"""
        w = gate.module.weight
        track = torch.is_grad_enabled() and w.requires_grad
        key, hit = (w._version, w.data_ptr(), cx, ch), self._parts_cache.get(('3', id(gate)))
        if not track and hit is not None and hit[0] == key:
            return hit[1], hit[2]
        wx, wh = w[:, :cx].contiguous(), w[:, w.shape[1] - ch:].contiguous()
        if not track:
            wx, wh = wx.detach(), wh.detach()
            self._parts_cache[('3', id(gate))] = (key, wx, wh)
        return wx, wh

    def input_terms(self, x_all, extra_terms, views):
        """The input half of every step's gate pre-activations in ONE convolution per gate: the steps' inputs do not
        depend on the recurrence, so conv(x_i, W[:, :C]) + bias + (constant extra term) is evaluated for all views at
        once (a [B*views]-image batch fills the machine; the per-step convolutions see only the hidden state).
        x_all [B, views, C, ...] -> three tensors [B, views, hidden, ...] (update, reset, out)."""
        b, cx = x_all.shape[0], x_all.shape[2]
        flat = x_all.reshape(b * views, *x_all.shape[2:])
        out = []
        for gate, term in zip((self.update_gate, self.reset_gate, self.out_gate), extra_terms):
            wx, _ = self._split_weights(gate, cx, self.hidden_dim)
            fan_in = int(math.prod(gate.module.weight.shape[1:]))
            pre = ops.eq_conv(flat, wx, gate.bias, precision=gate.precision, fan_in=fan_in)
            out.append(pre.view(b, views, *pre.shape[1:]) + term.unsqueeze(1))
        return tuple(out)

    def step_hidden(self, h_cur, pre_u, pre_r, pre_o, main_channels):
        """one recurrence step given the input halves of the three gates: only conv(h, W[:, -hidden:]) remains"""
        def hconv(gate, h):
            _, wh = self._split_weights(gate, main_channels, self.hidden_dim)
            return ops.eq_conv(h, wh, None, precision=gate.precision, fan_in=int(math.prod(gate.module.weight.shape[1:])))
        update, h_reset = ops.gru_gates1(pre_u + hconv(self.update_gate, h_cur), pre_r + hconv(self.reset_gate, h_cur), h_cur)
        return ops.gru_gates2(h_cur, update, pre_o + hconv(self.out_gate, h_reset))

    def forward(self, x, h_cur, extra=None):
        if extra is not None:           # (x: the main channels only; extra = extra_terms(...) of the remaining inputs)
            update, h_reset = ops.gru_gates1(self._gate3(self.update_gate, x, h_cur, extra[0]),
                                             self._gate3(self.reset_gate, x, h_cur, extra[1]), h_cur)
            return ops.gru_gates2(h_cur, update, self._gate3(self.out_gate, x, h_reset, extra[2]))
        prec = self.update_gate.precision
        if (ops.get_default_precision() if prec is None else prec) == ops.PRECISION_FP32 or not x.is_cuda:
            x_in = torch.cat([x, h_cur], dim=1)
            update, h_reset = ops.gru_gates1(self.update_gate(x_in), self.reset_gate(x_in), h_cur)
            x_out = self.out_gate(torch.cat([x, h_reset], dim=1))
            return ops.gru_gates2(h_cur, update, x_out)
        cx = x.shape[1]
        cxp = (cx + 3) // 4 * 4
        if cxp != cx:
            xp = ops.empty_cl((x.shape[0], cxp, *x.shape[2:]), x.device)
            xp[:, :cx] = x
            xp[:, cx:] = 0
        else:
            xp = x
        update, h_reset = ops.gru_gates1(self._gate('u', self.update_gate, xp, cx, h_cur),
                                         self._gate('r', self.reset_gate, xp, cx, h_cur), h_cur)
        x_out = self._gate('o', self.out_gate, xp, cx, h_reset)
        return ops.gru_gates2(h_cur, update, x_out)
