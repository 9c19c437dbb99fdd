"""Conv blocks + the block-config DSL.  API/checkpoint mirror of reference
``latentfusion/modules/blocks.py`` (create_blocks :10-75, InputBlock :78-90, OutputBlock :107-119,
Block :136-164).  A Block here is two fused kernels (+ one for the optional resize) instead of the
reference's ~14 elementwise passes."""
import torch
from torch import nn

from . import Interpolate, PixelNorm, EqualizedConv2d, EqualizedConv3d
from .. import ops


def count_blocks(config):
    return sum(1 for tok in config if isinstance(tok, int)) - 1


def create_blocks(config, conv_module, scale_factor, scale_mode='bilinear', kernel_size=3,
                  skip_connections=False, skip_connect_start=1, skip_connect_end=None,
                  in_views=1, skip_connection_views=None) -> nn.ModuleList:
    """Walk a config such as ``[64, 'D', 128, 'U', 64]``: ints open a Block(prev -> int); 'U'/'D'/'I'
    set the resize applied at the END of the next block (x2 / x0.5 / `scale_factor`)."""
    if conv_module is EqualizedConv3d and scale_mode == 'bilinear':
        scale_mode = 'trilinear'
    if skip_connection_views is None:
        skip_connection_views = in_views
    total = count_blocks(config)
    skip_connect_end = total if skip_connect_end is None else min(total, skip_connect_end)

    resize = {'I': scale_factor, 'U': 2.0, 'D': 0.5}
    blocks, made, pending, width = [], 0, 1.0, config[0]
    for tok in config[1:]:
        if isinstance(tok, str) and not tok.isdigit():
            if tok not in resize:
                raise ValueError(f"Unknown block type {tok!r}")
            pending = resize[tok]
            continue
        tok = int(tok)
        extra = width * skip_connection_views if (
            skip_connections and skip_connect_start <= made < skip_connect_end) else 0
        fan = width * in_views if made == 0 else width
        blocks.append(Block(fan + extra, tok, kernel_size=kernel_size, conv_module=conv_module,
                            scale_mode=scale_mode, scale_factor=pending))
        width, made, pending = tok, made + 1, 1.0
    return nn.ModuleList(blocks)


class InputBlock(nn.Module):
    """1x1 (by default) convolution + LeakyReLU in one kernel: lifts raw inputs to the first width of a config."""

    def __init__(self, in_channels, out_channels, conv_module, kernel_size=1, relu_slope=0.2, padding=0):
        super().__init__()
        self.conv = conv_module(in_channels, out_channels, kernel_size, padding=padding)
        self.activation = nn.LeakyReLU(relu_slope)

    def forward(self, x):
        return self.conv(x, act=True, slope=self.activation.negative_slope)


class OutputBlock(nn.Module):
    """Plain convolution head (logits), optionally followed by a caller-supplied activation module."""

    def __init__(self, in_channels, out_channels, conv_module, kernel_size=1, padding=0, activation=None):
        super().__init__()
        self.conv = conv_module(in_channels, out_channels, kernel_size, padding=padding)
        self.activation = activation

    def forward(self, x):
        y = self.conv(x)
        return y if not self.activation else self.activation(y)


def _bind_conv(base, conv_module, name):
    """`base` with its conv_module argument fixed (the reference spells these out as ...2d / ...3d classes)."""

    def __init__(self, in_channels, out_channels, *args, **kwargs):
        base.__init__(self, in_channels, out_channels, conv_module, *args, **kwargs)
    return type(name, (base,), {'__init__': __init__, '__module__': __name__})


InputBlock2d = _bind_conv(InputBlock, EqualizedConv2d, 'InputBlock2d')
InputBlock3d = _bind_conv(InputBlock, EqualizedConv3d, 'InputBlock3d')
OutputBlock2d = _bind_conv(OutputBlock, EqualizedConv2d, 'OutputBlock2d')
OutputBlock3d = _bind_conv(OutputBlock, EqualizedConv3d, 'OutputBlock3d')


class Block(nn.Module):
    """(conv -> LeakyReLU -> PixelNorm) x 2 -> optional resize."""

    def __init__(self, in_channels, out_channels, kernel_size=3, padding=1, relu_slope=0.2,
                 conv_module=EqualizedConv3d, scale_factor=1.0, scale_mode='bilinear'):
        super().__init__()
        self.activation = nn.LeakyReLU(relu_slope)
        self.norm = PixelNorm()
        self.conv1 = conv_module(in_channels, out_channels, kernel_size, padding=padding)
        self.conv2 = conv_module(out_channels, out_channels, kernel_size, padding=padding)
        self.interpolate = None
        if scale_factor != 1.0 and scale_factor is not None:
            self.interpolate = Interpolate(scale_factor, mode=scale_mode)

    def forward(self, x):
        slope = self.activation.negative_slope
        # (3-D: conv1's epilogue also leaves the split-planar form conv2's TMA staging reads)
        x = self.conv1(x, act=True, slope=slope, norm=True, emit_split=isinstance(self.conv1, EqualizedConv3d))
        # conv2 is the only consumer of conv1's output: its bwd-data kernel can apply conv1's activation/norm backward
        # (emit_out_split: the consumer of this block — the Photographer's depth collapse — reads conv2's activation
        #  backward inputs from the split-planar twin in its fused backward kernel)
        x = self.conv2(ops.mark_single_consumer(x), act=True, slope=slope, norm=True,
                       emit_split=(getattr(self, 'emit_out_split', False) and self.interpolate is None
                                   and torch.is_grad_enabled()))
        if self.interpolate is not None:
            x = self.interpolate(x)
        return x
