"""U-Net assembled from a token list (``block_config = [encoder_tokens, decoder_tokens]``; integers are channel
widths, ``D``/``U`` halve/double the resolution after the block, see ``blocks.create_blocks``).

Constructor arguments, attribute names (``input_block``, ``down_blocks``, ``up_blocks``, ``output_block``) and
therefore ``state_dict`` keys are those of the reference's ``latentfusion/modules/unet.py:8-127``, so its checkpoints
load unchanged.  Behaviour kept from there: the resizes inside a U-Net are always (bi|tri)linear (no ``scale_mode`` is
forwarded), decoder block 0 gets no skip connection, and ``out_channels`` may be ``None`` (no head), an int (one
head) or a tuple (several heads concatenated along channels).  The convolutions are the fused lfb200 layers.
"""
import torch
from torch import nn

from . import EqualizedConv2d, EqualizedConv3d
from .blocks import InputBlock, OutputBlock, count_blocks, create_blocks

_HALVING, _DOUBLING = ('I', 'D'), ('I', 'U')


def _width(spec, fallback):
    """channel count described by a ctor argument: None -> fallback, int -> itself, sequence -> its sum"""
    if spec is None:
        return fallback
    return spec if isinstance(spec, int) else sum(spec)


class BaseUNet(nn.Module):

    def __init__(self, in_channels, out_channels, block_config, conv_module):
        super().__init__()
        self._in_channels, self._out_channels, self._conv_module = in_channels, out_channels, conv_module
        self.block_config = block_config
        enc, dec = block_config
        self.input_block = None if in_channels is None else InputBlock(in_channels, enc[0], conv_module=conv_module)
        self.down_blocks = create_blocks(enc, conv_module, 0.5)
        self.up_blocks = create_blocks(dec, conv_module, 2.0, skip_connections=True,
                                       skip_connect_end=min(count_blocks(enc), count_blocks(dec)))
        self.output_block = self._make_heads(dec[-1], out_channels, conv_module)

    @staticmethod
    def _make_heads(width, out_channels, conv_module):
        if out_channels is None:
            return None
        if isinstance(out_channels, int):
            return OutputBlock(width, out_channels, conv_module=conv_module)
        return nn.ModuleList(OutputBlock(width, c, conv_module=conv_module) for c in out_channels)

    # ---- (de)serialisation in the reference's checkpoint layout
    def create_checkpoint(self):
        args = dict(in_channels=self._in_channels, out_channels=self._out_channels, block_config=self.block_config,
                    conv_module=None)
        return {'args': args, 'state_dict': self.cpu().state_dict()}

    @classmethod
    def from_checkpoint(cls, checkpoint):
        args = {k: v for k, v in checkpoint['args'].items() if k != 'conv_module'}
        net = cls(**args)
        net.load_state_dict(checkpoint['state_dict'])
        return net

    # ---- geometry of the network
    down_block_config = property(lambda self: self.block_config[0])
    up_block_config = property(lambda self: self.block_config[1])
    in_channels = property(lambda self: _width(self._in_channels, self.block_config[0][0]))
    out_channels = property(lambda self: _width(self._out_channels, self.block_config[1][-1]))

    def bottleneck_size(self, in_size):
        return in_size >> sum(self.down_block_config.count(t) for t in _HALVING)

    def output_size(self, in_size):
        return self.bottleneck_size(in_size) << sum(self.up_block_config.count(t) for t in _DOUBLING)

    # ---- forward
    def _encode(self, z):
        if self.input_block is not None:
            z = self.input_block(z)
        pyramid = []                                   # deepest first
        for block in self.down_blocks:
            z = block(z)
            pyramid.insert(0, z)
        return z, pyramid

    @staticmethod
    def _inject(z, code):
        """broadcast a per-sample code vector over the bottleneck and append it as extra channels"""
        if code.size(0) != z.size(0):
            raise ValueError("z_inject: batch size mismatch")
        code = code.reshape(*code.shape, *(1,) * (z.dim() - 2)).expand(-1, -1, *z.shape[2:])
        return torch.cat((z, code), dim=1)

    def _heads(self, z):
        if self.output_block is None:
            return z
        if isinstance(self.output_block, OutputBlock):
            return self.output_block(z)
        return torch.cat([head(z) for head in self.output_block], dim=1)

    def forward(self, z, z_inject=None, return_intermediate=False):
        z, pyramid = self._encode(z)
        if z_inject is not None:
            z = self._inject(z, z_inject)
        for level, block in enumerate(self.up_blocks):
            if 0 < level < len(pyramid):               # level 0 consumes the bottleneck itself
                z = torch.cat((z, pyramid[level]), dim=1)
            z = block(z)
        z = self._heads(z)
        return (z, pyramid) if return_intermediate else z


class UNet2d(BaseUNet):
    def __init__(self, in_channels, out_channels, block_config):
        super().__init__(in_channels, out_channels, block_config, conv_module=EqualizedConv2d)


class UNet3d(BaseUNet):
    def __init__(self, in_channels, out_channels, block_config):
        super().__init__(in_channels, out_channels, block_config, conv_module=EqualizedConv3d)
