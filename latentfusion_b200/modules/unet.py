"""Config-driven U-Net.  API/checkpoint mirror of reference ``latentfusion/modules/unet.py``
(BaseUNet :8-127).  ``block_config`` is ``[down_tokens, up_tokens]``."""
import torch
from torch import nn

from . import EqualizedConv2d, EqualizedConv3d
from .blocks import create_blocks, InputBlock, OutputBlock, count_blocks


class BaseUNet(nn.Module):

    def __init__(self, in_channels, out_channels, block_config, conv_module):
        super().__init__()
        self._in_channels, self._out_channels = in_channels, out_channels
        self.block_config = block_config
        self._conv_module = conv_module
        down, up = block_config
        self.input_block = (InputBlock(in_channels, down[0], conv_module=conv_module)
                            if in_channels is not None else None)
        # NB: scale_mode is not forwarded, so resizes inside a U-Net are always (bi|tri)linear.
        self.down_blocks = create_blocks(down, conv_module, 0.5)
        self.up_blocks = create_blocks(up, conv_module, 2.0, skip_connections=True,
                                       skip_connect_end=min(count_blocks(down), count_blocks(up)))
        if out_channels is None:
            self.output_block = None
        elif isinstance(out_channels, int):
            self.output_block = OutputBlock(up[-1], out_channels, conv_module=conv_module)
        else:
            self.output_block = nn.ModuleList(
                [OutputBlock(up[-1], c, conv_module=conv_module) for c in out_channels])

    @classmethod
    def from_checkpoint(cls, checkpoint):
        checkpoint['args'].pop('conv_module', None)
        model = cls(**checkpoint['args'])
        model.load_state_dict(checkpoint['state_dict'])
        return model

    def create_checkpoint(self):
        return {'args': {'in_channels': self._in_channels, 'out_channels': self._out_channels,
                         'block_config': self.block_config, 'conv_module': None},
                'state_dict': self.cpu().state_dict()}

    @property
    def down_block_config(self):
        return self.block_config[0]

    @property
    def up_block_config(self):
        return self.block_config[1]

    @property
    def in_channels(self):
        if self._in_channels is None:
            return self.down_block_config[0]
        return self._in_channels if isinstance(self._in_channels, int) else sum(self._in_channels)

    @property
    def out_channels(self):
        if self._out_channels is None:
            return self.up_block_config[-1]
        return self._out_channels if isinstance(self._out_channels, int) else sum(self._out_channels)

    def bottleneck_size(self, in_size):
        down = self.down_block_config
        return in_size // (2 ** (down.count('I') + down.count('D')))

    def output_size(self, in_size):
        up = self.up_block_config
        return self.bottleneck_size(in_size) * (2 ** (up.count('I') + up.count('U')))

    def forward(self, z, z_inject=None, return_intermediate=False):
        if self.input_block is not None:
            z = self.input_block(z)
        skips = []
        for block in self.down_blocks:
            z = block(z)
            skips.insert(0, z)
        if z_inject is not None:
            assert z_inject.size(0) == z.size(0)
            tail = [1] * (z.dim() - 2)
            z = torch.cat((z, z_inject.view(*z_inject.shape, *tail).expand(-1, -1, *z.shape[2:])), dim=1)
        for i, block in enumerate(self.up_blocks):
            if 1 <= i < len(skips):
                z = torch.cat((z, skips[i]), dim=1)
            z = block(z)
        if isinstance(self.output_block, OutputBlock):
            z = self.output_block(z)
        elif self.output_block is not None:
            z = torch.cat([head(z) for head in self.output_block], dim=1)
        return (z, skips) if return_intermediate else z


class UNet2d(BaseUNet):
    def __init__(self, in_channels, out_channels, block_config):
        super().__init__(in_channels, out_channels, block_config, conv_module=EqualizedConv2d)


class UNet3d(BaseUNet):
    def __init__(self, in_channels, out_channels, block_config):
        super().__init__(in_channels, out_channels, block_config, conv_module=EqualizedConv3d)
