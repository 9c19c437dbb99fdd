"""Multi-view latent-volume fusion along the view axis.  API/checkpoint mirror of reference
``latentfusion/recon/fusion.py`` (get_fuser :17-38, pool_tensor :45-57, PoolFuser :77-84,
ConcatFuser :87-92, BlendFuser :95-149, GRUFuser :152-201, LSTMFuser :204-246)."""
import abc

import torch
from torch import nn

from .. import ops
from ..modules import unet, EqualizedConv2d, EqualizedConv3d
from ..modules.geometry import CameraToObjectTransform, Camera
from ..modules.gru import ConvGRUCell
from ..modules.lstm import ConvLSTMCell
from ..three.batchview import bv2b, b2bv
from . import utils


def get_fuser(fuser_type, in_channels, cube_size, block_config=None, conv_module=EqualizedConv3d):
    if fuser_type.startswith('pool:'):
        return PoolFuser(fuser_type.split(':')[1])
    if fuser_type == 'concat':
        return ConcatFuser()
    if fuser_type == 'blend':
        return BlendFuser(block_config, in_channels=in_channels, cube_size=cube_size, conv_module=conv_module)
    if fuser_type == 'gru':
        return GRUFuser(in_channels=in_channels, cube_size=cube_size, conv_module=conv_module)
    if fuser_type == 'lstm':
        return LSTMFuser(in_channels=in_channels, cube_size=cube_size, conv_module=conv_module)
    raise ValueError(f"Unknown fuser type {fuser_type!r}")


def from_checkpoint(checkpoint):
    return globals()[checkpoint['type']].from_checkpoint(checkpoint)


def pool_tensor(tensor, pool_type, dim=1):
    """View-axis pooling of [B,V,C,D,H,W] (one HBM pass, ``lf_fuse_pool_fwd``)."""
    if dim != 1:
        tensor = tensor.transpose(dim, 1)
    out = ops.fuse_pool(tensor, pool_type)
    return out if dim == 1 else out.transpose(dim, 1)


class Fuser(nn.Module, abc.ABC):

    @classmethod
    def from_checkpoint(cls, checkpoint):
        return cls()

    def create_checkpoint(self):
        return {'type': self.__class__.__qualname__}

    @abc.abstractmethod
    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera: Camera):
        raise NotImplementedError


class _ParamFuser(Fuser):
    @classmethod
    def from_checkpoint(cls, checkpoint):
        model = cls(**checkpoint['args'])
        model.load_state_dict(checkpoint['state_dict'])
        return model

    def _args(self):
        raise NotImplementedError

    def create_checkpoint(self):
        return {**super().create_checkpoint(), 'args': self._args(), 'state_dict': self.cpu().state_dict()}


class PoolFuser(Fuser):
    def __init__(self, pool_type='mean'):
        super().__init__()
        self.pool_type = pool_type

    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera):
        return pool_tensor(z_obj, self.pool_type, dim=1), {}


class ConcatFuser(Fuser):
    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera):
        n, v, c, d, h, w = z_obj.size()
        return z_obj.reshape(n, 1, v * c, d, h, w), {}


class BlendFuser(_ParamFuser):
    def __init__(self, block_config, in_channels, cube_size=1.0, conv_module=EqualizedConv3d):
        super().__init__()
        self.block_config, self.in_channels, self.cube_size = block_config, in_channels, cube_size
        self.unet = unet.BaseUNet(in_channels + 1, 1, block_config, conv_module=conv_module)
        self.transform_block = CameraToObjectTransform(cube_size)

    def _args(self):
        return {'block_config': self.block_config, 'in_channels': self.in_channels, 'cube_size': self.cube_size}

    def compute_blend_weights(self, z_cam, camera):
        num_views = z_cam.shape[1]
        z_cam = bv2b(z_cam)
        w = torch.cat((z_cam, utils.get_normalized_voxel_depth(z_cam)), dim=1)
        w = b2bv(self.transform_block(self.unet(w), camera), num_views)
        return torch.softmax(w, dim=1)

    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera):
        weights = self.compute_blend_weights(z_cam_mid[-1], camera)
        return torch.sum(z_obj * weights, dim=1, keepdim=True), {'blend_weights': weights.squeeze(2)}


class GRUFuser(_ParamFuser):
    def __init__(self, in_channels, cube_size=1.0, conv_module=EqualizedConv3d):
        super().__init__()
        self.in_channels, self.cube_size, self.conv_module = in_channels, cube_size, conv_module
        coord_channels = 2 if conv_module is EqualizedConv2d else 3
        self.gru = ConvGRUCell(in_channels + coord_channels, in_channels, kernel_size=3, bias=True,
                               conv_module=conv_module)

    def _args(self):
        return {'in_channels': self.in_channels, 'cube_size': self.cube_size}

    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera):
        h = z_obj[:, 0]
        coords = (utils.get_normalized_pixel_coords(h) if self.conv_module is EqualizedConv2d
                  else utils.get_normalized_voxel_coords(h))
        for i in range(1, z_obj.shape[1]):
            h = self.gru(torch.cat((z_obj[:, i], coords), dim=1), h)
        return h.unsqueeze(1), {}


class LSTMFuser(_ParamFuser):
    def __init__(self, in_channels, cube_size=1.0, conv_module=EqualizedConv3d):
        super().__init__()
        self.in_channels, self.cube_size = in_channels, cube_size
        self.lstm = ConvLSTMCell(in_channels + 3, in_channels, kernel_size=3, bias=True, conv_module=conv_module)

    def _args(self):
        return {'in_channels': self.in_channels, 'cube_size': self.cube_size}

    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera):
        h = z_obj[:, 0]
        c = torch.zeros_like(h)
        coords = utils.get_normalized_voxel_coords(h)
        for i in range(1, z_obj.shape[1]):
            h, c = self.lstm(torch.cat((z_obj[:, i], coords), dim=1), (h, c))
        return h.unsqueeze(1), {}
