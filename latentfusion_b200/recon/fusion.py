"""Fusing the per-view object-space cubes ``z_obj [B, V, C, D, H, W]`` into one cube per object.

Same factory (``get_fuser``), class names, constructor arguments, ``forward(z_obj, z_cam_mid, z_obj_mid, camera)
-> (z_fused [B, 1, C', D, H, W], extras)`` contract and checkpoint layout as the reference's
``latentfusion/recon/fusion.py`` (factory :17-38, pooling :45-84, concat :87-92, blend :95-149, GRU :152-201,
LSTM :204-246).  Pooling is one pass over the V cubes (``lf_fuse_pool_fwd``); the recurrent fusers run their gate
convolutions on the fused conv kernels and their gate arithmetic in ``lf_gru_gates{1,2}``.
"""
import abc

import torch
from torch import nn

from .. import ops
from ..modules import EqualizedConv2d, EqualizedConv3d, unet
from ..modules.geometry import Camera, CameraToObjectTransform
from ..modules.gru import ConvGRUCell
from ..modules.lstm import ConvLSTMCell
from ..three.batchview import b2bv, bv2b
from . import utils


def pool_tensor(tensor, pool_type, dim=1):
    """max / mean / median / abs_max over the view axis (any ``dim`` is moved to 1 and back)."""
    moved = tensor if dim == 1 else tensor.transpose(dim, 1)
    pooled = ops.fuse_pool(moved, pool_type)
    return pooled if dim == 1 else pooled.transpose(dim, 1)


class Fuser(nn.Module, abc.ABC):
    """Parameter-free fusers checkpoint as just their class name."""

    @abc.abstractmethod
    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera: Camera):
        ...

    def create_checkpoint(self):
        return {'type': type(self).__qualname__}

    @classmethod
    def from_checkpoint(cls, checkpoint):
        return cls()


class _ParamFuser(Fuser):
    """Fusers with weights additionally store their constructor arguments (``ARG_NAMES``) and a state_dict."""
    ARG_NAMES = ()

    def create_checkpoint(self):
        args = {name: getattr(self, name) for name in self.ARG_NAMES}
        return dict(super().create_checkpoint(), args=args, state_dict=self.cpu().state_dict())

    @classmethod
    def from_checkpoint(cls, checkpoint):
        fuser = cls(**checkpoint['args'])
        fuser.load_state_dict(checkpoint['state_dict'])
        return fuser


class PoolFuser(Fuser):
    def __init__(self, pool_type='mean'):
        super().__init__()
        self.pool_type = pool_type

    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera):
        return pool_tensor(z_obj, self.pool_type), {}


class ConcatFuser(Fuser):
    """Views stacked along channels (C' = V * C); a pure reshape."""

    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera):
        batch, views, channels = z_obj.shape[:3]
        return z_obj.reshape(batch, 1, views * channels, *z_obj.shape[3:]), {}


class BlendFuser(_ParamFuser):
    """Per-voxel softmax over the views of a score predicted in CAMERA space (U-Net on the camera-space features plus a
    depth coordinate channel) and resampled into object space; the cubes are averaged with those weights."""
    ARG_NAMES = ('block_config', 'in_channels', 'cube_size')

    def __init__(self, block_config, in_channels, cube_size=1.0, conv_module=EqualizedConv3d):
        super().__init__()
        self.block_config, self.in_channels, self.cube_size = block_config, in_channels, cube_size
        self.unet = unet.BaseUNet(in_channels + 1, 1, block_config, conv_module=conv_module)
        self.transform_block = CameraToObjectTransform(cube_size)

    def compute_blend_weights(self, z_cam, camera):
        views = z_cam.shape[1]
        flat = bv2b(z_cam)
        scores = self.unet(torch.cat((flat, utils.get_normalized_voxel_depth(flat)), dim=1))
        return torch.softmax(b2bv(self.transform_block(scores, camera), views), dim=1)

    def compute_blend_scores(self, z_cam, camera):
        views = z_cam.shape[1]
        flat = bv2b(z_cam)
        scores = self.unet(torch.cat((flat, utils.get_normalized_voxel_depth(flat)), dim=1))
        return b2bv(self.transform_block(scores, camera), views)

    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera):
        if z_obj.is_cuda:           # softmax over the views and the weighted sum in one pass (lf_softmax_blend_*)
            fused, weights = ops.view_softmax_blend(self.compute_blend_scores(z_cam_mid[-1], camera), z_obj)
            return fused, {'blend_weights': weights.squeeze(2)}
        weights = self.compute_blend_weights(z_cam_mid[-1], camera)
        return (z_obj * weights).sum(dim=1, keepdim=True), {'blend_weights': weights.squeeze(2)}


def _scan_views(z_obj, coords, step, state):
    """Run a recurrent cell over views 1..V-1 (view 0 seeds the state); each input is [cube_i, coordinates]."""
    for i in range(1, z_obj.shape[1]):
        state = step(torch.cat((z_obj[:, i], coords), dim=1), state)
    return state


class GRUFuser(_ParamFuser):
    ARG_NAMES = ('in_channels', 'cube_size')

    def __init__(self, in_channels, cube_size=1.0, conv_module=EqualizedConv3d):
        super().__init__()
        self.in_channels, self.cube_size, self.conv_module = in_channels, cube_size, conv_module
        self._planar = conv_module is EqualizedConv2d
        self.gru = ConvGRUCell(in_channels + (2 if self._planar else 3), in_channels, kernel_size=3, bias=True,
                               conv_module=conv_module)

    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera):
        seed = z_obj[:, 0]
        coords = (utils.get_normalized_pixel_coords if self._planar else utils.get_normalized_voxel_coords)(seed)
        if not self._planar and self.in_channels % 4 == 0 and self.gru.splits_inputs(seed):
            # the coordinate channels are the same constant field at every step: convolve them once per gate
            views = z_obj.shape[1] - 1
            if views == 0:
                return seed.unsqueeze(1), {}
            # ... and the view inputs do not depend on the recurrence: their half of every gate is one batched
            # convolution per gate; the per-step convolutions only see the hidden state
            pre_u, pre_r, pre_o = self.gru.input_terms(z_obj[:, 1:], self.gru.extra_terms(coords, self.in_channels), views)
            state = seed
            for i in range(views):
                state = self.gru.step_hidden(state, pre_u[:, i], pre_r[:, i], pre_o[:, i], self.in_channels)
            return state.unsqueeze(1), {}
        return _scan_views(z_obj, coords, self.gru, seed).unsqueeze(1), {}


class LSTMFuser(_ParamFuser):
    ARG_NAMES = ('in_channels', 'cube_size')

    def __init__(self, in_channels, cube_size=1.0, conv_module=EqualizedConv3d):
        super().__init__()
        self.in_channels, self.cube_size = in_channels, cube_size
        self.lstm = ConvLSTMCell(in_channels + 3, in_channels, kernel_size=3, bias=True, conv_module=conv_module)

    def forward(self, z_obj, z_cam_mid, z_obj_mid, camera):
        seed = z_obj[:, 0]
        hidden, _ = _scan_views(z_obj, utils.get_normalized_voxel_coords(seed), self.lstm,
                                (seed, torch.zeros_like(seed)))
        return hidden.unsqueeze(1), {}


_RECURRENT_OR_LEARNED = {'gru': GRUFuser, 'lstm': LSTMFuser}


def get_fuser(fuser_type, in_channels, cube_size, block_config=None, conv_module=EqualizedConv3d):
    """'pool:<max|mean|median|abs_max>', 'concat', 'blend', 'gru' or 'lstm'."""
    kind, _, option = fuser_type.partition(':')
    if kind == 'pool' and option:
        return PoolFuser(option)
    if fuser_type == 'concat':
        return ConcatFuser()
    if fuser_type == 'blend':
        return BlendFuser(block_config, in_channels=in_channels, cube_size=cube_size, conv_module=conv_module)
    if fuser_type in _RECURRENT_OR_LEARNED:
        return _RECURRENT_OR_LEARNED[fuser_type](in_channels=in_channels, cube_size=cube_size, conv_module=conv_module)
    raise ValueError(f"Unknown fuser type {fuser_type!r}")


def from_checkpoint(checkpoint):
    return globals()[checkpoint['type']].from_checkpoint(checkpoint)
