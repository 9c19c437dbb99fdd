"""Sculptor (views -> latent cube) and Photographer (latent cube + cameras -> depth/mask).

API/checkpoint mirror of reference ``latentfusion/recon/models.py`` (load_models :32-70,
autoencode :73-81, Sculptor :84-258, Photographer :261-505).  Same constructor arguments, same
``state_dict`` keys, same forward/encode/decode signatures and return values.

B200-first differences (values identical, work removed):
* ``Photographer.decode`` hands the *un-replicated* cube to the resampler (the reference materialises
  N copies, models.py:493-494); object-space blocks therefore also run once per object, not per camera.
* ``Sculptor.forward`` resamples the last camera-space volume once (the reference does it twice,
  models.py:209-214) and skips the intermediate resamples nobody consumes unless ``keep_mid``.
* every conv+He+bias+LeakyReLU+PixelNorm chain is one kernel; the 2D<->3D projections are single
  depth-expand / depth-collapse GEMMs.
"""
import torch
from torch import nn
from torch.nn import functional as F

from .. import ops
from ..modules import unet, EqualizedConv3d
from ..modules.blocks import create_blocks, OutputBlock3d, OutputBlock2d
from ..modules.geometry import (TileProjection2d3d, FactorProjection2d3d, CameraToObjectTransform, Camera,
                                ObjectToCameraTransform, FactorProjection3d2d)
from ..three import b2bv, bv2b
from . import fusion
from .utils import get_normalized_voxel_depth


def gan_normalize(tensor):
    return tensor * 2.0 - 1.0


def _get_activation(activation_type, relu_slope=0.2):
    table = {None: None, 'none': None, 'lrelu': lambda: nn.LeakyReLU(relu_slope), 'relu': nn.ReLU, 'tanh': nn.Tanh}
    if activation_type not in table:
        raise ValueError(f'Unknown activation type {activation_type}')
    make = table[activation_type]
    return make() if make else None


def _module_device(module):
    return next(module.parameters()).device


def load_models(checkpoint, kwargs=None, device=None, return_generator=False):
    """Rebuild (sculptor, fuser, photographer, discriminator[, generator]) from a reference checkpoint
    dict.  The GAN discriminator is a training-only extra outside this path: it is not rebuilt (None)."""
    if kwargs is None:
        kwargs = checkpoint['args']
    mods = checkpoint['modules']
    s_args, p_args = mods['sculptor']['args'], mods['photographer']['args']
    s_args.setdefault('input_color', True)                      # legacy checkpoints
    if 'input_depth' not in s_args:
        s_args['input_depth'] = kwargs['generator_input_depth']
    if 'input_mask' not in s_args:
        s_args['input_mask'] = kwargs['generator_input_mask']
    for key in ('predict_color', 'predict_depth', 'predict_mask'):
        if key not in p_args:
            p_args[key] = kwargs[key]
    sculptor = Sculptor.from_checkpoint(mods['sculptor']).to(device)
    photographer = Photographer.from_checkpoint(mods['photographer']).to(device)
    fuser = fusion.from_checkpoint(mods['fuser']).to(device)
    discriminator = None
    if return_generator:
        generator = unet.UNet2d.from_checkpoint(mods['generator']).to(device) if 'generator' in mods else None
        return sculptor, fuser, photographer, discriminator, generator
    return sculptor, fuser, photographer, discriminator


def autoencode(sculptor, fuser, photographer, camera, color, depth=None, mask=None):
    z_obj, _ = sculptor.encode(fuser, camera, color, depth, mask)
    y, z_pix, _ = photographer.decode(z_obj, camera, return_latent=True, interpret_logits=True)
    return {k: v.squeeze(1) for k, v in y.items()}, z_pix.squeeze(1)


class Sculptor(nn.Module):

    def __init__(self, in_size, image_config, camera_config, object_config, relu_slope=0.2, cube_size=1.0,
                 cube_activation_type=None, projection_type='tile', input_color=True, input_depth=False,
                 input_mask=True, scale_mode='bilinear', **kwargs):
        super().__init__()
        self.image_config, self.camera_config, self.object_config = image_config, camera_config, object_config
        self.input_color, self.input_depth, self.input_mask = input_color, input_depth, input_mask
        self.relu_slope, self.cube_size = relu_slope, cube_size
        self.cube_activation_type, self.projection_type, self.scale_mode = cube_activation_type, projection_type, scale_mode
        self.in_channels = 3 * bool(input_color) + bool(input_mask) + bool(input_depth)
        self.in_size = in_size
        # set True to also return the per-block object-space intermediates (skip-connection / blend users)
        self.keep_mid = False

        self.image_encoder = unet.UNet2d(self.in_channels, None, self.image_config)
        proj = {'tile': TileProjection2d3d, 'factor': FactorProjection2d3d}.get(projection_type)
        if proj is None:
            raise ValueError(f"Unknown projection type {projection_type!r}")
        self.projection_block = proj(in_channels=self.image_encoder.out_channels,
                                     out_channels=self.camera_config[0], out_size=self.image_out_size)
        self.camera_blocks = create_blocks(self.camera_config, EqualizedConv3d, 0.5, scale_mode=scale_mode)
        self.transform_block = CameraToObjectTransform(cube_size)
        self.object_blocks = (create_blocks(self.object_config, EqualizedConv3d, 0.5, scale_mode=scale_mode)
                              if self.object_config else nn.ModuleList())
        self.output_block = OutputBlock3d(self.out_channels, self.out_channels,
                                          activation=_get_activation(cube_activation_type))

    @property
    def image_out_size(self):
        return self.image_encoder.output_size(self.in_size)

    @property
    def camera_out_size(self):
        return self.image_out_size // (2 ** self.camera_config.count('D'))

    @property
    def out_size(self):
        if self.object_config:
            return self.camera_out_size // (2 ** self.object_config.count('D'))
        return self.camera_out_size

    @property
    def image_bottleneck_size(self):
        return self.image_encoder.bottleneck_size(self.in_size)

    @property
    def out_channels(self):
        return self.object_config[-1] if self.object_config else self.camera_config[-1]

    @classmethod
    def from_checkpoint(cls, checkpoint):
        model = cls(**checkpoint['args'])
        model.load_state_dict(checkpoint['state_dict'])
        return model

    def create_checkpoint(self):
        keys = ('in_channels', 'in_size', 'image_config', 'camera_config', 'object_config', 'relu_slope',
                'cube_size', 'cube_activation_type', 'projection_type', 'input_color', 'input_depth',
                'input_mask', 'scale_mode')
        return {'args': {k: getattr(self, k) for k in keys}, 'state_dict': self.cpu().state_dict()}

    def forward(self, x, camera: Camera, keep_mid=None):
        keep_mid = self.keep_mid if keep_mid is None else keep_mid
        z = self.projection_block(self.image_encoder(x))
        z_cam_mid, z_obj_mid = [], []
        last = len(self.camera_blocks) - 1
        z_obj = None
        for i, block in enumerate(self.camera_blocks):
            z = block(z)
            if keep_mid or i == last:
                z_obj = self.transform_block(z, camera)     # the last one doubles as the main path
                z_cam_mid.append(z_obj)
        if z_obj is None or not len(self.camera_blocks):
            z_obj = self.transform_block(z, camera)
        z = z_obj
        for block in self.object_blocks:
            z = block(z)
            z_obj_mid.append(z)
        return self.output_block(z), z_cam_mid, z_obj_mid

    def encode(self, fuser, camera, color, depth=None, mask=None, data_parallel=False):
        device = _module_device(self)
        num_views = color.shape[1] if color.dim() == 5 else 1
        parts = []
        if self.input_color:
            parts.append(bv2b(color) if color.dim() == 5 else color)
        if self.input_depth:
            parts.append(bv2b(depth) if depth.dim() == 5 else depth)
        if self.input_mask:
            parts.append(gan_normalize(bv2b(mask) if mask.dim() == 5 else mask))
        x = torch.cat([p.to(device) for p in parts], dim=1)
        needs_mid = isinstance(fuser, fusion.BlendFuser)
        z_obj, z_cam_mid, z_obj_mid = self(x, camera.to(device), keep_mid=needs_mid or self.keep_mid)
        z_obj = b2bv(z_obj, num_views)
        z_cam_mid = [b2bv(z, num_views) for z in z_cam_mid]
        z_obj_mid = [b2bv(z, num_views) for z in z_obj_mid]
        return fuser(z_obj, z_cam_mid, z_obj_mid, camera)


class Photographer(nn.Module):

    def __init__(self, in_size, image_config, camera_config, object_config, projection_type='sum',
                 occlusion_config=False, in_views=1, skip_connections=False, relu_slope=0.2, cube_size=1.0,
                 predict_color=False, predict_depth=True, predict_mask=True, scale_mode='bilinear', **kwargs):
        super().__init__()
        self.image_config, self.camera_config = image_config, camera_config
        self.occlusion_config, self.object_config = occlusion_config, object_config
        self.projection_type = projection_type
        self.predict_color, self.predict_depth, self.predict_mask = predict_color, predict_depth, predict_mask
        self.in_views, self.relu_slope, self.skip_connections = in_views, relu_slope, skip_connections
        self.cube_size, self.scale_mode, self.in_size = cube_size, scale_mode, in_size
        self.out_channels = [3] * bool(predict_color) + [1] * bool(predict_depth) + [1] * bool(predict_mask)

        self.object_blocks = (create_blocks(self.object_config, EqualizedConv3d, 2.0, in_views=in_views,
                                            skip_connections=skip_connections, scale_mode=scale_mode)
                              if self.object_config else nn.ModuleList())
        self.transform_block = ObjectToCameraTransform(cube_size)
        self.occlusion_module = (unet.UNet3d(self.object_config[-1] + 1, 1, occlusion_config)
                                 if occlusion_config else None)
        self.camera_blocks = create_blocks(self.camera_config, EqualizedConv3d, 2.0,
                                           skip_connections=skip_connections, skip_connect_start=True,
                                           skip_connection_views=in_views, scale_mode=scale_mode)
        self.projection_block = (FactorProjection3d2d(self.camera_config[-1], self.image_config[0][0],
                                                      out_size=self.camera_out_size)
                                 if projection_type == 'factor' else None)
        self.image_decoder = unet.UNet2d(None, None, self.image_config)
        self.output_blocks = nn.ModuleList([OutputBlock2d(self.image_decoder.out_channels, c)
                                            for c in self.out_channels])

    @property
    def object_out_size(self):
        return self.in_size * (2 ** self.object_config.count('U'))

    @property
    def camera_out_size(self):
        return self.object_out_size * (2 ** self.camera_config.count('U'))

    @property
    def out_size(self):
        return self.image_decoder.output_size(self.camera_out_size)

    @property
    def image_bottleneck_size(self):
        return self.image_decoder.bottleneck_size(self.camera_out_size)

    @classmethod
    def from_checkpoint(cls, checkpoint):
        model = cls(**checkpoint['args'])
        model.load_state_dict(checkpoint['state_dict'])
        return model

    def create_checkpoint(self):
        keys = ('image_config', 'camera_config', 'occlusion_config', 'object_config', 'projection_type',
                'relu_slope', 'out_channels', 'in_views', 'in_size', 'skip_connections', 'cube_size',
                'predict_color', 'predict_depth', 'predict_mask', 'scale_mode')
        return {'args': {k: getattr(self, k) for k in keys}, 'state_dict': self.cpu().state_dict()}

    def _compute_depth_weights(self, z_cam):
        logits = self.occlusion_module(torch.cat((z_cam, get_normalized_voxel_depth(z_cam)), dim=1))
        resized = F.interpolate(logits, z_cam.size(-1))
        return torch.softmax(logits, dim=2), torch.softmax(resized, dim=2)

    def _depth_from_weight(self, depth_weights):
        return (get_normalized_voxel_depth(depth_weights) * depth_weights).sum(dim=2)

    def forward(self, z_obj, camera, z_cam_mid=None, z_obj_mid=None, return_latent=False):
        if len(camera) % z_obj.shape[0] != 0:
            raise ValueError(f"batch dimension of z_obj and camera much match. ({z_obj.shape[0]} != {len(camera)})")
        if z_cam_mid is None and self.skip_connections:
            raise ValueError("z_cam_intermediate required for skip connections.")
        if z_obj_mid is None and self.skip_connections:
            raise ValueError("z_obj_intermediate required for skip connections.")
        if z_obj.shape[0] > 1 and z_obj.stride(0) == 0:
            z_obj = z_obj[:1]                       # an expand()ed cube: sample the one real copy
        if self.skip_connections:
            z_cam_mid = [self.transform_block(z_cam, camera) for z_cam in z_cam_mid]

        z = z_obj
        for i, block in enumerate(self.object_blocks):
            if self.skip_connections and i >= 1:
                z = torch.cat((z, z_obj_mid[-i - 1]), dim=1)
            z = block(z)
        z = self.transform_block(z, camera)
        for i, block in enumerate(self.camera_blocks):
            if self.skip_connections:
                z = torch.cat((z, z_cam_mid[-i - 1]), dim=1)
            z = block(z)

        z_depth = None
        if self.occlusion_module:
            z_weights, resized = self._compute_depth_weights(z)
            z_depth = self._depth_from_weight(z_weights)
            z = z * resized

        if self.projection_type == 'sum':
            z = z.sum(dim=2)
        elif self.projection_type == 'factor':
            if not self.occlusion_module:          # the projection is then the only consumer of the camera block
                z = ops.mark_single_consumer(z)
            z = self.projection_block(z)
        y = self.image_decoder(z)
        if len(self.output_blocks):
            y = torch.cat([head(y) for head in self.output_blocks], dim=1)
        return (y, z, z_depth) if return_latent else (y, None, z_depth)

    def interpret_logits(self, logits, apply_mask=False):
        y, base = {}, 0
        if self.predict_color:
            y['color_logits'] = logits[:, base:base + 3]
            y['color'] = torch.tanh(y['color_logits'])
            base += 3
        if self.predict_depth:
            y['depth_logits'] = logits[:, base:base + 1]
            y['depth'] = torch.tanh(y['depth_logits'])
            base += 1
        if self.predict_mask:
            y['mask_logits'] = logits[:, base:base + 1]
            y['mask'] = torch.sigmoid(y['mask_logits'])
        else:
            y['mask'] = (y['depth'].detach() > -1.0).float()
            y['mask_logits'] = 100 * y['mask'] + (-100) * (1.0 - y['mask'])
        if apply_mask and self.predict_mask:
            keep = y['mask'] > 0.5
            if self.predict_depth:
                y['depth'] = (y['depth'] + 1) * keep - 1
            if self.predict_color:
                y['color'] = y['color'] * keep
        return y

    def decode(self, z_obj, camera, interpret_logits=True, return_latent=False, data_parallel=False,
               apply_mask=False):
        """z_obj [B,1,C,S,S,S] (one cube per object), camera with B*num_views entries (object-major)."""
        num_views = camera.length // z_obj.shape[0]
        if z_obj.shape[1] == 1:
            cubes = z_obj[:, 0]                     # shared per object — never replicated
        else:
            cubes = z_obj.reshape(-1, *z_obj.shape[2:])
        y, z, z_depth = self(cubes, camera, return_latent=return_latent)
        if z is not None:
            z = b2bv(z, num_views)
        if interpret_logits:
            y = {k: b2bv(v, num_views) for k, v in self.interpret_logits(y, apply_mask=apply_mask).items()}
        return y, z, z_depth
