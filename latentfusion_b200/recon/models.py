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
from ..modules.blocks import create_blocks, Block, OutputBlock3d, OutputBlock2d
from ..modules.geometry import (TileProjection2d3d, FactorProjection2d3d, CameraToObjectTransform, Camera,
                                ObjectToCameraTransform, FactorProjection3d2d)
from ..three import b2bv, bv2b
from . import fusion
from .utils import get_normalized_voxel_depth


def gan_normalize(tensor):
    """[0, 1] -> [-1, 1] (how masks are fed to the encoder)."""
    return tensor * 2.0 - 1.0


_ACTIVATIONS = {'lrelu': lambda slope: nn.LeakyReLU(slope), 'relu': lambda slope: nn.ReLU(),
                'tanh': lambda slope: nn.Tanh()}


def _get_activation(activation_type, relu_slope=0.2):
    if activation_type in (None, 'none'):
        return None
    if activation_type not in _ACTIVATIONS:
        raise ValueError(f'Unknown activation type {activation_type}')
    return _ACTIVATIONS[activation_type](relu_slope)


def _resolution(size, tokens, token):
    """`size` after the 'U' (x2) or 'D' (/2) steps listed in a block config"""
    steps = list(tokens).count(token)
    return size << steps if token == 'U' else size >> steps


def _as_batch(t):
    """[B, V, ...] -> [B*V, ...]; 4-D tensors are already flat"""
    return bv2b(t) if t.dim() == 5 else t


class _Checkpointable:
    """create_checkpoint()/from_checkpoint() in the reference's layout: {'args': ctor kwargs, 'state_dict': ...}."""
    CKPT_FIELDS = ()

    def create_checkpoint(self):
        return {'args': {name: getattr(self, name) for name in self.CKPT_FIELDS},
                'state_dict': self.cpu().state_dict()}

    @classmethod
    def from_checkpoint(cls, checkpoint):
        net = cls(**checkpoint['args'])
        net.load_state_dict(checkpoint['state_dict'])
        return net


def load_models(checkpoint, kwargs=None, device=None, return_generator=False):
    """(sculptor, fuser, photographer, discriminator[, generator]) from a reference training checkpoint.  The GAN
    discriminator is a training-only extra outside this path and is returned as None."""
    run_args = checkpoint['args'] if kwargs is None else kwargs
    parts = checkpoint['modules']
    # checkpoints written before these switches existed take them from the training arguments
    legacy = {'sculptor': {'input_color': True, 'input_depth': run_args.get('generator_input_depth'),
                           'input_mask': run_args.get('generator_input_mask')},
              'photographer': {k: run_args.get(k) for k in ('predict_color', 'predict_depth', 'predict_mask')}}
    for part, defaults in legacy.items():
        for key, value in defaults.items():
            if key not in parts[part]['args']:
                if value is None:
                    raise KeyError(f"checkpoint lacks {part} argument {key!r} and the run arguments do not define it")
                parts[part]['args'][key] = value
    sculptor = Sculptor.from_checkpoint(parts['sculptor']).to(device)
    photographer = Photographer.from_checkpoint(parts['photographer']).to(device)
    fuser = fusion.from_checkpoint(parts['fuser']).to(device)
    if not return_generator:
        return sculptor, fuser, photographer, None
    generator = unet.UNet2d.from_checkpoint(parts['generator']).to(device) if 'generator' in parts else None
    return sculptor, fuser, photographer, None, generator


def autoencode(sculptor, fuser, photographer, camera, color, depth=None, mask=None):
    """views -> cube -> the same views again; returns (interpreted outputs, pre-decoder latent), view axis dropped"""
    cube, _ = sculptor.encode(fuser, camera, color, depth, mask)
    outputs, latent, _ = photographer.decode(cube, camera, return_latent=True, interpret_logits=True)
    return {name: value.squeeze(1) for name, value in outputs.items()}, latent.squeeze(1)


class Sculptor(_Checkpointable, nn.Module):
    CKPT_FIELDS = ('in_channels', 'in_size', 'image_config', 'camera_config', 'object_config', 'relu_slope',
                   'cube_size', 'cube_activation_type', 'projection_type', 'input_color', 'input_depth',
                   'input_mask', 'scale_mode')
    _LIFTS = {'tile': TileProjection2d3d, 'factor': FactorProjection2d3d}

    def __init__(self, in_size, image_config, camera_config, object_config, relu_slope=0.2, cube_size=1.0,
                 cube_activation_type=None, projection_type='tile', input_color=True, input_depth=False,
                 input_mask=True, scale_mode='bilinear', **kwargs):
        super().__init__()
        if projection_type not in self._LIFTS:
            raise ValueError(f"Unknown projection type {projection_type!r}")
        for name, value in dict(in_size=in_size, image_config=image_config, camera_config=camera_config,
                                object_config=object_config, relu_slope=relu_slope, cube_size=cube_size,
                                cube_activation_type=cube_activation_type, projection_type=projection_type,
                                input_color=input_color, input_depth=input_depth, input_mask=input_mask,
                                scale_mode=scale_mode).items():
            setattr(self, name, value)
        self.in_channels = 3 * bool(input_color) + bool(input_depth) + bool(input_mask)
        self.keep_mid = False      # True: also return the per-block object-space intermediates (skip / blend users)

        self.image_encoder = unet.UNet2d(self.in_channels, None, image_config)
        self.projection_block = self._LIFTS[projection_type](in_channels=self.image_encoder.out_channels,
                                                             out_channels=camera_config[0],
                                                             out_size=self.image_out_size)
        self.camera_blocks = create_blocks(camera_config, EqualizedConv3d, 0.5, scale_mode=scale_mode)
        self.transform_block = CameraToObjectTransform(cube_size)
        self.object_blocks = (create_blocks(object_config, EqualizedConv3d, 0.5, scale_mode=scale_mode)
                              if object_config else nn.ModuleList())
        self.output_block = OutputBlock3d(self.out_channels, self.out_channels,
                                          activation=_get_activation(cube_activation_type))

    # sizes along the pipeline: image encoder -> camera-space blocks -> object-space blocks
    image_out_size = property(lambda self: self.image_encoder.output_size(self.in_size))
    image_bottleneck_size = property(lambda self: self.image_encoder.bottleneck_size(self.in_size))
    camera_out_size = property(lambda self: _resolution(self.image_out_size, self.camera_config, 'D'))
    out_size = property(lambda self: _resolution(self.camera_out_size, self.object_config or (), 'D'))
    out_channels = property(lambda self: (self.object_config or self.camera_config)[-1])

    def forward(self, x, camera: Camera, keep_mid=None):
        keep_mid = self.keep_mid if keep_mid is None else keep_mid
        z = self.projection_block(self.image_encoder(x))
        cam_mid, obj_mid, resampled = [], [], None
        final = len(self.camera_blocks) - 1
        for index, block in enumerate(self.camera_blocks):
            z = block(z)
            if keep_mid or index == final:
                # (the reference resamples the last camera-space volume twice; once is enough: same values)
                resampled = self.transform_block(z, camera)
                cam_mid.append(resampled)
        z = resampled if resampled is not None else self.transform_block(z, camera)
        for block in self.object_blocks:
            z = block(z)
            obj_mid.append(z)
        return self.output_block(z), cam_mid, obj_mid

    def encode(self, fuser, camera, color, depth=None, mask=None, data_parallel=False):
        """color/depth/mask [B, V, c, H, W] (or flat [B*V, c, H, W]) -> fuser(z_obj [B, V, C, D, H, W], ...)."""
        device = next(self.parameters()).device
        views = color.shape[1] if color.dim() == 5 else 1
        sources = ((self.input_color, color, None), (self.input_depth, depth, None),
                   (self.input_mask, mask, gan_normalize))
        planes = []
        for enabled, tensor, prepare in sources:
            if enabled:
                flat = _as_batch(tensor)
                planes.append((prepare(flat) if prepare else flat).to(device))
        want_mid = self.keep_mid or isinstance(fuser, fusion.BlendFuser)
        cube, cam_mid, obj_mid = self(torch.cat(planes, dim=1), camera.to(device), keep_mid=want_mid)
        per_view = lambda t: b2bv(t, views)                                    # noqa: E731
        return fuser(per_view(cube), [per_view(t) for t in cam_mid], [per_view(t) for t in obj_mid], camera)


class Photographer(_Checkpointable, nn.Module):
    CKPT_FIELDS = ('image_config', 'camera_config', 'occlusion_config', 'object_config', 'projection_type',
                   'relu_slope', 'out_channels', 'in_views', 'in_size', 'skip_connections', 'cube_size',
                   'predict_color', 'predict_depth', 'predict_mask', 'scale_mode')

    def __init__(self, in_size, image_config, camera_config, object_config, projection_type='sum',
                 occlusion_config=False, in_views=1, skip_connections=False, relu_slope=0.2, cube_size=1.0,
                 predict_color=False, predict_depth=True, predict_mask=True, scale_mode='bilinear', **kwargs):
        super().__init__()
        for name, value in dict(in_size=in_size, image_config=image_config, camera_config=camera_config,
                                object_config=object_config, projection_type=projection_type,
                                occlusion_config=occlusion_config, in_views=in_views,
                                skip_connections=skip_connections, relu_slope=relu_slope, cube_size=cube_size,
                                predict_color=predict_color, predict_depth=predict_depth, predict_mask=predict_mask,
                                scale_mode=scale_mode).items():
            setattr(self, name, value)
        # one head per predicted quantity, in the channel order (colour, depth, mask)
        self.out_channels = [3] * bool(predict_color) + [1] * bool(predict_depth) + [1] * bool(predict_mask)

        blocks3d = dict(conv_module=EqualizedConv3d, scale_factor=2.0, scale_mode=scale_mode,
                        skip_connections=skip_connections)
        self.object_blocks = (create_blocks(object_config, in_views=in_views, **blocks3d)
                              if object_config else nn.ModuleList())
        self.transform_block = ObjectToCameraTransform(cube_size)
        self.occlusion_module = unet.UNet3d(object_config[-1] + 1, 1, occlusion_config) if occlusion_config else None
        self.camera_blocks = create_blocks(camera_config, skip_connect_start=True, skip_connection_views=in_views,
                                           **blocks3d)
        if projection_type == 'factor' and len(self.camera_blocks) and not occlusion_config:
            self.camera_blocks[-1].emit_out_split = True        # see Block.forward / ops._EqConv (fused collapse backward)
        self.projection_block = (FactorProjection3d2d(camera_config[-1], image_config[0][0],
                                                      out_size=self.camera_out_size)
                                 if projection_type == 'factor' else None)
        self.image_decoder = unet.UNet2d(None, None, image_config)
        self.output_blocks = nn.ModuleList(OutputBlock2d(self.image_decoder.out_channels, c)
                                           for c in self.out_channels)

    object_out_size = property(lambda self: _resolution(self.in_size, self.object_config, 'U'))
    camera_out_size = property(lambda self: _resolution(self.object_out_size, self.camera_config, 'U'))
    out_size = property(lambda self: self.image_decoder.output_size(self.camera_out_size))
    image_bottleneck_size = property(lambda self: self.image_decoder.bottleneck_size(self.camera_out_size))

    # ---- optional occlusion reasoning (off in the released recipe): a softmax over depth of a U-Net score
    def _compute_depth_weights(self, z_cam):
        scores = self.occlusion_module(torch.cat((z_cam, get_normalized_voxel_depth(z_cam)), dim=1))
        return torch.softmax(scores, dim=2), torch.softmax(F.interpolate(scores, z_cam.size(-1)), dim=2)

    def _depth_from_weight(self, depth_weights):
        return (depth_weights * get_normalized_voxel_depth(depth_weights)).sum(dim=2)

    def forward(self, z_obj, camera, z_cam_mid=None, z_obj_mid=None, return_latent=False):
        """z_obj [B, C, S, S, S] (one cube per object; a stride-0 expand()ed batch is collapsed to its single copy),
        camera with B*views entries -> (logits [B*views, heads, P, P], pre-decoder latent or None, depth or None)."""
        if len(camera) % z_obj.shape[0] != 0:
            raise ValueError(f"batch dimension of z_obj and camera much match. ({z_obj.shape[0]} != {len(camera)})")
        if self.skip_connections:
            for given, label in ((z_cam_mid, 'z_cam_intermediate'), (z_obj_mid, 'z_obj_intermediate')):
                if given is None:
                    raise ValueError(f"{label} required for skip connections.")
            z_cam_mid = [self.transform_block(t, camera) for t in z_cam_mid]
        if z_obj.shape[0] > 1 and z_obj.stride(0) == 0:
            z_obj = z_obj[:1]

        z = z_obj
        for depth, block in enumerate(self.object_blocks):
            if self.skip_connections and depth > 0:
                z = torch.cat((z, z_obj_mid[-depth - 1]), dim=1)
            z = block(z)
        # the single cube is read through L2 by every camera.  In the pose loop (frozen weights, no skip tensors to
        # concatenate) the frustum volumes' only reader is the first camera block's depth-batched convolution, which stages
        # the split-planar layout: K1 then writes that layout directly and the dense fp32 copy never exists.
        first = self.camera_blocks[0] if len(self.camera_blocks) else None
        split_only = bool(z.is_cuda and not self.skip_connections and isinstance(first, Block)
                          and ops.o2c_split_ok(z.shape[1], z.shape[-1], len(camera), first.conv1))
        z = self.transform_block(z, camera, split_only=split_only)
        for depth, block in enumerate(self.camera_blocks):
            if self.skip_connections:
                z = torch.cat((z, z_cam_mid[-depth - 1]), dim=1)
            z = block(z)

        z_depth = None
        if self.occlusion_module:
            weights, weights_resized = self._compute_depth_weights(z)
            z_depth, z = self._depth_from_weight(weights), z * weights_resized

        if self.projection_type == 'factor':
            # without the occlusion branch the projection is the only consumer of the camera block's output
            z = self.projection_block(z if self.occlusion_module else ops.mark_single_consumer(z))
        elif self.projection_type == 'sum':
            z = ops.depth_sum(z) if z.is_cuda else z.sum(dim=2)
        y = self.image_decoder(z)
        if len(self.output_blocks):
            y = self._heads(y)
        return y, (z if return_latent else None), z_depth

    def _heads(self, features):
        """depth / mask / colour logits: one pass for all heads when they are plain 1x1 convolutions"""
        convs = [head.conv for head in self.output_blocks]
        plain = all(head.activation is None and conv.module.kernel_size == (1, 1)
                    for head, conv in zip(self.output_blocks, convs))
        if (plain and features.is_cuda
                and ops.heads_supported(features.shape[1], sum(conv.module.out_channels for conv in convs))):
            return ops.fused_heads(features, [conv.module.weight for conv in convs], [conv.bias for conv in convs])
        return torch.cat([head(features) for head in self.output_blocks], dim=1)

    def interpret_logits(self, logits, apply_mask=False):
        """split the head channels and squash them: colour/depth tanh, mask sigmoid; without a mask head the mask is
        'depth above the far plane'.  apply_mask gates depth (to the far plane) and colour by mask > 0.5."""
        out, cursor = {}, 0
        for name, width, enabled, squash in (('color', 3, self.predict_color, torch.tanh),
                                             ('depth', 1, self.predict_depth, torch.tanh),
                                             ('mask', 1, self.predict_mask, torch.sigmoid)):
            if enabled:
                out[f'{name}_logits'] = logits[:, cursor:cursor + width]
                out[name] = squash(out[f'{name}_logits'])
                cursor += width
        if not self.predict_mask:
            out['mask'] = (out['depth'].detach() > -1.0).float()
            out['mask_logits'] = 100 * out['mask'] + (-100) * (1.0 - out['mask'])
        elif apply_mask:
            inside = out['mask'] > 0.5
            if self.predict_depth:
                out['depth'] = (out['depth'] + 1) * inside - 1
            if self.predict_color:
                out['color'] = out['color'] * inside
        return out

    def decode(self, z_obj, camera, interpret_logits=True, return_latent=False, data_parallel=False,
               apply_mask=False):
        """z_obj [B, 1, C, S, S, S] (a cube per object, never replicated per view) or [B, V, ...]; cameras are
        object-major.  Returns (outputs with a view axis, latent with a view axis or None, depth or None)."""
        views = camera.length // z_obj.shape[0]
        cubes = z_obj[:, 0] if z_obj.shape[1] == 1 else z_obj.flatten(0, 1)
        y, latent, z_depth = self(cubes, camera, return_latent=return_latent)
        if latent is not None:
            latent = b2bv(latent, views)
        if interpret_logits:
            y = {name: b2bv(value, views) for name, value in self.interpret_logits(y, apply_mask=apply_mask).items()}
        return y, latent, z_depth
