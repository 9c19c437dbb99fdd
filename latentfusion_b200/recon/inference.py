"""The object-level front door: reconstruct once, then render / image-based-render / autoencode against it.

Method names and arguments follow the reference class of the same name (``latentfusion/recon/inference.py``:
from_checkpoint :17-29, preprocessing :54-71, build_latent_object :73-84, compute_latent_code :86-99, render_full
:101-120, render_latent_object :122-128, render_ibr_basic :130-149, render_ibr :151-192, _render_reprojections
:194-217) so notebooks and scripts written against it keep working; everything below the method boundary runs on
the lfb200 kernels.  Two deliberate differences: ``eval()`` freezes the weights (the reference back-propagates
into every convolution weight on each pose iteration and throws the result away), and the IBR colour
branch is forward-only (``latentfusion_b200/ibr.py``).
"""
from pathlib import Path

import torch

from .. import ibr
from ..observation import Observation
from ..three.batchview import b2bv, bv2b
from . import models

# the preprocessing stages an Observation goes through before the networks see it: (meta flag, how to get there)
_STAGES = (
    ('is_zoomed', lambda self, obs: obs.zoom(self.camera_dist, self.input_size)),
    ('is_prepared', lambda self, obs: obs.prepare()),
    ('is_normalized', lambda self, obs: obs.normalize()),
)


def _per_object(tensors):
    """Drop the leading object axis of every entry (this façade handles one object at a time)."""
    return {name: value.squeeze(0) for name, value in tensors.items()}


class LatentFusionModel(object):

    def __init__(self, sculptor, fuser, photographer, camera_dist, device, generator=None):
        self.device, self.camera_dist, self.input_size = device, camera_dist, sculptor.in_size
        self.sculptor, self.fuser, self.photographer = (m.to(device) for m in (sculptor, fuser, photographer))
        self.generator = None if generator is None else generator.to(device)
        self.eval()

    @classmethod
    def from_checkpoint(cls, checkpoint, device='cpu'):
        if isinstance(checkpoint, (str, Path)):
            # trusted file; the reference pickles pathlib objects in 'args', which weights_only=True rejects
            checkpoint = torch.load(checkpoint, map_location='cpu', weights_only=False)
        nets = models.load_models(checkpoint, device=device, return_generator=True)
        return cls(nets[0], nets[1], nets[2], checkpoint['args']['camera_dist'], device, generator=nets[4])

    # ------------------------------------------------------------------ mode
    def train(self, train):
        for net in (self.sculptor, self.fuser, self.photographer, self.generator):
            if net is None:
                continue
            net.train(train)
            net.requires_grad_(bool(train))       # frozen in eval: no weight gradients in the pose loop
        return self

    def eval(self):
        return self.train(False)

    # ------------------------------------------------------------------ preprocessing
    def _advance(self, observation, upto):
        for flag, step in _STAGES[:upto]:
            if not observation.meta[flag]:
                observation = step(self, observation)
        return observation

    def zoom_observation(self, observation):
        return self._advance(observation, 1)

    def preprocess_observation(self, observation):
        return self._advance(observation, len(_STAGES))

    # ------------------------------------------------------------------ reconstruct / autoencode
    def _inputs(self, observation, axis):
        return {name: getattr(observation, name).unsqueeze(axis) for name in ('color', 'depth', 'mask')}

    @torch.no_grad()
    def build_latent_object(self, observation: Observation):
        obs = self.preprocess_observation(observation).to(self.device)
        return self.sculptor.encode(self.fuser, camera=obs.camera, **self._inputs(obs, 0))[0]     # views of ONE object

    def compute_latent_code(self, observation, camera):
        obs = self.preprocess_observation(observation)
        if len(obs) == 1:
            obs = obs.expand(len(camera))
        return models.autoencode(self.sculptor, self.fuser, self.photographer, camera=camera,
                                 **self._inputs(obs, 1))[1]                                       # one view per object

    # ------------------------------------------------------------------ render
    def render_latent_object(self, z_obj, camera, return_latent=True, apply_mask=True):
        y, z, _ = self.photographer.decode(z_obj, camera, return_latent=return_latent, apply_mask=apply_mask)
        return y, (z.squeeze(0) if return_latent else z)

    def render_full(self, z_obj, camera, input_obs=None, p=0.5):
        """Full-frame depth / mask (/ colour when reference views are given) for full-frame cameras."""
        # the reference passes (dist, size) in this order; the crop box only depends on their product
        cam = camera.zoom(None, self.camera_dist, self.input_size).to(self.device)
        options = dict(apply_mask=True, return_latent=False)
        pred = (self.render_latent_object(z_obj, cam, **options) if input_obs is None
                else self.render_ibr_basic(z_obj, input_obs, cam, p=p, **options))[0]
        # (the reference hands the [1, V, ...] render straight to Camera.uncrop, whose 2-D sampler rejects 5-D input;
        #  the object axis is dropped here so the method actually runs — one object at a time, as everywhere in this class)
        pred = {k: (v.squeeze(0) if v.dim() == 5 else v) for k, v in pred.items()}
        full = {'depth': cam.uncrop(cam.denormalize_depth(pred['depth']) * pred['mask'])[0],
                'mask': cam.uncrop(pred['mask'])[0]}
        if 'color' in pred:
            full['color'] = cam.uncrop(pred['color'] / 2 + 0.5)[0]
        return full

    # ------------------------------------------------------------------ image-based colour (forward only)
    @torch.no_grad()
    def render_ibr_basic(self, z_obj, input_obs, camera_out, return_latent=True, apply_mask=True, p=0.5):
        """Camera-distance-weighted blend of the reprojected reference views (no generator network)."""
        views = self.preprocess_observation(input_obs)
        y, z = ibr.render_latent_ibr2(self.photographer, z_obj,
                                      views.camera.clone().to(self.device), camera_out.clone().to(self.device),
                                      b2bv(views.color, batch_size=1).to(self.device),
                                      p=p, weight_type='cam_dist', return_latent=return_latent, apply_mask=apply_mask)
        return _per_object(y), (z.squeeze(0) if return_latent else z)

    @torch.no_grad()
    def render_ibr(self, z_obj, input_obs, camera_out, return_latent=True):
        """Generator-refined blend: per-view weights and a bounded flow predicted from the reprojections."""
        if self.generator is None:
            raise ValueError("render_ibr needs the IBR generator network (a checkpoint with a 'generator' entry)")
        views = self.preprocess_observation(input_obs)
        y, z, image_reproj, depth_reproj, _, depth_out, _, dist_t = self._render_reprojections(
            z_obj, views.color.to(self.device), views.camera.to(self.device), camera_out.to(self.device))
        height, width = image_reproj.shape[-2:]
        similarity = (1.0 - 2 * dist_t)[:, :, None, None, None].expand(-1, -1, -1, height, width)
        per_view = torch.cat((image_reproj, depth_reproj, similarity), dim=2)            # [Vo, Vi, C+2, H, W]
        stacked = torch.cat((depth_out, per_view.flatten(1, 2)), dim=1)                  # views folded into channels
        y['color'] = ibr.warp_blend_logits(self.generator(stacked), image_reproj, 5)[0]
        return _per_object(y), (z.squeeze(0) if return_latent else z)

    def _render_reprojections(self, z_obj, color_in, camera_in, camera_out, return_latent=True):
        decode = self.photographer.decode
        depth_in = decode(z_obj, camera_in)[0]['depth']
        y_out, z_out, _ = decode(z_obj, camera_out, return_latent=return_latent)
        image, depth, dist_r, dist_t = ibr.reproject_views_batch(color_in.unsqueeze(0), depth_in, y_out['depth'],
                                                                 camera_in, camera_out)
        visible = y_out['mask'].unsqueeze(2)
        image, depth = image * visible, (depth + 1.0) * visible - 1.0                     # background -> far plane
        flat = [bv2b(t) for t in (image, depth, y_out['mask'], y_out['depth'], dist_r, dist_t)]
        return (y_out, z_out, *flat)
