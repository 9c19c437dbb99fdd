"""LatentFusionModel façade.  API mirror of reference ``latentfusion/recon/inference.py``
(from_checkpoint :17-29, build_latent_object :73-84, compute_latent_code :86-99, render_full :101-120,
render_latent_object :122-128, render_ibr_basic :130-149, render_ibr :151-192, _render_reprojections :194-217).
The IBR colour branch runs on the forward-only kernels of ``latentfusion_b200/ibr.py``."""
from pathlib import Path

import torch

from . import models
from .. import ibr
from ..observation import Observation
from ..three.batchview import b2bv, bv2b


class LatentFusionModel(object):

    @classmethod
    def from_checkpoint(cls, checkpoint, device='cpu'):
        if isinstance(checkpoint, (Path, str)):
            # reference checkpoints pickle pathlib objects inside 'args' -> weights_only must be False
            checkpoint = torch.load(checkpoint, map_location='cpu', weights_only=False)
        kwargs = checkpoint['args']
        sculptor, fuser, photographer, _, generator = models.load_models(
            checkpoint, device=device, return_generator=True)
        return cls(sculptor, fuser, photographer, kwargs['camera_dist'], device, generator=generator)

    def __init__(self, sculptor, fuser, photographer, camera_dist, device, generator=None):
        self.device = device
        self.sculptor = sculptor.to(device)
        self.fuser = fuser.to(device)
        self.photographer = photographer.to(device)
        self.generator = generator.to(device) if generator is not None else None
        self.camera_dist = camera_dist
        self.input_size = sculptor.in_size
        self.eval()

    def eval(self):
        return self.train(False)

    def train(self, train):
        # Inference façade: eval() also freezes the weights.  The reference leaves requires_grad=True, so
        # every pose-refinement backward also computes (and discards) all conv weight gradients
        # (SURVEY.md §3.1); nothing ever reads them.  train(True) re-enables them.
        for m in (self.sculptor, self.photographer, self.fuser, self.generator):
            if m is not None:
                m.train(train)
                m.requires_grad_(bool(train))
        return self

    def zoom_observation(self, observation):
        if not observation.meta['is_zoomed']:
            return observation.zoom(self.camera_dist, self.input_size)
        return observation

    def preprocess_observation(self, observation):
        if not observation.meta['is_zoomed']:
            observation = observation.zoom(self.camera_dist, self.input_size)
        if not observation.meta['is_prepared']:
            observation = observation.prepare()
        if not observation.meta['is_normalized']:
            observation = observation.normalize()
        return observation

    def build_latent_object(self, observation: Observation):
        observation = self.preprocess_observation(observation).to(self.device)
        with torch.no_grad():
            z_obj, _ = self.sculptor.encode(self.fuser, camera=observation.camera,
                                            color=observation.color.unsqueeze(0),
                                            depth=observation.depth.unsqueeze(0),
                                            mask=observation.mask.unsqueeze(0))
        return z_obj

    def compute_latent_code(self, observation, camera):
        observation = self.preprocess_observation(observation)
        if len(observation) == 1:
            observation = observation.expand(len(camera))
        _, feats = models.autoencode(self.sculptor, self.fuser, self.photographer, camera=camera,
                                     color=observation.color.unsqueeze(1),
                                     depth=observation.depth.unsqueeze(1),
                                     mask=observation.mask.unsqueeze(1))
        return feats

    def render_full(self, z_obj, camera, input_obs=None, p=0.5):
        # (argument order kept from the reference: the zoom box depends only on dist*size)
        camera_zoom = camera.zoom(None, self.camera_dist, self.input_size).to(self.device)
        if input_obs is None:
            pred, _ = self.render_latent_object(z_obj, camera_zoom, apply_mask=True, return_latent=False)
        else:
            pred, _ = self.render_ibr_basic(z_obj, input_obs, camera_zoom, apply_mask=True, return_latent=False, p=p)
        mask = pred['mask']
        depth = camera_zoom.denormalize_depth(pred['depth']) * mask
        out = {'depth': camera_zoom.uncrop(depth)[0], 'mask': camera_zoom.uncrop(mask)[0]}
        if 'color' in pred:
            out['color'] = camera_zoom.uncrop(pred['color'] / 2 + 0.5)[0]
        return out

    def render_latent_object(self, z_obj, camera, return_latent=True, apply_mask=True):
        y, z, _ = self.photographer.decode(z_obj, camera, return_latent=return_latent, apply_mask=apply_mask)
        if return_latent:
            z = z.squeeze(0)       # one object
        return y, z

    # ---- image-based colour rendering (reference inference.py:130-217); forward only
    def render_ibr_basic(self, z_obj, input_obs, camera_out, return_latent=True, apply_mask=True, p=0.5):
        input_obs = self.preprocess_observation(input_obs)
        with torch.no_grad():
            y_ibr, z_ibr = ibr.render_latent_ibr2(
                self.photographer, z_obj, input_obs.camera.clone().to(self.device), camera_out.clone().to(self.device),
                b2bv(input_obs.color, batch_size=1).to(self.device), p=p, weight_type='cam_dist',
                return_latent=return_latent, apply_mask=apply_mask)
        if return_latent:
            z_ibr = z_ibr.squeeze(0)
        return {k: v.squeeze(0) for k, v in y_ibr.items()}, z_ibr

    def render_ibr(self, z_obj, input_obs, camera_out, return_latent=True):
        if self.generator is None:
            raise ValueError("render_ibr needs the IBR generator network (a checkpoint with a 'generator' entry)")
        input_obs = self.preprocess_observation(input_obs)
        with torch.no_grad():
            (y_out, z_out, image_reproj, depth_reproj, _, depth_ibr_out, _, cam_dist_t) = self._render_reprojections(
                z_obj, input_obs.color.to(self.device), input_obs.camera.to(self.device), camera_out.to(self.device))
            if return_latent:
                z_out = z_out.squeeze(0)
            cam_sims = 1.0 - cam_dist_t * 2
            x = torch.cat((image_reproj, depth_reproj,
                           cam_sims[:, :, None, None, None].expand(-1, -1, -1, *image_reproj.shape[-2:])), dim=2)
            x = x.view(-1, x.shape[1] * x.shape[2], x.shape[3], x.shape[4])      # views -> channels
            x = torch.cat((depth_ibr_out, x), dim=1)
            color_ibr, _, _, _ = ibr.warp_blend_logits(self.generator(x), image_reproj, 5)
        y_out['color'] = color_ibr
        return {k: v.squeeze(0) for k, v in y_out.items()}, z_out

    def _render_reprojections(self, z_obj, color_in, camera_in, camera_out, return_latent=True):
        y_in, _, _ = self.photographer.decode(z_obj, camera_in)
        y_out, z_out, _ = self.photographer.decode(z_obj, camera_out, return_latent=return_latent)
        mask_out, depth_out = y_out['mask'], y_out['depth']
        image_reproj, depth_reproj, cam_dist_r, cam_dist_t = ibr.reproject_views_batch(
            color_in.unsqueeze(0), y_in['depth'], y_out['depth'], camera_in, camera_out)
        image_reproj = image_reproj * mask_out.unsqueeze(2)
        depth_reproj = (depth_reproj + 1.0) * mask_out.unsqueeze(2) - 1.0
        return (y_out, z_out, bv2b(image_reproj), bv2b(depth_reproj), bv2b(mask_out), bv2b(depth_out),
                bv2b(cam_dist_r), bv2b(cam_dist_t))
