"""LatentFusionModel façade.  API mirror of reference ``latentfusion/recon/inference.py``
(from_checkpoint :17-29, build_latent_object :73-84, compute_latent_code :86-99, render_full :101-120,
render_latent_object :122-128).  The IBR colour branch (:130-217) is outside this path."""
from pathlib import Path

import torch

from . import models
from ..observation import Observation


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
        if input_obs is not None:
            raise NotImplementedError("image-based colour rendering is outside the reconstruct->render path")
        # (argument order kept from the reference: the zoom box depends only on dist*size)
        camera_zoom = camera.zoom(None, self.camera_dist, self.input_size).to(self.device)
        pred, _ = self.render_latent_object(z_obj, camera_zoom, apply_mask=True, return_latent=False)
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
