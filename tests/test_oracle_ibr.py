"""CPU: the oracle's restatement of the IBR colour branch (oracle/lf_oracle.py, ibr_* functions; reference
latentfusion/ibr.py) against golden vectors produced by the unmodified reference (oracle/make_golden_ibr.py)."""
import os

import pytest
import torch

from tests import parity_helpers as ph
from oracle import lf_oracle as O

IBR_GOLDEN = os.path.join(ph.ROOT, 'tests', 'golden', 'ibr_p24.npz')
TOL = dict(atol=1e-5, rtol=1e-4)


@pytest.fixture(scope='module')
def g():
    return ph.Golden(IBR_GOLDEN)


def test_warp_field_and_reprojection(g):
    cam_in, cam_out = ph.oracle_camera(g.cam('cam_in')), ph.oracle_camera(g.cam('cam_out'))
    grid = O.ibr_warp_field(cam_in, cam_out, g['depth_out'])
    torch.testing.assert_close(grid, g['warp_field'], atol=2e-5, rtol=1e-4)
    img, dep = O.ibr_reproject_views(g['image_in'], g['depth_in'], g['depth_out'], cam_in, cam_out)
    assert img.shape == g['image_reproj'].shape and dep.shape == g['depth_reproj'].shape
    torch.testing.assert_close(img, g['image_reproj'], **TOL)
    torch.testing.assert_close(dep, g['depth_reproj'], **TOL)
    # the fixture must exercise both the interior and the zero-padded exterior of the source views
    inside = (g['warp_field'].abs() <= 1).all(-1).float().mean()
    assert 0.2 < inside < 0.98


@pytest.mark.parametrize('weight_type', ['cam_dist', 'cam_angle', 'cam_hybrid', 'depth'])
def test_render_ibr_weight_types(g, weight_type):
    cam_in, cam_out = ph.oracle_camera(g.cam('cam_in')), ph.oracle_camera(g.cam('cam_out'))
    fake, _ = O.ibr_render(cam_in, cam_out, g['image_in'], g['depth_in'], g['depth_out'], p=0.5,
                           weight_type=weight_type, eps=1e-2)
    torch.testing.assert_close(fake, g[f'render_ibr.{weight_type}'][0], **TOL)


def test_blend_and_warp_blend_logits(g):
    vi = g.meta['VI']
    out, wts = O.ibr_blend_logits(g['logits'][:, :vi], g['image_reproj'])
    torch.testing.assert_close(out, g['blend.image'], **TOL)
    torch.testing.assert_close(wts, g['blend.weights'], **TOL)
    out, wts, dx, dy = O.ibr_warp_blend_logits(g['logits'], g['image_reproj'], 5)
    torch.testing.assert_close(out, g['warp_blend.image'], **TOL)
    torch.testing.assert_close(wts, g['warp_blend.weights'], **TOL)
    torch.testing.assert_close(dx, g['warp_blend.dx'], **TOL)
    torch.testing.assert_close(dy, g['warp_blend.dy'], **TOL)
