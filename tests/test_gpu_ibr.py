"""GPU: the IBR colour branch (latentfusion_b200/ibr.py -> csrc/ibr.cu) against the golden vectors of the unmodified
reference (tests/golden/ibr_p24.npz), against the CPU oracle at another size, and — at the full render size — against
the same computation spelled with torch ops on the device (depth_to_warp_field + F.grid_sample)."""
import os

import pytest
import torch
import torch.nn.functional as F

from tests import parity_helpers as ph
from tests.parity_helpers import OUT_TOL

pytestmark = pytest.mark.gpu
IBR_GOLDEN = os.path.join(ph.ROOT, 'tests', 'golden', 'ibr_p24.npz')


@pytest.fixture(scope='module')
def g():
    return ph.Golden(IBR_GOLDEN)


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


def test_reproject_views_vs_golden(g, dev):
    from latentfusion_b200 import ibr
    cam_in, cam_out = ph.product_camera(g.cam('cam_in'), dev), ph.product_camera(g.cam('cam_out'), dev)
    with torch.no_grad():
        grid = ibr.depth_to_warp_field(cam_in, cam_out, g['depth_out'].to(dev))
        img, dep = ibr.reproject_views(g['image_in'].to(dev), g['depth_in'].to(dev), g['depth_out'].to(dev), cam_in, cam_out)
    torch.testing.assert_close(grid.cpu(), g['warp_field'], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(img.cpu(), g['image_reproj'], **OUT_TOL)
    torch.testing.assert_close(dep.cpu(), g['depth_reproj'], **OUT_TOL)


@pytest.mark.parametrize('weight_type', ['cam_dist', 'cam_angle', 'cam_hybrid', 'depth'])
def test_render_ibr_vs_golden(g, dev, weight_type):
    from latentfusion_b200 import ibr
    cam_in, cam_out = ph.product_camera(g.cam('cam_in'), dev), ph.product_camera(g.cam('cam_out'), dev)
    with torch.no_grad():
        fake, reproj = ibr.render_ibr(cam_in, cam_out, g['image_in'].to(dev)[None], g['depth_in'].to(dev)[None],
                                      g['depth_out'].to(dev)[None], p=0.5, weight_type=weight_type, eps=1e-2)
    torch.testing.assert_close(fake.cpu(), g[f'render_ibr.{weight_type}'], **OUT_TOL)
    torch.testing.assert_close(reproj[0].cpu(), g['image_reproj'], **OUT_TOL)


def test_blend_and_warp_blend_vs_golden(g, dev):
    from latentfusion_b200 import ibr
    vi = g.meta['VI']
    reproj, logits = g['image_reproj'].to(dev), g['logits'].to(dev)
    out, wts = ibr.blend_logits(logits[:, :vi], reproj)
    torch.testing.assert_close(out.cpu(), g['blend.image'], **OUT_TOL)
    torch.testing.assert_close(wts.cpu(), g['blend.weights'], **OUT_TOL)
    out, wts, dx, dy = ibr.warp_blend_logits(logits, reproj, 5)
    torch.testing.assert_close(out.cpu(), g['warp_blend.image'], **OUT_TOL)
    torch.testing.assert_close(wts.cpu(), g['warp_blend.weights'], **OUT_TOL)
    torch.testing.assert_close(dx.cpu(), g['warp_blend.dx'], **OUT_TOL)
    torch.testing.assert_close(dy.cpu(), g['warp_blend.dy'], **OUT_TOL)


def test_reproject_vs_oracle_other_size(dev):
    from oracle import lf_oracle as O
    from latentfusion_b200 import ibr
    P, VI, VO, C = 40, 5, 3, 3
    cams_in, _ = ph.synthetic_cameras(VI, P // 2, seed=31, perturb=False)
    cams_out, _ = ph.synthetic_cameras(VO, P // 2, seed=32)
    din, dout = ph.cam_to_dict(cams_in), ph.cam_to_dict(cams_out)
    torch.manual_seed(33)
    image = torch.rand(VI, C, P, P) * 2 - 1
    depth_in = (torch.rand(VI, 1, P, P) - 0.5) * 1.2
    depth_out = (torch.rand(VO, 1, P, P) - 0.5) * 1.2
    ref_img, ref_dep = O.ibr_reproject_views(image, depth_in, depth_out, ph.oracle_camera(din), ph.oracle_camera(dout))
    with torch.no_grad():
        img, dep = ibr.reproject_views(image.to(dev), depth_in.to(dev), depth_out.to(dev),
                                       ph.product_camera(din, dev), ph.product_camera(dout, dev))
    # white-noise images: |d out / d coord| ~ 1 per pixel; a 1-ulp difference of the projected pixel coordinate
    # (values ~300 px) is ~3e-5 px
    torch.testing.assert_close(img.cpu(), ref_img, atol=3e-4, rtol=1e-3)
    torch.testing.assert_close(dep.cpu(), ref_dep, atol=3e-4, rtol=1e-3)


def test_full_size_reprojection_matches_torch_ops_on_device(dev):
    """BASELINE configs[1] extents of the colour branch: 16 reference views into 8 output views at 128^2."""
    from latentfusion_b200 import ibr
    P, VI, VO, C = 128, 16, 8, 3
    cams_in, _ = ph.synthetic_cameras(VI, P // 2, seed=41, perturb=False)
    cams_out, _ = ph.synthetic_cameras(VO, P // 2, seed=42)
    cin, cout = cams_in.to(dev), cams_out.to(dev)
    torch.manual_seed(43)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, P), torch.linspace(-1, 1, P), indexing='ij')
    image = torch.stack([torch.sin(3 * xx + k) * torch.cos(2 * yy - k) for k in range(VI * C)]).view(VI, C, P, P).to(dev)
    depth_in = (0.5 - 0.8 * (xx ** 2 + yy ** 2)).expand(VI, 1, P, P).contiguous().to(dev)
    depth_out = (0.4 - 0.7 * (xx ** 2 + yy ** 2)).expand(VO, 1, P, P).contiguous().to(dev)
    with torch.no_grad():
        img, dep = ibr.reproject_views(image, depth_in, depth_out, cin, cout)
        grid = ibr.depth_to_warp_field(cin, cout, depth_out).reshape(VO * VI, P, P, 2)
        ref = F.grid_sample(image[None].expand(VO, -1, -1, -1, -1).reshape(VO * VI, C, P, P), grid, mode='bilinear',
                            align_corners=False).view(VO, VI, C, P, P)
    assert img.shape == (VO, VI, C, P, P) and dep.shape == (VO, VI, 1, P, P)
    torch.testing.assert_close(img, ref, atol=2e-4, rtol=1e-3)
    assert torch.isfinite(dep).all() and dep.min() >= -1 and dep.max() <= 1


def test_ibr_is_forward_only_and_has_no_cpu_path(dev):
    from latentfusion_b200 import ibr
    g = ph.Golden(IBR_GOLDEN)
    cam_in, cam_out = ph.product_camera(g.cam('cam_in'), dev), ph.product_camera(g.cam('cam_out'), dev)
    image = g['image_in'].to(dev).requires_grad_(True)
    with pytest.raises(NotImplementedError):
        ibr.reproject_views(image, g['depth_in'].to(dev), g['depth_out'].to(dev), cam_in, cam_out)
    with pytest.raises((RuntimeError, ValueError)):
        ibr.reproject_views(g['image_in'], g['depth_in'], g['depth_out'], ph.product_camera(g.cam('cam_in'), 'cpu'),
                            ph.product_camera(g.cam('cam_out'), 'cpu'))


def _ref_blend_logits(logits, image_reproj):
    """latentfusion/ibr.py:231-234 in torch ops (fp64 on the device)"""
    w = torch.softmax(logits, dim=1).unsqueeze(2)
    return (w * image_reproj).sum(dim=1), w


def _ref_warp_blend_logits(logits, image_reproj, flow_size):
    """latentfusion/ibr.py:237-249 in torch ops"""
    vi = image_reproj.shape[1]
    h, w = image_reproj.shape[-2:]
    bl, fx, fy = torch.split(logits, vi, dim=1)
    wts = torch.softmax(bl, dim=1).unsqueeze(2)
    dx, dy = flow_size / w * torch.tanh(fx), flow_size / h * torch.tanh(fy)
    yy, xx = torch.meshgrid([torch.linspace(-1, 1, h, device=logits.device, dtype=logits.dtype),
                             torch.linspace(-1, 1, w, device=logits.device, dtype=logits.dtype)], indexing='ij')
    grid = torch.stack((xx[None, None] + dx, yy[None, None] + dy), dim=-1).clamp(-1, 1)
    samp = F.grid_sample(image_reproj.flatten(0, 1), grid.flatten(0, 1), mode='bilinear', align_corners=False)
    samp = samp.view(*image_reproj.shape)
    return (wts * samp).sum(dim=1), wts, dx, dy


def test_blend_heads_backward_vs_torch_autograd(dev):
    """The two heads the IBR generator is trained through (tools/train/train_ibr.py:367-376): gradients of
    blend_logits / warp_blend_logits w.r.t. the logits, incl. the paths through the returned weights and flows,
    against fp64 autograd of the reference formulation."""
    from latentfusion_b200 import ibr
    torch.manual_seed(11)
    b, vi, c, h, w, flow = 3, 5, 3, 20, 28, 6.0
    image_reproj = torch.rand(b, vi, c, h, w, device=dev) * 2 - 1
    for name in ('blend', 'warp'):
        logits = (torch.randn(b, vi if name == 'blend' else 3 * vi, h, w, device=dev) * 1.5).requires_grad_(True)
        l64 = logits.detach().double().requires_grad_(True)
        g_img = torch.randn(b, c, h, w, device=dev)
        g_aux = torch.randn(b, vi, h, w, device=dev)
        if name == 'blend':
            img, wts = ibr.blend_logits(logits, image_reproj)
            rimg, rw = _ref_blend_logits(l64, image_reproj.double())
            loss = (img * g_img).sum() + (wts.squeeze(2) * g_aux).sum()
            rloss = (rimg * g_img.double()).sum() + (rw.squeeze(2) * g_aux.double()).sum()
        else:
            img, wts, dx, dy = ibr.warp_blend_logits(logits, image_reproj, flow)
            rimg, rw, rdx, rdy = _ref_warp_blend_logits(l64, image_reproj.double(), flow)
            loss = (img * g_img).sum() + (wts.squeeze(2) * g_aux).sum() + (dx * g_aux).sum() - 0.5 * (dy * g_aux).sum()
            rloss = ((rimg * g_img.double()).sum() + (rw.squeeze(2) * g_aux.double()).sum() + (rdx * g_aux.double()).sum()
                     - 0.5 * (rdy * g_aux.double()).sum())
        torch.testing.assert_close(img.double(), rimg, atol=1e-5, rtol=1e-4)
        loss.backward()
        rloss.backward()
        err = (logits.grad.double() - l64.grad).abs()
        # a sample position within rounding of a pixel boundary takes the other bilinear cell in fp32: its flow gradient
        # then differs by a finite jump — allow a handful of such pixels
        assert float((err > 1e-4 + 1e-3 * l64.grad.abs()).float().mean()) < 2e-4, (name, float(err.max()))
