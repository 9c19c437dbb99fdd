"""dev probe: camera-gradient error at config B vs the reference's fp64 golden for forward/backward precision combos."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import parity_helpers as ph
from latentfusion_b200 import ops

dev = torch.device('cuda:0')
g, model, z_obj, target = ph.config_b_case(dev, smooth='--smooth' in sys.argv)
g32 = torch.cat([g['grad.log_quaternion'], g['grad.translation'], g['grad.viewport']], 1).double()
g64 = torch.cat([g['grad64.log_quaternion'], g['grad64.translation'], g['grad64.viewport']], 1)
names = ('ov_depth', 'depth', 'iou', 'mask')
w = g.meta['weights']


def run(fwd, bwd, torch_loss=False):
    ops.set_default_precision(fwd)
    ops._bwd_precision_override = bwd
    cam = ph.product_camera(g.cam('hyp_cam'), dev, requires_grad=True)
    y, latent = model.render_latent_object(z_obj, cam, return_latent=True, apply_mask=True)
    if torch_loss:
        from latentfusion_b200.pose import estimation
        losses = estimation.default_pose_loss(target, cam.denormalize_depth(y['depth'].squeeze(0)), y['mask_logits'].squeeze(0), cam)
        sum(w[k] * losses[k] for k in names).mean().backward()
    else:
        terms = ops.pose_loss_terms(y['depth_logits'].squeeze(0)[:, 0], y['mask_logits'].squeeze(0)[:, 0], cam.viewport,
                                    cam.translation[:, 2], target.depth, target.mask, cam.z_span, 0.01, cam.width, cam.height)
        sum(w[k] * terms[:, i] for i, k in enumerate(names)).mean().backward()
    ours = torch.cat([cam.log_quaternion.grad, cam.translation.grad, cam.viewport.grad], 1).cpu().double()
    dl = (y['depth_logits'].cpu().double() - g['render64.depth_logits']).abs().max()
    out = []
    for sl in (slice(0, 3), slice(3, 6), slice(6, 10)):
        sc = g64[:, sl].abs().max()
        out.append(float((ours[:, sl] - g64[:, sl]).abs().max() / sc))
    return float(dl), out


ref = []
for sl in (slice(0, 3), slice(3, 6), slice(6, 10)):
    ref.append(float((g32[:, sl] - g64[:, sl]).abs().max() / g64[:, sl].abs().max()))
print('reference fp32 vs fp64 rel grad err (lq, t, vp):', ['%.2e' % v for v in ref])
dl, e = run(0, 0, True)
print('torch loss head, fwd 0 bwd 0: grad rel err', ['%.2e' % v for v in e])
dl, e = run(1, 1, True)
print('torch loss head, fwd 1 bwd 1: grad rel err', ['%.2e' % v for v in e])
for fwd, bwd in ((0, 0), (1, 1), (1, 2), (2, 2)):
    dl, e = run(fwd, bwd)
    print(f'fwd {fwd} bwd {bwd}: depth-logit abs err vs fp64 {dl:.2e}; grad rel err', ['%.2e' % v for v in e])
