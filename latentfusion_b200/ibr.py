"""Image-based rendering of the reference views into novel views — API mirror of reference ``latentfusion/ibr.py``
(depth_to_warp_field :11-52, reproject_views :55-93, reproject_views_batch :96-139, render_latent_ibr :142-155,
render_latent_ibr2 :158-178, render_ibr :181-228, blend_logits :231-234, warp_blend_logits :237-249).

The arithmetic runs in three lfb200 kernels (``csrc/ibr.cu``): the per-(output view, input view) reprojection of
colour and depth, the view blend, and the flow-refined blend.  The reprojection is forward-only (the reference's IBR
training keeps the reconstruction networks frozen and reprojects without grad, tools/train/train_ibr.py:319-340; asking
it for a gradient raises); ``blend_logits`` / ``warp_blend_logits`` — the heads the IBR generator is trained through,
train_ibr.py:367-376 — are differentiable w.r.t. their logits (``lf_ibr_blend_bwd`` / ``lf_ibr_warp_blend_bwd``).
The [V_o, V_i] view weights (camera-distance softmaxes) are a handful of floats and stay in torch.
"""
import math

import torch

from . import ops, three
from .three.batchview import bv2b


def outer_distance(x1, x2, metric='cosine', eps=1e-8):
    """reference latentfusion/distances.py:27-42 (the metrics this module uses)."""
    if metric == 'cosine':
        w1, w2 = torch.norm(x1, dim=1, keepdim=True), torch.norm(x2, dim=1, keepdim=True)
        return 1.0 - (x1 @ x2.t()) / (w1 @ w2.t()).clamp(min=eps)
    if metric == 'euclidean':
        return torch.cdist(x1, x2)
    if metric == 'inner':
        return -(x1 @ x2.t())
    raise ValueError(f'Unknown type {metric!r}')


def depth_to_warp_field(source_cam, target_cam, target_depth):
    """Warp field [V_o, V_i, H, W, 2] (F.grid_sample convention) taking source-view images to the target views.
    Provided for API parity and inspection; ``reproject_views`` generates it in registers instead."""
    height, width = target_depth.shape[-2:]
    xx, yy, zz = target_cam.depth_camera_coords(target_cam.denormalize_depth(target_depth))
    cam_coords = three.grid_to_coords(torch.stack((xx, yy, zz), dim=-1))
    obj_coords = three.transform_coords(cam_coords, target_cam.cam_to_obj)
    vo, vi = target_cam.length, source_cam.length
    obj_coords = bv2b(obj_coords[:, None].expand(-1, vi, -1, -1))
    obj_to_pix = bv2b(source_cam.obj_to_image[None].expand(vo, -1, -1, -1))
    pix = three.transform_coords(obj_coords, obj_to_pix)
    vp = source_cam.viewport.repeat(vo, 1)
    sw, sh = vp[:, 2] - vp[:, 0], vp[:, 3] - vp[:, 1]
    grid = torch.stack((((pix[..., 0] - vp[:, 0, None]) / sw[:, None]) * 2 - 1,
                        ((pix[..., 1] - vp[:, 1, None]) / sh[:, None]) * 2 - 1), dim=-1)
    return grid.view(vo, vi, height, width, 2)


def reproject_views(image_in, depth_in, depth_out, camera_in, camera_out):
    """image_in [V_i,C,H,W], depth_in [V_i,1,H,W], depth_out [V_o,1,H,W] ->
    (image_reproj [V_o,V_i,C,H,W], depth_reproj [V_o,V_i,1,H,W]).  One kernel launch."""
    return ops.ibr_reproject(image_in, depth_in, depth_out, camera_in.ibr_block(), camera_out.ibr_block())


def reproject_views_batch(image_in, depth_in, depth_out, camera_in, camera_out):
    """image_in/depth_in [B,V_i,C,H,W], depth_out [B,V_o,1,H,W]; cameras are flat over (B, V)."""
    num_objects, in_views, out_views = image_in.shape[0], image_in.shape[1], depth_out.shape[1]
    images, depths, dists_r, dists_t = [], [], [], []
    for i in range(num_objects):
        cin = camera_in[i * in_views:(i + 1) * in_views]
        cout = camera_out[i * out_views:(i + 1) * out_views]
        dists_r.append(three.quaternion.angular_distance(cout.quaternion, cin.quaternion, eps=1e-2) / math.pi)
        dists_t.append(outer_distance(cout.position, cin.position, metric='cosine') / 2.0)
        img, dep = reproject_views(image_in[i], depth_in[i], depth_out[i], cin, cout)
        images.append(img)
        depths.append(dep)
    return torch.stack(images, 0), torch.stack(depths, 0), torch.stack(dists_r, 0), torch.stack(dists_t, 0)


def _view_weights(cam_in, cam_out, weight_type, p, eps, depth_reproj=None, depth_out=None):
    """[V_o, V_i] (camera-based) or [V_o, V_i, H, W] (depth-based) softmax weights; ibr.py:196-222."""
    if weight_type == 'cam_dist':
        d = outer_distance(cam_out.position, cam_in.position, metric='cosine', eps=eps) / 2.0
    elif weight_type == 'cam_angle':
        d = three.quaternion.angular_distance(cam_out.quaternion, cam_in.quaternion) / math.pi
    elif weight_type == 'cam_hybrid':
        dt = outer_distance(cam_out.position, cam_in.position, metric='cosine') / 2.0
        dr = (three.quaternion.angular_distance(cam_out.quaternion, cam_in.quaternion) / (math.pi / 8)).clamp(0.0, 1.0)
        d = 1.0 - (1.0 - dt) * (1.0 - dr)
    elif weight_type == 'depth':
        diff = (depth_reproj - depth_out.unsqueeze(1).expand_as(depth_reproj)).abs()
        return torch.softmax(1.0 / ((diff / diff.max()) ** p + eps), dim=1).squeeze(2)
    else:
        raise ValueError(f'Unknown weight_type {weight_type}')
    return torch.softmax(1.0 / (d ** p).clamp(min=eps), dim=1)


def render_ibr(camera_in, camera_out, image_in, depth_fake_in, depth_fake_out, p=0.5, weight_type='cam_dist', eps=1e-2):
    """image_in [B,V_i,C,H,W], depths [B,V,1,H,W] -> (image_ibr [B,V_o,C,H,W], image_reproj [B,V_o,V_i,C,H,W])."""
    batch = image_in.shape[0]
    nin, nout = camera_in.length // batch, camera_out.length // batch
    ibrs, reprojs = [], []
    for i in range(batch):
        cin, cout = camera_in[i * nin:(i + 1) * nin], camera_out[i * nout:(i + 1) * nout]
        image_reproj, depth_reproj = reproject_views(image_in[i], depth_fake_in[i], depth_fake_out[i], cin, cout)
        weights = _view_weights(cin, cout, weight_type, p, eps, depth_reproj, depth_fake_out[i])
        ibrs.append(ops.ibr_blend(image_reproj, weights))
        reprojs.append(image_reproj)
    return torch.stack(ibrs, 0), torch.stack(reprojs, 0)


def render_latent_ibr(photographer, z_obj, camera_in, camera_out, image_in, p=0.5, weight_type='cam_dist', eps=0.0001):
    fake_in, _, _ = photographer.decode(z_obj, camera_in)
    fake_out, _, _ = photographer.decode(z_obj, camera_out)
    image_ibr, image_reproj = render_ibr(camera_in, camera_out, image_in, fake_in['depth'], fake_out['depth'],
                                         p, weight_type, eps)
    return image_ibr, fake_out['depth'], fake_out['mask'], image_reproj


def render_latent_ibr2(photographer, z_obj, camera_in, camera_out, image_in, p=0.5, weight_type='cam_dist',
                       return_latent=True, eps=0.0001, apply_mask=False):
    y_in, _, _ = photographer.decode(z_obj, camera_in, apply_mask=apply_mask)
    y_out, z_out, _ = photographer.decode(z_obj, camera_out, return_latent=return_latent, apply_mask=apply_mask)
    image_ibr, _ = render_ibr(camera_in, camera_out, image_in, y_in['depth'], y_out['depth'], p, weight_type, eps)
    y_out['color'] = image_ibr * (y_out['mask'] > 0.5) if apply_mask else image_ibr
    return y_out, z_out


def blend_logits(logits, image_reproj):
    """logits [B,V_i,H,W], image_reproj [B,V_i,C,H,W] -> (image [B,C,H,W], weights [B,V_i,1,H,W])."""
    weights = torch.softmax(logits, dim=1)
    return ops.ibr_blend(image_reproj, weights), weights.unsqueeze(2)


def warp_blend_logits(logits, image_reproj, flow_size):
    """logits [B,3*V_i,H,W] -> (image, blend_weights [B,V_i,1,H,W], flow_dx, flow_dy).  One kernel launch."""
    return ops.ibr_warp_blend(logits, image_reproj, flow_size)
