"""Camera sampling / parameterisation and masked loss reductions.  API mirror of the parts of
reference ``latentfusion/pose/utils.py`` used by the estimators (perturb_camera :12-16,
sample_cameras_with_estimate :28-45, parameterize/deparameterize_camera :48-71, flip_camera :74-78,
zero_invalid_pixels :81-96, iou_loss :99-108, reduce_loss_mask :111-117)."""
import math

import torch
from torch import nn

from .. import three
from ..modules.geometry import Camera


def perturb_camera(camera, translation_std, quaternion_std):
    camera = camera.clone()
    camera.translation.data += torch.randn_like(camera.translation) * translation_std
    camera.log_quaternion.data += torch.randn_like(camera.log_quaternion) * quaternion_std
    return camera


def sample_cameras_with_estimate(n, camera_est, translation_std=0.0, hemisphere=False, upright=False) -> Camera:
    device = camera_est.device
    translation = camera_est.translation.expand(n, -1)
    translation = translation + torch.randn_like(translation) * translation_std
    quaternion = three.orientation.evenly_distributed_quats(n, hemisphere=hemisphere, upright=upright)
    extrinsic = three.to_extrinsic_matrix(translation.cpu(), quaternion).to(device)
    return Camera(camera_est.intrinsic.expand(n, -1, -1), extrinsic, camera_est.z_span,
                  width=camera_est.width, height=camera_est.height,
                  viewport=camera_est.viewport.expand(n, -1))


def parameterize_camera(camera, optimize_rotation=True, optimize_translation=True, optimize_viewport=False):
    out = camera.clone()
    if optimize_rotation:
        out.log_quaternion = nn.Parameter(out.log_quaternion)
    if optimize_translation:
        out.translation = nn.Parameter(out.translation)
    if optimize_viewport:
        out.viewport = nn.Parameter(out.viewport)
    return out


def deparameterize_camera(camera):
    out = camera.clone()
    out.log_quaternion = out.log_quaternion.detach()
    out.translation = out.translation.detach()
    out.viewport = out.viewport.detach()
    return out


def flip_camera(camera, axis=(0.0, 0.0, 1.0)):
    ax = torch.tensor([axis], dtype=torch.float32, device=camera.device).expand(len(camera), -1)
    return camera.clone().rotate(three.quaternion.from_axis_angle(ax, math.pi))


def zero_invalid_pixels(tensor, invalid_mask):
    """Drop pixels that have a positive mask but no depth reading (sensor holes)."""
    return tensor * (~invalid_mask).float()


def iou_loss(input_mask, target_mask, eps=1e-4):
    inter = torch.sum(input_mask * target_mask, dim=(1, 2, 3))
    union = torch.sum(input_mask, dim=(1, 2, 3)) + torch.sum(target_mask, dim=(1, 2, 3)) - inter
    return torch.log(union.clamp(min=eps)) - torch.log(inter.clamp(min=eps))


def reduce_loss_mask(loss, mask, eps=1e-4):
    if loss.dim() == 4:
        loss = loss.squeeze(1)
    if mask.dim() == 4:
        mask = mask.squeeze(1)
    return (loss * mask).sum(dim=(-2, -1)).clamp(min=eps / 10) / mask.sum(dim=(-2, -1)).clamp(min=eps)
