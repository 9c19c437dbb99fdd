"""Initial translation from a depth map + mask.  API mirror of reference
``latentfusion/pose/initialization.py`` (:35-97); the disk erosion is a min-pool here instead of
skimage (host-side, once per target)."""
import torch
from torch.nn import functional as F

from .. import three
from ..modules.geometry import Camera


def _disk(size, device):
    r = torch.arange(-size, size + 1, device=device, dtype=torch.float32)
    yy, xx = torch.meshgrid(r, r, indexing='ij')
    return (yy ** 2 + xx ** 2) <= size ** 2


def _erode_mask(mask, size=5):
    """Binary erosion of a [1,H,W] bool mask with a radius-`size` disk."""
    foot = _disk(size, mask.device)
    inv = (~mask).float().unsqueeze(0)
    hit = F.conv2d(inv, foot.float().view(1, 1, *foot.shape), padding=size) > 0   # any background under the disk
    return (mask & ~hit.squeeze(0))


def _reject_outliers_mad(data, m=2.0):
    median = data.median()
    mad = torch.median(torch.abs(data - median))
    keep = torch.abs(data - median) / mad < m
    return data[keep], int((~keep).sum())


def _mask_boxes(masks, pad=0.0):
    out = []
    for m in masks:
        ys, xs = torch.nonzero(m.squeeze(), as_tuple=True)
        out.append(torch.stack((xs.min() - pad, ys.min() - pad, xs.max() + pad, ys.max() + pad)).float())
    return torch.stack(out, dim=0)


def estimate_translation(depth, mask, intrinsic):
    mask_b = mask.bool()
    zs = torch.zeros(depth.shape[0], device=depth.device)
    for i in range(depth.shape[0]):
        vals = depth[i][_erode_mask(mask_b[i], size=3) & (depth[i] > 0.0)]
        vals, _ = _reject_outliers_mad(vals, m=3.0)
        zs[i] = (vals.min() + vals.max()) / 2.0
    boxes = _mask_boxes(mask)
    cu, cv = (boxes[:, 2] + boxes[:, 0]) / 2.0, (boxes[:, 3] + boxes[:, 1]) / 2.0
    x = (cu - intrinsic[..., 0, 2]) / intrinsic[..., 0, 0] * zs
    y = (cv - intrinsic[..., 1, 2]) / intrinsic[..., 1, 1] * zs
    return x, y, zs


def estimate_initial_pose(depth, mask, intrinsic, width, height) -> Camera:
    translation = torch.stack(estimate_translation(depth, mask, intrinsic), dim=-1)
    rotation = three.quaternion.identity(intrinsic.shape[0], intrinsic.device)
    return Camera(intrinsic, three.to_extrinsic_matrix(translation, rotation), height=height, width=width)
