"""Helpers around the networks.  API mirror of reference ``latentfusion/recon/utils.py``
(optimal_camera_dist :13-22, voxel/pixel coordinate channels :35-61)."""
import math

import torch


def optimal_camera_dist(focal_length, size, radius, slack=1.5):
    """Distance at which a sphere of `radius` just fills a `size`-pixel crop (+ slack)."""
    theta = math.atan2(size / 2.0, focal_length)          # half field of view
    x = radius * math.cos(theta) / math.sin(theta)
    d = math.sqrt(x ** 2 + radius ** 2 - 2 * x * radius * math.cos(math.pi / 2.0 - theta))
    return d + slack


def _expand_like(coords, ref, spatial_dims):
    lead = ref.dim() - spatial_dims - 1
    return coords.view(*([1] * lead), *coords.shape).expand(*ref.shape[:lead], -1, *ref.shape[-spatial_dims:])


def get_normalized_voxel_coords(tensor):
    """[..., 3, D, H, W] channels (z, y, x), each linspace(-1, 1)."""
    d, h, w = tensor.shape[-3:]
    dev = tensor.device
    z, y, x = torch.meshgrid(torch.linspace(-1.0, 1.0, d, device=dev), torch.linspace(-1.0, 1.0, h, device=dev),
                             torch.linspace(-1.0, 1.0, w, device=dev), indexing='ij')
    return _expand_like(torch.stack((z, y, x), dim=0), tensor, 3)


def get_normalized_pixel_coords(tensor):
    h, w = tensor.shape[-2:]
    dev = tensor.device
    y, x = torch.meshgrid(torch.linspace(-1.0, 1.0, h, device=dev), torch.linspace(-1.0, 1.0, w, device=dev),
                          indexing='ij')
    return _expand_like(torch.stack((y, x), dim=0), tensor, 2)


def get_normalized_voxel_depth(tensor):
    b, _, d, h, w = tensor.shape
    return torch.linspace(-1.0, 1.0, d, device=tensor.device).view(1, 1, d, 1, 1).expand(b, 1, d, h, w)


def mask_normalized_depth(depth, mask):
    return ((depth / 2.0 + 0.5) * mask) * 2.0 - 1.0
