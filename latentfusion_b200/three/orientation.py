"""Viewpoint sampling.  API mirror of the parts of reference ``latentfusion/three/orientation.py``
used on the path (evenly_distributed_points :126-163, evenly_distributed_quats :166-169,
random_quat_from_ray :71-93)."""
import math

import torch

from . import core
from . import quaternion as _q


def random_quat_from_ray(forward, up=None):
    n = forward.shape[0]
    if up is None:
        down = core.uniform_unit_vector(n)
    else:
        down = -(torch.tensor(up).unsqueeze(0).expand(n, 3) + forward)
    right = core.normalize(torch.cross(down, forward, dim=-1))
    down = core.normalize(torch.cross(forward, right, dim=-1))
    return _q.mat_to_quat(torch.stack([right, down, forward], dim=1))


def evenly_distributed_points(n: int, hemisphere=False, pole=(0.0, 0.0, 1.0)):
    """Sunflower (golden-angle) points on the sphere."""
    idx = torch.arange(0, n, dtype=torch.float32) + 0.5
    phi = torch.acos(1 - 2 * idx / n / 2) if hemisphere else torch.acos(1 - 2 * idx / n)
    theta = math.pi * (1 + 5 ** 0.5) * idx
    pts = torch.stack((torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)), dim=1)
    if hemisphere:
        default = torch.tensor([(0.0, 0.0, 1.0)]).expand(n, 3)
        pole_t = torch.tensor([pole]).expand(n, 3)
        if (default[0] + pole_t[0]).abs().sum() < 1e-5:
            pts = -pts
        elif (default[0] - pole_t[0]).abs().sum() >= 1e-5:
            axis = torch.cross(pole_t, default, dim=-1)
            angle = torch.acos(core.inner_product(pole_t, default))
            pts = _q.rotate_vector(_q.from_axis_angle(axis, angle), pts)
    return pts


def evenly_distributed_quats(n: int, hemisphere=False, hemisphere_pole=(0.0, 0.0, 1.0),
                             upright=False, upright_up=(0.0, 0.0, 1.0)):
    rays = evenly_distributed_points(n, hemisphere, hemisphere_pole)
    return random_quat_from_ray(-rays, upright_up if upright else None)
