"""Rigid transforms as 4x4 matrices.  API mirror of reference ``latentfusion/three/rigid.py``."""
import torch

from . import quaternion as _q
from .core import ensure_batch_dim, dehomogenize


def intrinsic_to_3x4(matrix):
    m, squeezed = ensure_batch_dim(matrix, 2)
    out = torch.cat((m, m.new_zeros(m.shape[0], 3, 1)), dim=-1)
    return out.squeeze(0) if squeezed else out


def matrix_3x3_to_4x4(matrix):
    m, squeezed = ensure_batch_dim(matrix, 2)
    out = m.new_zeros(m.shape[0], 4, 4)
    out[:, :3, :3] = m
    out[:, 3, 3] = 1.0
    return out.squeeze(0) if squeezed else out


rotation_to_4x4 = matrix_3x3_to_4x4


def translation_to_4x4(translation):
    t, squeezed = ensure_batch_dim(translation, 1)
    out = torch.eye(4, device=t.device, dtype=t.dtype).repeat(t.shape[0], 1, 1)
    out = out + torch.nn.functional.pad(t.unsqueeze(2), (3, 0, 0, 1))
    return out.squeeze(0) if squeezed else out


def decompose(matrix):
    """4x4 -> (rotation 4x4, translation 4x4)."""
    m, squeezed = ensure_batch_dim(matrix, 2)
    R = m.clone()
    R[:, :3, 3] = 0.0
    R[:, 3, :3] = 0.0
    R[:, 3, 3] = 1.0
    T = torch.eye(4, device=m.device, dtype=m.dtype).repeat(m.shape[0], 1, 1)
    T[:, :, 3] = m[:, :, 3]
    if squeezed:
        return R.squeeze(0), T.squeeze(0)
    return R, T


def inverse_transform(matrix):
    m, squeezed = ensure_batch_dim(matrix, 2)
    Rt = m[:, :3, :3].transpose(1, 2)
    out = torch.zeros_like(m)
    out[:, :3, :3] = Rt
    out[:, :3, 3] = -(Rt @ m[:, :3, 3:4]).squeeze(2)
    out[:, 3, 3] = 1
    return out.squeeze(0) if squeezed else out


def translate_matrix(matrix, offset):
    m, squeezed = ensure_batch_dim(matrix, 2)
    inv = inverse_transform(m)
    inv[:, :3, 3] += offset
    out = inverse_transform(inv)
    return out.squeeze(0) if squeezed else out


def scale_matrix(matrix, scale):
    m, squeezed = ensure_batch_dim(matrix, 2)
    inv = inverse_transform(m)
    inv[:, :3, 3] *= scale
    out = inverse_transform(inv)
    return out.squeeze(0) if squeezed else out


def extrinsic_to_position(extrinsic):
    m, squeezed = ensure_batch_dim(extrinsic, 2)
    R, T = decompose(m)
    pos = dehomogenize((R.transpose(2, 1) @ T[:, :, 3, None]).squeeze(-1))
    return pos.squeeze(0) if squeezed else pos


def random_translation(n, x_bound, y_bound, z_bound):
    lo = torch.tensor([x_bound[0], y_bound[0], z_bound[0]])
    hi = torch.tensor([x_bound[1], y_bound[1], z_bound[1]])
    return torch.rand(n, 3) * (hi - lo) + lo


def to_extrinsic_matrix(translation, quaternion):
    return translation_to_4x4(translation) @ rotation_to_4x4(_q.quat_to_mat(quaternion))


def extrinsic_to_quat(extrinsic):
    R, _ = decompose(extrinsic)
    return _q.mat_to_quat(R[..., :3, :3])
