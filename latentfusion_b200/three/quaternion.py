"""Quaternion algebra, (w, x, y, z) convention.  API mirror of reference
``latentfusion/three/quaternion.py`` (qexp :287-311, qlog :314-334, quat_to_mat :39-93,
mat_to_quat :96-180, qmul :205-227, angular_distance :372-377).  Host-side math on [N,4] tensors."""
import math

import torch
from torch.nn import functional as F

from . import core


def identity(n: int, device='cpu'):
    return torch.tensor((1.0, 0.0, 0.0, 0.0), device=device).repeat(n, 1)


def normalize(quaternion, eps: float = 1e-12):
    if quaternion.shape[-1] != 4:
        raise ValueError(f"Input must be a tensor of shape (*, 4). Got {quaternion.shape}")
    return F.normalize(quaternion, p=2.0, dim=-1, eps=eps)


def quat_to_mat(quaternion):
    q, squeezed = core.ensure_batch_dim(quaternion, 1)
    w, x, y, z = normalize(q).unbind(-1)
    x2, y2, z2 = 2.0 * x, 2.0 * y, 2.0 * z
    m = torch.stack((1.0 - (y2 * y + z2 * z), y2 * x - z2 * w, z2 * x + y2 * w,
                     y2 * x + z2 * w, 1.0 - (x2 * x + z2 * z), z2 * y - x2 * w,
                     z2 * x - y2 * w, z2 * y + x2 * w, 1.0 - (x2 * x + y2 * y)), dim=-1).view(-1, 3, 3)
    return m.squeeze(0) if squeezed else m


def mat_to_quat(rotation_matrix, eps: float = 1e-8):
    """Shepperd's branch selection (largest of trace / diagonal), vectorised with torch.where."""
    m, squeezed = core.ensure_batch_dim(rotation_matrix, 2)
    if m.shape[-2:] != (3, 3):
        raise ValueError(f"Input size must be a (*, 3, 3) tensor. Got {m.shape}")
    tiny = torch.finfo(m.dtype).tiny
    m00, m01, m02 = m[..., 0, 0:1], m[..., 0, 1:2], m[..., 0, 2:3]
    m10, m11, m12 = m[..., 1, 0:1], m[..., 1, 1:2], m[..., 1, 2:3]
    m20, m21, m22 = m[..., 2, 0:1], m[..., 2, 1:2], m[..., 2, 2:3]
    trace = m00 + m11 + m22

    def div(a, b):
        return a / b.clamp(min=tiny)

    s0 = torch.sqrt(trace + 1.0) * 2.0
    q0 = torch.cat((0.25 * s0, div(m21 - m12, s0), div(m02 - m20, s0), div(m10 - m01, s0)), dim=-1)
    s1 = torch.sqrt(1.0 + m00 - m11 - m22 + eps) * 2.0
    q1 = torch.cat((div(m21 - m12, s1), 0.25 * s1, div(m01 + m10, s1), div(m02 + m20, s1)), dim=-1)
    s2 = torch.sqrt(1.0 + m11 - m00 - m22 + eps) * 2.0
    q2 = torch.cat((div(m02 - m20, s2), div(m01 + m10, s2), 0.25 * s2, div(m12 + m21, s2)), dim=-1)
    s3 = torch.sqrt(1.0 + m22 - m00 - m11 + eps) * 2.0
    q3 = torch.cat((div(m10 - m01, s3), div(m02 + m20, s3), div(m12 + m21, s3), 0.25 * s3), dim=-1)
    q = torch.where(trace > 0.0, q0,
                    torch.where((m00 > m11) & (m00 > m22), q1, torch.where(m11 > m22, q2, q3)))
    return q.squeeze(0) if squeezed else q


def random(k: int = 1, device='cpu'):
    u = torch.rand(k, 3, device=device)
    a, b = torch.sqrt(1.0 - u[:, 0]), torch.sqrt(u[:, 0])
    t1, t2 = 2.0 * math.pi * u[:, 1], 2.0 * math.pi * u[:, 2]
    return torch.stack((torch.cos(t2) * b, torch.sin(t1) * a, torch.cos(t1) * a, torch.sin(t2) * b), dim=1)


def qmul(q1, q2):
    """Hamilton product with the reference's operand convention (its outer product is q2 x q1)."""
    assert q1.shape[-1] == 4 and q2.shape[-1] == 4
    a, b = q2.reshape(-1, 4), q1.reshape(-1, 4)
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw - ay * bz + az * by,
                        aw * by + ax * bz + ay * bw - az * bx,
                        aw * bz - ax * by + ay * bx + az * bw), dim=1).view(q1.shape)


def rotate_vector(quat, vector):
    assert quat.shape[-1] == 4 and vector.shape[-1] == 3
    shape = vector.shape
    q, v = quat.reshape(-1, 4), vector.reshape(-1, 3)
    uv = torch.cross(q[:, 1:], v, dim=1)
    uuv = torch.cross(q[:, 1:], uv, dim=1)
    return (v + 2 * (q[:, :1] * uv + uuv)).view(shape)


def from_axis_angle(axis, angle):
    if torch.is_tensor(axis) and isinstance(angle, float):
        angle = torch.full((axis.shape[0],), angle, dtype=axis.dtype, device=axis.device)
    axis = axis / axis.norm(dim=-1, keepdim=True)
    half = angle / 2.0
    s = torch.sin(half)
    return torch.stack((torch.cos(half), s * axis[..., 0], s * axis[..., 1], s * axis[..., 2]), dim=-1)


def qexp(q, eps=1e-8):
    """exp of a quaternion given as (s; v) [*,4] or as its pure part v [*,3]."""
    if q.shape[1] == 4:
        s, v = q[:, :1], q[:, 1:]
    else:
        s, v = torch.zeros_like(q[:, :1]), q
    theta = v.norm(dim=-1, keepdim=True)
    xyz = 1.0 / theta.clamp(min=eps) * torch.sin(theta) * v
    return torch.exp(s) * torch.cat((torch.cos(theta), xyz), dim=-1)


def qlog(q, eps=1e-8):
    mag = q.norm(dim=-1, keepdim=True)
    s, v = q[..., :1], q[..., 1:]
    xyz = v / v.norm(dim=-1, keepdim=True).clamp(min=eps) * core.acos_safe(s / mag.clamp(min=eps))
    return torch.cat((torch.log(mag), xyz), dim=-1)


def qdelta(n, std, device=None):
    omega = torch.cat((torch.zeros(n, 1, device=device), torch.randn(n, 3, device=device)), dim=-1)
    return qexp(std / 2.0 * omega)


def perturb(q, std):
    q, squeezed = core.ensure_batch_dim(q, num_dims=1)
    out = qmul(qdelta(q.shape[0], std, device=q.device), q)
    return out.squeeze(0) if squeezed else out


def angular_distance(q1, q2, eps: float = 1e-7):
    dot = normalize(q1) @ normalize(q2).t()
    return 2 * core.acos_safe(dot.abs(), eps=eps)
