"""Small vector helpers (host-side, differentiable torch ops on tiny tensors).
API mirror of reference ``latentfusion/three/core.py``."""
import torch


def acos_safe(t, eps: float = 1e-7):
    return torch.acos(t.clamp(-1.0 + eps, 1.0 - eps))


def ensure_batch_dim(tensor, num_dims: int):
    if tensor.dim() == num_dims:
        return tensor.unsqueeze(0), True
    return tensor, False


def normalize(vector, dim: int = -1):
    return vector / vector.norm(p=2.0, dim=dim, keepdim=True)


def uniform(n: int, min_val: float, max_val: float):
    return torch.rand(n) * (max_val - min_val) + min_val


def uniform_unit_vector(n):
    return normalize(torch.randn(n, 3), dim=1)


def inner_product(a, b):
    return (a * b).sum(dim=-1)


def homogenize(coords):
    return torch.cat((coords, torch.ones_like(coords[..., :1])), dim=-1)


def dehomogenize(coords):
    return coords[..., :-1] / coords[..., -1:]


def transform_coords(coords, transform):
    coords, squeezed = ensure_batch_dim(coords, 2)
    out = dehomogenize((transform @ homogenize(coords).transpose(1, 2)).transpose(1, 2))
    return out.squeeze(0) if squeezed else out


def grid_to_coords(grid):
    return grid.reshape(grid.shape[0], -1, grid.shape[-1])
