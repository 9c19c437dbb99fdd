"""Building blocks.  API mirror of reference ``latentfusion/modules/__init__.py``."""
import torch
from torch import nn

from .. import ops


class PixelNorm(nn.Module):
    """x / sqrt(mean_c x^2 + 1e-8)  (reference modules/__init__.py:8-15).  On the hot path this never
    runs standalone: Block / the projections fuse it into the producing convolution's epilogue.  The
    standalone module exists for API parity."""

    def forward(self, x):
        return x / torch.sqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


class Interpolate(nn.Module):
    """F.interpolate(scale_factor 2 | 0.5, nearest | (bi|tri)linear, align_corners=False)
    (reference modules/__init__.py:18-36) as one gather kernel."""

    def __init__(self, scale_factor, mode='nearest'):
        super().__init__()
        self.scale_factor = scale_factor
        self.mode = mode
        self.align_corners = False if mode in ('bilinear', 'trilinear') else None

    def forward(self, x):
        return ops.interpolate(x, self.scale_factor, self.mode)

    def extra_repr(self):
        return f"scale_factor={self.scale_factor}"


from .equalized import EqualizedConv2d, EqualizedConv3d  # noqa: E402,F401
