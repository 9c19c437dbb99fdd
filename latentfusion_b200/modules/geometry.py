"""Camera model, the two voxel resamplers and the 2D<->3D projections.

API mirror of reference ``latentfusion/modules/geometry.py`` (Camera :46-590,
CameraToObjectTransform :614-657, ObjectToCameraTransform :660-690, Tile/Factor projections :693-749).

What differs from the reference is *how* the hot operators run:

* the resamplers never materialise a sampling grid nor N copies of the cube — each camera is reduced to
  a 40-float constant block (``o2c_block`` / ``c2o_block``) and ``liblfb200`` generates sample positions
  in registers (csrc/resample.cu);
* the gradient of object->camera w.r.t. the camera comes back from the kernel as 17 numbers per camera
  (d/d cam_to_obj, viewport, znear) and autograd finishes the chain through the tiny matrix algebra
  below;
* the projections are depth-collapse / depth-expand GEMMs with the LeakyReLU + PixelNorm epilogue fused.

The camera algebra itself (quaternion exp -> rotation -> 4x4 products) stays as differentiable torch
ops on [N,·] tensors: it is host-level plumbing, a few hundred flops per camera.
"""
import abc

import torch
from torch import nn
from torch.nn import functional as F

from .. import ops, three
from .._lib import CAM_STRIDE
from ..three import quaternion as quat
from ..three.batchview import bv2b, b2bv
from . import PixelNorm
from .equalized import EqualizedConv2d


def _sample2d(image, grid, **kwargs):
    # 2-D crops of observations (pre-processing / loss head), not the voxel path.
    return F.grid_sample(image.float(), grid.float(), align_corners=False, **kwargs)


def bbox_to_grid(bbox, in_size, out_size):
    """Sampling grid (x,y in [-1,1]) covering ``bbox`` = (xmin, ymin, xmax, ymax) of an in_size image."""
    return bboxes_to_grid(bbox.view(1, 4), in_size, out_size)[0]


def bboxes_to_grid(boxes, in_size, out_size):
    """Crop grids of reference geometry.py:20-43.  As the reference executes (its TorchScript `bbox[i].item()` becomes
    an implicit int), the box edges are TRUNCATED toward zero before the grid is laid out — the crops of an
    Observation are taken on whole-pixel boxes while the camera keeps the exact viewport (quirks are
    reproduced, not fixed; pinned by tests/golden/facade_s16_c8.npz)."""
    boxes = torch.trunc(boxes.float())
    h, w = float(in_size[0]), float(in_size[1])
    oh, ow = int(out_size[0]), int(out_size[1])

    def rows(lo, hi, steps):
        # torch.linspace(lo_i, hi_i, steps) for every box i, with ATen's symmetric evaluation (start + i*step in the
        # first half, end - (steps-1-i)*step in the second), so the crops match the reference bit for bit
        i = torch.arange(steps, device=boxes.device, dtype=torch.float32)[None]
        step = (hi - lo)[:, None] / max(steps - 1, 1)
        return torch.where(i < steps // 2, lo[:, None] + i * step, hi[:, None] - (steps - 1 - i) * step)

    gx = rows(boxes[:, 0] / w, boxes[:, 2] / w, ow) * 2 - 1
    gy = rows(boxes[:, 1] / h, boxes[:, 3] / h, oh) * 2 - 1
    n = boxes.shape[0]
    return torch.stack((gx[:, None, :].expand(n, oh, ow), gy[:, :, None].expand(n, oh, ow)), dim=-1)


class Camera(nn.Module):
    """A batch of pin-hole cameras: log-quaternion rotation, translation, intrinsics and a viewport
    (crop box in pixels of the full width x height frame).  The near/far planes sit ``z_span`` either
    side of the translation's z."""

    def __init__(self, intrinsic, extrinsic, z_span=0.5, viewport=None, width=640, height=480,
                 log_quaternion=None, translation=None):
        super().__init__()
        if intrinsic.dim() == 2:
            intrinsic = intrinsic.unsqueeze(0)
        if intrinsic.shape[-2:] == (3, 3):
            intrinsic = three.intrinsic_to_3x4(intrinsic)
        n = intrinsic.shape[0]
        if viewport is None:
            viewport = torch.tensor((0, 0, width, height), dtype=torch.float32,
                                    device=intrinsic.device).view(1, 4).expand(n, -1)
        elif viewport.dim() == 1:
            viewport = viewport.unsqueeze(0)
        self.width, self.height, self.z_span = width, height, z_span
        self.register_buffer('viewport', viewport)
        self.register_buffer('intrinsic', intrinsic)

        if extrinsic is not None:
            if extrinsic.dim() == 2:
                extrinsic = extrinsic.unsqueeze(0)
            log_quaternion, translation = self._split_extrinsic(extrinsic)
        if translation is None:
            raise ValueError("translation must be given through extrinsic or explicitly.")
        if log_quaternion is None:
            raise ValueError("log_quaternion must be given through extrinsic or explicitly.")
        if translation.dim() == 1:
            translation = translation.unsqueeze(0)
        if log_quaternion.dim() == 1:
            log_quaternion = log_quaternion.unsqueeze(0)
        self.register_buffer('log_quaternion', log_quaternion)
        self.register_buffer('translation', translation)

    @staticmethod
    def _split_extrinsic(extrinsic):
        q = quat.mat_to_quat(extrinsic[:, :3, :3].contiguous())
        # a unit quaternion's log has zero real part: keep the 3-vector
        return quat.qlog(q)[:, 1:], extrinsic[:, :3, 3].contiguous()

    def _like(self, intrinsic=None, viewport=None, log_quaternion=None, translation=None):
        return Camera(self.intrinsic if intrinsic is None else intrinsic, None, self.z_span,
                      self.viewport if viewport is None else viewport,
                      width=self.width, height=self.height,
                      log_quaternion=self.log_quaternion if log_quaternion is None else log_quaternion,
                      translation=self.translation if translation is None else translation)

    # ---- (de)serialisation used by Observation.save/load and DataParallel scatter in the reference
    def to_kwargs(self):
        return {'intrinsic': self.intrinsic, 'extrinsic': self.extrinsic, 'z_span': self.z_span,
                'viewport': self.viewport, 'height': self.height, 'width': self.width}

    @classmethod
    def from_kwargs(cls, kwargs):
        return cls(**{k: torch.tensor(v, dtype=torch.float32) if isinstance(v, list) else v
                      for k, v in kwargs.items()})

    # ---- rotation / translation algebra
    @property
    def quaternion(self):
        return quat.qexp(self.log_quaternion)

    @quaternion.setter
    def quaternion(self, q):
        self.log_quaternion = quat.qlog(q)[:, 1:]

    @property
    def rotation_matrix(self):
        return three.matrix_3x3_to_4x4(quat.quat_to_mat(quat.normalize(self.quaternion)))

    @property
    def translation_matrix(self):
        return three.translation_to_4x4(self.translation)

    @property
    def inv_translation_matrix(self):
        return three.translation_to_4x4(-self.translation)

    @property
    def extrinsic(self):
        return self.translation_matrix @ self.rotation_matrix

    @extrinsic.setter
    def extrinsic(self, extrinsic):
        lq, t = self._split_extrinsic(extrinsic)
        self.log_quaternion.copy_(lq)
        self.translation.copy_(t)

    @property
    def obj_to_cam(self):
        return self.translation_matrix @ self.rotation_matrix

    @property
    def cam_to_obj(self):
        return self.rotation_matrix.transpose(2, 1) @ self.inv_translation_matrix

    @property
    def obj_to_image(self):
        return self.intrinsic @ self.obj_to_cam

    @property
    def position(self):
        return -(self.rotation_matrix[:, :3, :3].transpose(2, 1) @ self.translation.unsqueeze(-1)).squeeze(-1)

    @property
    def direction(self):
        p = self.position
        return p / p.norm(dim=1, keepdim=True)

    def rotate(self, q):
        self.quaternion = quat.qmul(self.quaternion, q)
        return self

    def translate(self, offset):
        offset, _ = three.ensure_batch_dim(offset, 1)
        pos = three.homogenize(self.position + offset.expand_as(self.position)).unsqueeze(-1)
        self.translation = three.dehomogenize(-(self.rotation_matrix @ pos).squeeze(2))
        return self

    # ---- simple accessors
    @property
    def device(self):
        return self.intrinsic.device

    @property
    def length(self):
        return self.intrinsic.size(0)

    def __len__(self):
        return self.length

    @property
    def viewport_height(self):
        return self.viewport[:, 3] - self.viewport[:, 1]

    @property
    def viewport_width(self):
        return self.viewport[:, 2] - self.viewport[:, 0]

    @property
    def viewport_centroid(self):
        return torch.stack(((self.viewport[:, 2] + self.viewport[:, 0]) / 2.0,
                            (self.viewport[:, 3] + self.viewport[:, 1]) / 2.0), dim=-1)

    @property
    def u0(self):
        return self.intrinsic[:, 0, 2]

    @property
    def v0(self):
        return self.intrinsic[:, 1, 2]

    @property
    def fu(self):
        return self.intrinsic[:, 0, 0]

    @property
    def fv(self):
        return self.intrinsic[:, 1, 1]

    @property
    def fov_u(self):
        return torch.atan2(self.fu, self.viewport_width / 2.0)

    @property
    def fov_v(self):
        return torch.atan2(self.fv, self.viewport_height / 2.0)

    @property
    def znear(self):
        return self.translation[:, 2] - self.z_span

    @property
    def zfar(self):
        return self.translation[:, 2] + self.z_span

    @property
    def z_bounds(self):
        return self.znear, self.zfar

    # ---- constant blocks consumed by csrc/resample.cu (layout: include/lfb200.h)
    def o2c_block(self, cube_size):
        if self.intrinsic.is_cuda:
            # one kernel (+ one for the analytic VJP) instead of ~60 tiny launches and their autograd
            return ops.camera_o2c_block(self.log_quaternion, self.translation, self.viewport, self.intrinsic,
                                        self.z_span, cube_size)
        n, dev = self.length, self.device
        m = self.cam_to_obj[:, :3, :].reshape(n, 12)
        vp = torch.stack((self.viewport[:, 0], self.viewport[:, 1], self.viewport_width,
                          self.viewport_height), dim=1)
        k = torch.stack((self.u0, self.v0, self.fu, self.fv), dim=1)
        tail = torch.zeros(n, CAM_STRIDE - 23, device=dev)
        # (torch.full, not torch.tensor(list): no host->device copy, so the call is CUDA-graph capturable)
        const = torch.cat((torch.full((n, 1), float(self.z_span), device=dev),
                           torch.full((n, 1), cube_size / 2.0, device=dev)), dim=1)
        return torch.cat((m, vp, k, self.znear.unsqueeze(1), const, tail), dim=1)

    def c2o_block(self, cube_size):
        n, dev = self.length, self.device
        m = self.obj_to_cam[:, :3, :].reshape(n, 12)
        vp = torch.stack((self.viewport[:, 0], self.viewport[:, 1], self.viewport_width,
                          self.viewport_height), dim=1)
        k = self.intrinsic.reshape(n, 12)
        tail = torch.zeros(n, CAM_STRIDE - 31, device=dev)
        cube = torch.full((n, 1), float(cube_size), device=dev)
        return torch.cat((m, vp, k, self.znear.unsqueeze(1), self.zfar.unsqueeze(1), cube, tail), dim=1)

    # ---- image <-> viewport crops (2-D; observation pre-processing and the loss head)
    def ibr_block(self, eps=0.01):
        """[n, 48] constant block of the IBR kernels (layout in include/lfb200.h): cam_to_obj, obj_to_cam,
        obj_to_image (3x4 each), viewport origin/size, principal point/focal, depth normalisation bounds."""
        n, dev = self.length, self.device
        vp = torch.stack((self.viewport[:, 0], self.viewport[:, 1], self.viewport_width, self.viewport_height), dim=1)
        k = torch.stack((self.u0, self.v0, self.fu, self.fv), dim=1)
        zb = torch.stack((self.znear - eps, self.zfar + eps), dim=1)
        return torch.cat((self.cam_to_obj[:, :3, :].reshape(n, 12), self.obj_to_cam[:, :3, :].reshape(n, 12),
                          self.obj_to_image.reshape(n, 12), vp, k, zb, torch.zeros(n, 2, device=dev)), dim=1).float()

    def _full_viewport(self):
        z = torch.zeros(self.length, 1, device=self.device)
        return torch.cat((z, z, z + float(self.width), z + float(self.height)), dim=1)

    def uncrop(self, image=None, scale_mode='nearest', scale=1.0):
        full = self._like(viewport=self._full_viewport())
        if image is None:
            return full
        w, h = int(self.width * scale), int(self.height * scale)
        vp = self.viewport * scale
        ys = torch.arange(0, h, device=self.device, dtype=torch.float32)
        xs = torch.arange(0, w, device=self.device, dtype=torch.float32)
        gy = (ys[None, :] - vp[:, 1, None]) / (self.viewport_height * scale)[:, None] * 2 - 1
        gx = (xs[None, :] - vp[:, 0, None]) / (self.viewport_width * scale)[:, None] * 2 - 1
        n = image.shape[0]
        grid = torch.stack((gx[:, None, :].expand(n, h, w), gy[:, :, None].expand(n, h, w)), dim=-1)
        return _sample2d(image, grid, mode=scale_mode, padding_mode='border'), full

    def crop_to_viewport(self, image, target_size, scale_mode='nearest'):
        grid = bboxes_to_grid(self.viewport, (self.height, self.width), (target_size, target_size))
        return _sample2d(image, grid, mode=scale_mode)

    def zoom(self, image, target_size, target_dist, target_fu=None, target_fv=None, image_scale=1.0,
             zs=None, centroid_uvs=None, scale_mode='bilinear'):
        """Re-frame as if the object were seen from ``target_dist`` into a ``target_size`` crop: returns
        the camera whose viewport is the corresponding box around the projected object origin (and the
        cropped image when one is given)."""
        if zs is None:
            zs = self.translation[:, 2]
        fu, fv = self.fu, self.fv
        target_fu = fu if target_fu is None else target_fu
        target_fv = fv if target_fv is None else target_fv
        box_u = target_dist * (1.0 / zs) / fu * target_fu * target_size / self.width * image_scale
        box_v = target_dist * (1.0 / zs) / fv * target_fv * target_size / self.height * image_scale
        if centroid_uvs is None:
            origin = torch.tensor((0, 0, 0, 1.0), device=self.device).view(1, 4, 1).expand(self.length, -1, -1)
            uvw = self.intrinsic @ self.obj_to_cam @ origin
            centroid_uvs = (uvw[:, :2] / uvw[:, 2, None]).squeeze(-1).clone().float()
        cu, cv = centroid_uvs[:, 0] / self.width, centroid_uvs[:, 1] / self.height
        boxes = torch.stack(((cu - box_u / 2) * float(self.width), (cv - box_v / 2) * float(self.height),
                             (cu + box_u / 2) * float(self.width), (cv + box_v / 2) * float(self.height)), dim=-1)
        zoomed = self._like(viewport=boxes)
        if image is None:
            return zoomed
        grid = bboxes_to_grid(boxes, (self.height, self.width), (target_size, target_size))
        return _sample2d(image, grid, mode=scale_mode), zoomed

    # ---- batching helpers
    def __getitem__(self, item):
        return self._like(self.intrinsic[item], self.viewport[item], self.log_quaternion[item],
                          self.translation[item])

    def __setitem__(self, item, value):
        self.intrinsic[item] = value.intrinsic
        self.viewport[item] = value.viewport
        self.log_quaternion[item] = value.log_quaternion
        self.translation[item] = value.translation

    def __iter__(self):
        return iter([self[i] for i in range(len(self))])

    def split(self, sections):
        parts = zip(torch.split(self.intrinsic, sections), torch.split(self.viewport, sections),
                    torch.split(self.log_quaternion, sections), torch.split(self.translation, sections))
        return [self._like(k, vp, lq, t) for k, vp, lq, t in parts]

    @classmethod
    def cat(cls, cameras):
        first = cameras[0]
        return cls(torch.cat([c.intrinsic for c in cameras], dim=0), None, first.z_span,
                   torch.cat([c.viewport for c in cameras], dim=0),
                   log_quaternion=torch.cat([c.log_quaternion for c in cameras], dim=0),
                   translation=torch.cat([c.translation for c in cameras], dim=0),
                   width=first.width, height=first.height)

    @classmethod
    def vcat(cls, cameras, batch_size=-1):
        first = cameras[0]

        def join(name):
            return bv2b(torch.cat([b2bv(getattr(c, name), batch_size=batch_size) for c in cameras], dim=1))
        return cls(join('intrinsic'), None, first.z_span, join('viewport'),
                   log_quaternion=join('log_quaternion'), translation=join('translation'),
                   width=first.width, height=first.height)

    def repeat(self, n):
        return self._like(self.intrinsic.repeat(n, 1, 1), self.viewport.repeat(n, 1),
                          self.log_quaternion.repeat(n, 1), self.translation.repeat(n, 1))

    def repeat_interleave(self, n):
        r = torch.repeat_interleave
        return self._like(r(self.intrinsic, n, dim=0), r(self.viewport, n, dim=0),
                          r(self.log_quaternion, n, dim=0), r(self.translation, n, dim=0))

    def clone(self):
        return self._like(self.intrinsic.clone(), self.viewport.clone(), self.log_quaternion.clone(),
                          self.translation.clone())

    def detach(self):
        return self._like(self.intrinsic.detach(), self.viewport.detach(), self.log_quaternion.detach(),
                          self.translation.detach())

    # ---- coordinate grids (API parity; the resamplers do NOT use these — they generate them in-kernel)
    def pixel_coords_uvz(self, out_size):
        if isinstance(out_size, int):
            out_size = (out_size,) * 3
        n = self.length
        tz, tv, tu = torch.meshgrid([torch.linspace(0.0, 1.0, s, device=self.device) for s in out_size],
                                    indexing='ij')
        u = tu[None] * self.viewport_width.view(n, 1, 1, 1) + self.viewport[:, 0].view(n, 1, 1, 1)
        v = tv[None] * self.viewport_height.view(n, 1, 1, 1) + self.viewport[:, 1].view(n, 1, 1, 1)
        z = tz[None] * self.z_span + self.znear.view(n, 1, 1, 1)
        return u, v, z

    def pixel_coords_uv(self, out_size):
        if isinstance(out_size, int):
            out_size = (out_size,) * 2
        n = self.length
        tv, tu = torch.meshgrid([torch.linspace(0.0, 1.0, s, device=self.device) for s in out_size],
                                indexing='ij')
        u = tu[None] * self.viewport_width.view(n, 1, 1) + self.viewport[:, 0].view(n, 1, 1)
        v = tv[None] * self.viewport_height.view(n, 1, 1) + self.viewport[:, 1].view(n, 1, 1)
        return u, v

    def camera_coords(self, out_size):
        u, v, z = self.pixel_coords_uvz(out_size)
        s = (-1, 1, 1, 1)
        return (u - self.u0.view(s)) / self.fu.view(s) * z, (v - self.v0.view(s)) / self.fv.view(s) * z, z

    def depth_camera_coords(self, depth):
        u, v = self.pixel_coords_uv((depth.shape[-2], depth.shape[-1]))
        z = depth.view_as(u)
        s = (-1, 1, 1)
        return (u - self.u0.view(s)) / self.fu.view(s) * z, (v - self.v0.view(s)) / self.fv.view(s) * z, z

    def depth_object_coords(self, depth):
        grid = torch.stack(self.depth_camera_coords(depth), dim=-1)
        obj = three.transform_coords(three.grid_to_coords(grid), self.cam_to_obj).view_as(grid)
        return obj[..., 0], obj[..., 1], obj[..., 2]

    def denormalize_depth(self, depth, eps=0.01):
        shape = (*depth.shape[:-3], 1, 1, 1)
        lo, hi = (self.znear - eps).view(shape), (self.zfar + eps).view(shape)
        return (depth / 2.0 + 0.5) * (hi - lo) + lo

    def normalize_depth(self, depth, eps=0.01):
        lo, hi = (self.znear - eps).view(-1, 1, 1, 1), (self.zfar + eps).view(-1, 1, 1, 1)
        return ((depth - lo) / (hi - lo)).clamp(0, 1) * 2.0 - 1.0

    def __repr__(self):
        return f"Camera(count={self.intrinsic.size(0)})"


class BaseTransformBlock(abc.ABC, nn.Module):

    def __init__(self, cube_size):
        super().__init__()
        self.cube_size = cube_size

    def get_obj_coords(self, size, device=None):
        lin = torch.linspace(-self.cube_size / 2, self.cube_size / 2, size, device=device)
        z, y, x = torch.meshgrid(lin, lin, lin, indexing='ij')
        return torch.stack((x, y, z, torch.ones_like(x)), dim=-1).view(-1, 4)


class CameraToObjectTransform(BaseTransformBlock):
    """Camera-frustum volume [V,C,S,S,S] -> object cube [V,C,S,S,S] (K2, ``lf_resample_c2o_*``)."""

    def __init__(self, cube_size, padding_mode='border'):
        super().__init__(cube_size)
        if padding_mode != 'border':
            raise ValueError("only padding_mode='border' is implemented (all the path uses)")
        self.padding_mode = padding_mode

    def forward(self, cam_volume, camera: Camera):
        return ops.resample_c2o(cam_volume, camera.c2o_block(self.cube_size))


class ObjectToCameraTransform(BaseTransformBlock):
    """Object cube [B,C,S,S,S] -> per-camera frustum volumes [N,C,S,S,S] (K1, ``lf_resample_o2c_*``).
    ``B`` may be 1 (or any divisor of N): the cube is shared, never replicated."""

    def __init__(self, cube_size, padding_mode='border'):
        super().__init__(cube_size)
        if padding_mode != 'border':
            raise ValueError("only padding_mode='border' is implemented (all the path uses)")
        self.padding_mode = padding_mode

    def forward(self, obj_volume, camera: Camera, split_only=False):
        """split_only (internal, see ops.resample_o2c): the caller's single reader stages the split-planar form."""
        if camera.length % obj_volume.shape[0] != 0:
            raise ValueError(f"number of cameras ({camera.length}) must be a multiple of the number of "
                             f"object volumes ({obj_volume.shape[0]})")
        return ops.resample_o2c(obj_volume, camera.o2c_block(self.cube_size), split_only=split_only)


class _Projection(nn.Module):
    def __init__(self, conv_in, conv_out, relu_slope, norm_module):
        super().__init__()
        self.conv = EqualizedConv2d(conv_in, conv_out, kernel_size=1, padding=0)
        self.activation = nn.LeakyReLU(relu_slope)
        self.norm = norm_module()
        if not isinstance(self.norm, PixelNorm):
            raise ValueError("projections fuse PixelNorm; other norms are not implemented")


class TileProjection2d3d(_Projection):
    """1x1 conv + LeakyReLU + PixelNorm, then tile along depth (zero-copy expand)."""

    def __init__(self, in_channels, out_channels, out_size, relu_slope=0.2, norm_module=PixelNorm):
        super().__init__(in_channels, out_channels, relu_slope, norm_module)
        self.out_size, self.out_channels = out_size, out_channels

    def forward(self, x):
        x = self.conv(x, act=True, slope=self.activation.negative_slope, norm=True)
        return x.unsqueeze(2).expand(-1, -1, self.out_size, -1, -1)


class FactorProjection2d3d(_Projection):
    """1x1 conv to C*S channels (+LeakyReLU+PixelNorm over all C*S) viewed as [C,S]: run as a
    depth-expand GEMM writing the channels-last volume directly."""

    def __init__(self, in_channels, out_channels, out_size, relu_slope=0.2, norm_module=PixelNorm):
        super().__init__(in_channels, out_channels * out_size, relu_slope, norm_module)
        self.out_size, self.in_channels, self.out_channels = out_size, in_channels, out_channels

    def forward(self, x):
        return ops.eq_conv(x, self.conv.module.weight, self.conv.bias, act=True,
                           slope=self.activation.negative_slope, norm=True, kind=ops.KIND_EXPAND,
                           depth=self.out_size, precision=self.conv.precision)


class FactorProjection3d2d(_Projection):
    """[N,C,S,H,W] -> view [N,C*S,H,W] -> 1x1 conv + LeakyReLU + PixelNorm: a depth-collapse GEMM
    (K = C*S) reading the volume once."""

    def __init__(self, in_channels, out_channels, out_size, relu_slope=0.2, norm_module=PixelNorm):
        super().__init__(in_channels * out_size, out_channels, relu_slope, norm_module)
        self.out_size, self.in_channels, self.out_channels = out_size, in_channels, out_channels

    def forward(self, x):
        return ops.eq_conv(x, self.conv.module.weight, self.conv.bias, act=True,
                           slope=self.activation.negative_slope, norm=True, kind=ops.KIND_COLLAPSE,
                           depth=self.out_size, precision=self.conv.precision)
