"""TEST INFRASTRUCTURE — the CPU oracle for the LatentFusion reconstruct->render hot path.

This file is a *restatement* (plain PyTorch fp32 ops, functional style, parameters read from a
reference-format ``state_dict``) of what the reference computes on the path named by
BASELINE.json.  It is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  The product package (``latentfusion_b200``) never imports anything under ``oracle/``.

Pinning: the reference ships no tests and no golden vectors, so the oracle is
pinned against *outputs of the unmodified reference itself*, generated in the authoring container
by ``oracle/make_golden.py`` and committed under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every function here against those vectors.  The trilinear sampler / convolution arithmetic
that the reference delegates to ATen (``F.grid_sample``, ``F.conv{2,3}d``, ``F.interpolate`` —
torch is a third-party dependency, not under /root/reference, nominally "nightly ~1.6",
README.md:35-39) is additionally restated in plain C in ``oracle/resample_ref.c`` following the
published ``ATen/native/GridSampler.h`` semantics and cross-checked in the same test.

Every function cites the reference file:line it follows (paths relative to
/root/reference/latentfusion/).
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# L0: quaternion / rigid math                      three/quaternion.py, modules/geometry.py
# --------------------------------------------------------------------------------------


def qexp3(v, eps=1e-8):
    """exp of a pure quaternion (0; v) -> (w, x, y, z).  three/quaternion.py:287-311."""
    theta = v.norm(dim=-1, keepdim=True)
    return torch.cat((torch.cos(theta), 1.0 / theta.clamp(min=eps) * torch.sin(theta) * v), dim=-1)


def quat_to_rot(q):
    """(w,x,y,z) -> 3x3.  geometry.py:147-153 normalises (eps 1e-12) then quaternion.py:39-93
    normalises again before expanding the products."""
    q = F.normalize(F.normalize(q, dim=-1, eps=1e-12), dim=-1, eps=1e-12)
    w, x, y, z = q.unbind(-1)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    rows = [1.0 - (ty * y + tz * z), ty * x - tz * w, tz * x + ty * w,
            ty * x + tz * w, 1.0 - (tx * x + tz * z), tz * y - tx * w,
            tz * x - ty * w, tz * y + tx * w, 1.0 - (tx * x + ty * y)]
    return torch.stack(rows, dim=-1).view(-1, 3, 3)


class Cam:
    """Plain record of the ten learnable floats + intrinsics.  geometry.py:46-105."""

    def __init__(self, intrinsic, log_quaternion, translation, viewport, z_span=0.5,
                 width=640, height=480):
        self.intrinsic = intrinsic            # [n,3,4]
        self.log_quaternion = log_quaternion  # [n,3]
        self.translation = translation        # [n,3]
        self.viewport = viewport              # [n,4] xmin ymin xmax ymax
        self.z_span, self.width, self.height = z_span, width, height

    def __len__(self):
        return self.intrinsic.shape[0]

    # geometry.py:147-163, :207-213
    def rot4(self):
        R = quat_to_rot(qexp3(self.log_quaternion))
        out = torch.zeros(len(self), 4, 4, dtype=R.dtype)
        out[:, :3, :3] = R
        out[:, 3, 3] = 1.0
        return out

    def trans4(self, sign=1.0):
        out = torch.eye(4).repeat(len(self), 1, 1)
        out[:, :3, 3] = sign * self.translation
        return out

    def obj_to_cam(self):
        return self.trans4() @ self.rot4()

    def cam_to_obj(self):
        return self.rot4().transpose(1, 2) @ self.trans4(-1.0)

    # geometry.py:249-255
    def znear(self):
        return self.translation[:, 2] - self.z_span

    def zfar(self):
        return self.translation[:, 2] + self.z_span

    def vp_w(self):
        return self.viewport[:, 2] - self.viewport[:, 0]

    def vp_h(self):
        return self.viewport[:, 3] - self.viewport[:, 1]

    def with_viewport(self, viewport):
        return Cam(self.intrinsic, self.log_quaternion, self.translation, viewport,
                   self.z_span, self.width, self.height)

    def zoom(self, target_size, target_dist):
        """Viewport-only zoom (image=None).  geometry.py:294-347."""
        K = self.intrinsic
        zs = self.translation[:, 2]
        fu, fv = K[:, 0, 0], K[:, 1, 1]
        bbox_u = target_dist * (1.0 / zs) / fu * fu * target_size / self.width * 1.0
        bbox_v = target_dist * (1.0 / zs) / fv * fv * target_size / self.height * 1.0
        origin = torch.tensor((0, 0, 0, 1.0)).view(1, 4, 1).expand(len(self), -1, -1)
        uvs = K @ self.obj_to_cam() @ origin
        uvs = (uvs[:, :2] / uvs[:, 2, None]).squeeze(-1)
        cu, cv = uvs[:, 0] / self.width, uvs[:, 1] / self.height
        boxes = torch.stack(((cu - bbox_u / 2) * float(self.width),
                             (cv - bbox_v / 2) * float(self.height),
                             (cu + bbox_u / 2) * float(self.width),
                             (cv + bbox_v / 2) * float(self.height)), dim=-1)
        return self.with_viewport(boxes)

    def denormalize_depth(self, depth, eps=0.01):
        """geometry.py:555-558."""
        shape = (*depth.shape[:-3], 1, 1, 1)
        zn = (self.znear() - eps).view(shape)
        zf = (self.zfar() + eps).view(shape)
        return (depth / 2.0 + 0.5) * (zf - zn) + zn

    def uncrop(self, image, mode):
        """Paste a viewport crop back into the full frame.  geometry.py:261-285."""
        yy, xx = torch.meshgrid(torch.arange(0, self.height, dtype=torch.float32),
                                torch.arange(0, self.width, dtype=torch.float32), indexing='ij')
        vp = self.viewport
        yy = (yy[None] - vp[:, 1, None, None]) / self.vp_h()[:, None, None] * 2 - 1
        xx = (xx[None] - vp[:, 0, None, None]) / self.vp_w()[:, None, None] * 2 - 1
        grid = torch.stack((xx, yy), dim=-1)
        return F.grid_sample(image.float(), grid.float(), mode=mode, padding_mode='border',
                             align_corners=False)


def crop_boxes(image, boxes, in_hw, out_size, mode):
    """Camera.zoom's image branch: bbox -> grid -> grid_sample (zeros padding).
    geometry.py:20-43, :349-352."""
    h, w = in_hw
    n = boxes.shape[0]
    boxes = torch.trunc(boxes)      # geometry.py:24-27: `.item()` under torch.jit.script is an implicit int (truncation)
    lin = torch.linspace(0.0, 1.0, out_size)
    gx = (boxes[:, 0, None] / w + lin[None] * ((boxes[:, 2, None] - boxes[:, 0, None]) / w)) * 2 - 1
    gy = (boxes[:, 1, None] / h + lin[None] * ((boxes[:, 3, None] - boxes[:, 1, None]) / h)) * 2 - 1
    grid = torch.stack((gx[:, None, :].expand(n, out_size, out_size),
                        gy[:, :, None].expand(n, out_size, out_size)), dim=-1)
    return F.grid_sample(image.float(), grid, mode=mode, align_corners=False)


# --------------------------------------------------------------------------------------
# L1: the two voxel resamplers                               modules/geometry.py:599-690
# --------------------------------------------------------------------------------------


def o2c_grid(cam, size, cube_size=1.0):
    """Sampling grid of ObjectToCameraTransform.  geometry.py:469-493, :515-531, :669-685.
    Frustum depths are znear + t*z_span (t in [0,1]) — only the near half."""
    n = len(cam)
    lin = torch.linspace(0.0, 1.0, size)
    zp, vp, up = torch.meshgrid(lin, lin, lin, indexing='ij')
    u = up[None] * cam.vp_w().view(n, 1, 1, 1) + cam.viewport[:, 0].view(n, 1, 1, 1)
    v = vp[None] * cam.vp_h().view(n, 1, 1, 1) + cam.viewport[:, 1].view(n, 1, 1, 1)
    z = zp[None] * cam.z_span + cam.znear().view(n, 1, 1, 1)
    K = cam.intrinsic
    u0, v0 = K[:, 0, 2].view(n, 1, 1, 1), K[:, 1, 2].view(n, 1, 1, 1)
    fu, fv = K[:, 0, 0].view(n, 1, 1, 1), K[:, 1, 1].view(n, 1, 1, 1)
    y = (v - v0) / fv * z
    x = (u - u0) / fu * z
    pts = torch.stack((x, y, z, torch.ones_like(x)), dim=-1).view(n, -1, 4)
    obj = (cam.cam_to_obj() @ pts.transpose(2, 1))[:, :3].transpose(1, 2)
    return (obj / (cube_size / 2)).view(n, size, size, size, 3)


def c2o_grid(cam, size, cube_size=1.0):
    """Sampling grid of CameraToObjectTransform.  geometry.py:599-611, :625-654.
    The z coordinate is (z - znear)/(zfar - znear) in [0,1] with *no* 2x-1."""
    n = len(cam)
    lin = torch.linspace(-cube_size / 2, cube_size / 2, size)
    zc, yc, xc = torch.meshgrid(lin, lin, lin, indexing='ij')
    pts = torch.stack((xc, yc, zc, torch.ones_like(xc)), dim=-1).view(-1, 4)
    pts = pts.t()[None].expand(n, -1, -1)
    pix = cam.intrinsic @ (cam.obj_to_cam() @ pts)
    px = pix[:, 0] / pix[:, 2]
    py = pix[:, 1] / pix[:, 2]
    zn, zf = cam.znear().view(n, 1), cam.zfar().view(n, 1)
    g = torch.stack((((px - cam.viewport[:, 0, None]) / cam.vp_w()[:, None]) * 2 - 1,
                     ((py - cam.viewport[:, 1, None]) / cam.vp_h()[:, None]) * 2 - 1,
                     (pix[:, 2] - zn) / (zf - zn)), dim=-1)
    return g.view(n, size, size, size, 3)


def resample(vol, grid):
    """geometry.py:16-17 — always fp32, trilinear, border padding, align_corners=False."""
    return F.grid_sample(vol.float(), grid.float(), padding_mode='border', align_corners=False)


def object_to_camera(obj_vol, cam, cube_size=1.0):
    """geometry.py:669-690 (expands a single cube to all cameras)."""
    size = obj_vol.shape[-1]
    return resample(obj_vol.expand(len(cam), -1, -1, -1, -1), o2c_grid(cam, size, cube_size))


def camera_to_object(cam_vol, cam, cube_size=1.0):
    """geometry.py:625-657."""
    return resample(cam_vol, c2o_grid(cam, cam_vol.shape[-1], cube_size))


# --------------------------------------------------------------------------------------
# L1: equalised conv blocks             modules/equalized.py, blocks.py, __init__.py, unet.py
# --------------------------------------------------------------------------------------


def eq_conv(x, sd, prefix, padding=0):
    """conv (no bias) * sqrt(2/fan_in) + bias.  equalized.py:57-64, :66-75."""
    w = sd[prefix + '.module.weight']
    he = math.sqrt(2.0 / (w[0].numel()))
    conv = F.conv3d if w.dim() == 5 else F.conv2d
    y = conv(x, w, None, padding=padding) * he
    return y + sd[prefix + '.bias'].view(1, -1, *([1] * (w.dim() - 2)))


def pixel_norm(x):
    """modules/__init__.py:14-15."""
    return x / torch.sqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


def conv_block(x, sd, prefix, scale=1.0, mode='nearest', slope=0.2):
    """(conv -> lrelu -> pixelnorm) x2 -> optional interpolate.  blocks.py:152-164."""
    for name in ('conv1', 'conv2'):
        k = sd[f'{prefix}.{name}.module.weight'].shape[-1]
        x = pixel_norm(F.leaky_relu(eq_conv(x, sd, f'{prefix}.{name}', padding=k // 2), slope))
    if scale != 1.0:
        if mode == 'bilinear' and x.dim() == 5:
            mode = 'trilinear'                                   # blocks.py:34-35
        ac = False if mode in ('bilinear', 'trilinear') else None
        x = F.interpolate(x, scale_factor=scale, mode=mode, align_corners=ac)
    return x


def block_plan(config, skip_connections=False, skip_start=1, skip_end=None):
    """Token walk of create_blocks: returns [(c_in, c_out, scale)] per Block.  blocks.py:10-75
    (in_views = 1 only, which is all the path uses)."""
    n_blocks = sum(1 for t in config if isinstance(t, int)) - 1
    skip_end = n_blocks if skip_end is None else min(n_blocks, skip_end)
    plan, scale, c_in, k = [], 1.0, config[0], 0
    for tok in config[1:]:
        if isinstance(tok, int):
            extra = c_in if (skip_connections and skip_start <= k < skip_end) else 0
            plan.append((c_in + extra, tok, scale))
            c_in, k, scale = tok, k + 1, 1.0
        else:
            scale = {'U': 2.0, 'D': 0.5}[tok]
    return plan


def unet2d(x, sd, prefix, config, has_input_block, heads=0):
    """BaseUNet.forward.  unet.py:95-127.  Block scale mode is always bilinear here because
    BaseUNet never forwards scale_mode (unet.py:24-28)."""
    down, up = config
    if has_input_block:
        x = F.leaky_relu(eq_conv(x, sd, f'{prefix}.input_block.conv'), 0.2)   # blocks.py:78-90
    mids = []
    for i, (_, _, s) in enumerate(block_plan(down)):
        x = conv_block(x, sd, f'{prefix}.down_blocks.{i}', s, 'bilinear')
        mids.insert(0, x)
    n_down = len(mids)
    up_plan = block_plan(up, True, 1, min(n_down, sum(1 for t in up if isinstance(t, int)) - 1))
    for i, (_, _, s) in enumerate(up_plan):
        if 1 <= i < len(mids):
            x = torch.cat((x, mids[i]), dim=1)
        x = conv_block(x, sd, f'{prefix}.up_blocks.{i}', s, 'bilinear')
    return x


# --------------------------------------------------------------------------------------
# L2: Sculptor / fusers / Photographer                     recon/models.py, recon/fusion.py
# --------------------------------------------------------------------------------------


def sculptor_forward(sd, arch, x, cam):
    """Sculptor.forward, eval mode (no autocast).  models.py:198-224.
    arch: dict(image_config, camera_config, object_config, projection_type, scale_mode,
    cube_size)."""
    z = unet2d(x, sd, 'image_encoder', arch['image_config'], True)
    # projection 2D -> 3D: geometry.py:704-708 (tile) / :724-728 (factor)
    z = pixel_norm(F.leaky_relu(eq_conv(z, sd, 'projection_block.conv'), 0.2))
    c0 = arch['camera_config'][0]
    if arch['projection_type'] == 'tile':
        z = z.unsqueeze(2).expand(-1, -1, z.shape[-1], -1, -1)
    else:
        z = z.view(z.shape[0], c0, -1, z.shape[-2], z.shape[-1])
    z_cam_mid = []
    for i, (_, _, s) in enumerate(block_plan(arch['camera_config'])):
        z = conv_block(z, sd, f'camera_blocks.{i}', s, arch['scale_mode'])
        z_cam_mid.append(camera_to_object(z, cam, arch['cube_size']))
    z = camera_to_object(z, cam, arch['cube_size'])
    z_obj_mid = []
    for i, (_, _, s) in enumerate(block_plan(arch['object_config']) if arch['object_config'] else []):
        z = conv_block(z, sd, f'object_blocks.{i}', s, arch['scale_mode'])
        z_obj_mid.append(z)
    z = eq_conv(z, sd, 'output_block.conv')                       # blocks.py:93-105
    return z, z_cam_mid, z_obj_mid


def voxel_coords(depth, height, width):
    """recon/utils.py:35-43 — channel order is (z, y, x)."""
    z, y, x = torch.meshgrid(torch.linspace(-1.0, 1.0, depth), torch.linspace(-1.0, 1.0, height),
                             torch.linspace(-1.0, 1.0, width), indexing='ij')
    return torch.stack((z, y, x), dim=0)


def gru_cell(sd, prefix, x, h):
    """ConvGRUCell.forward.  modules/gru.py:36-43."""
    xin = torch.cat([x, h], dim=1)
    update = torch.sigmoid(eq_conv(xin, sd, f'{prefix}.update_gate', 1))
    reset = torch.sigmoid(eq_conv(xin, sd, f'{prefix}.reset_gate', 1))
    out = eq_conv(torch.cat([x, h * reset], dim=1), sd, f'{prefix}.out_gate', 1)
    return h * (1 - update) + out * update


def fuse(kind, z_obj, sd=None):
    """z_obj [B,V,C,D,H,W] -> [B,1,C,D,H,W].  fusion.py:45-57 (pool), :184-201 (gru),
    functional.py:47-49 (abs_max)."""
    if kind == 'pool:max':
        return z_obj.max(dim=1, keepdim=True)[0]
    if kind == 'pool:mean':
        return z_obj.mean(dim=1, keepdim=True)
    if kind == 'pool:median':
        return z_obj.median(dim=1, keepdim=True)[0]
    if kind == 'pool:abs_max':
        idx = z_obj.abs().max(dim=1, keepdim=True)[1]
        return torch.gather(z_obj, 1, idx)
    if kind == 'gru':
        h = z_obj[:, 0]
        coords = voxel_coords(*h.shape[-3:])[None].expand(h.shape[0], -1, -1, -1, -1)
        for i in range(1, z_obj.shape[1]):
            h = gru_cell(sd, 'gru', torch.cat((z_obj[:, i], coords), dim=1), h)
        return h.unsqueeze(1)
    if kind == 'lstm':                      # fusion.py:234-246, modules/lstm.py:41-56
        h = z_obj[:, 0]
        c = torch.zeros_like(h)
        coords = voxel_coords(*h.shape[-3:])[None].expand(h.shape[0], -1, -1, -1, -1).to(h.dtype)
        nh = h.shape[1]
        for i in range(1, z_obj.shape[1]):
            gates = eq_conv(torch.cat((z_obj[:, i], coords, h), dim=1), sd, 'lstm.conv', 1)
            gi, gf, go, gg = torch.split(gates, nh, dim=1)
            c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
            h = torch.sigmoid(go) * torch.tanh(c)
        return h.unsqueeze(1)
    if kind == 'concat':                    # fusion.py:87-92
        n, v, ch, d, hh, w = z_obj.shape
        return z_obj.reshape(n, 1, v * ch, d, hh, w)
    raise ValueError(kind)


def sculptor_encode(sd, arch, fuser_kind, fuser_sd, cam, color, mask):
    """Sculptor.encode for colour+mask inputs.  models.py:226-258.
    color [B,V,3,H,W], mask [B,V,1,H,W] in {0,1}; the mask is fed as 2m-1 (augment :55-56)."""
    B, V = color.shape[:2]
    x = torch.cat((color.flatten(0, 1), mask.flatten(0, 1) * 2.0 - 1.0), dim=1)
    z, _, _ = sculptor_forward(sd, arch, x, cam)
    return fuse(fuser_kind, z.view(B, V, *z.shape[1:]), fuser_sd)


def photographer_forward(sd, arch, z_obj, cam):
    """Photographer.forward, eval mode, no occlusion module, no skip connections.
    models.py:397-453.  z_obj [1,C,S,S,S] (one object) -> logits [N,2,P,P], latent [N,C',S,S]."""
    z = z_obj
    for i, (_, _, s) in enumerate(block_plan(arch['object_config']) if arch['object_config'] else []):
        z = conv_block(z, sd, f'object_blocks.{i}', s, arch['scale_mode'])
    z = object_to_camera(z, cam, arch['cube_size'])
    for i, (_, _, s) in enumerate(block_plan(arch['camera_config'])):
        z = conv_block(z, sd, f'camera_blocks.{i}', s, arch['scale_mode'])
    if arch['projection_type'] == 'sum':
        z = z.sum(dim=2)
    else:                                                          # geometry.py:744-749
        z = z.reshape(z.shape[0], z.shape[1] * z.shape[2], z.shape[3], z.shape[4])
        z = pixel_norm(F.leaky_relu(eq_conv(z, sd, 'projection_block.conv'), 0.2))
    y = unet2d(z, sd, 'image_decoder', arch['image_config'], False)
    heads = [eq_conv(y, sd, f'output_blocks.{i}.conv') for i in range(arch['num_heads'])]
    return torch.cat(heads, dim=1), z


def interpret_logits(logits, apply_mask=True):
    """depth+mask heads only.  models.py:455-484."""
    y = {'depth_logits': logits[:, 0:1], 'mask_logits': logits[:, 1:2]}
    y['depth'] = torch.tanh(y['depth_logits'])
    y['mask'] = torch.sigmoid(y['mask_logits'])
    if apply_mask:
        y['depth'] = (y['depth'] + 1) * (y['mask'] > 0.5) - 1
    return y


# --------------------------------------------------------------------------------------
# L4: one pose-refinement iteration               pose/estimation.py:70-118, :601-617, :703-713
# --------------------------------------------------------------------------------------


def pose_loss(target_depth, target_mask, pred_depth, pred_mask_logits, cam):
    """default_pose_loss (no latent term).  estimation.py:70-118; pose/utils.py:81-117.
    target_* are full-frame [1,1,H,W]; pred_* are viewport crops [N,1,P,P]."""
    depth = cam.uncrop(pred_depth, 'nearest')
    mask_logits = cam.uncrop(pred_mask_logits, 'bilinear')
    mask = torch.sigmoid(mask_logits)
    depth = depth * mask
    invalid = (target_depth == 0) & (target_mask > 0.1)
    t_depth = target_depth * target_mask                          # Observation.prepare, observation.py:251-264
    valid = (~invalid).float()
    out = {}
    overlap = mask * target_mask
    dl = F.l1_loss(depth, t_depth.expand_as(depth), reduction='none') * valid
    ovm, dls = overlap.squeeze(1), dl.squeeze(1)
    out['ov_depth'] = ((dls * ovm).sum(dim=(-2, -1)).clamp(min=1e-5)
                       / ovm.sum(dim=(-2, -1)).clamp(min=1e-4))
    out['depth'] = dl.mean(dim=(1, 2, 3))
    tm = target_mask * valid
    inter = (mask * tm).sum(dim=(1, 2, 3))
    union = mask.sum(dim=(1, 2, 3)) + tm.sum(dim=(1, 2, 3)) - inter
    out['iou'] = torch.log(union.clamp(min=1e-4)) - torch.log(inter.clamp(min=1e-4))
    out['mask'] = F.binary_cross_entropy_with_logits(
        mask_logits, target_mask.expand_as(mask), reduction='none').mean(dim=(1, 2, 3))
    return out


def refine_iteration(sd, arch, z_obj, cam, target_depth, target_mask, weights):
    """Forward of one body of GradientPoseEstimator._optimize_camera (estimation.py:601-617):
    render -> denormalise depth -> loss.  Returns (optim_loss [N], loss dict, render dict)."""
    logits, latent = photographer_forward(sd, arch, z_obj, cam)
    y = interpret_logits(logits, apply_mask=True)
    z_depth = cam.denormalize_depth(y['depth'])
    losses = pose_loss(target_depth, target_mask, z_depth, y['mask_logits'], cam)
    total = sum(weights.get(k, 0.0) * v for k, v in losses.items())
    return total, losses, y, latent


# --------------------------------------------------------------------------------------
# IBR colour branch                                     latentfusion/ibr.py
# --------------------------------------------------------------------------------------


def _cam_position(cam):
    """geometry.py:219-224: C = -R^T t."""
    R = quat_to_rot(qexp3(cam.log_quaternion))
    return -(R.transpose(1, 2) @ cam.translation[:, :, None]).squeeze(-1)


def _pixel_uv(cam, h, w):
    """geometry.py:495-513 (pixel_coords_uv): viewport-spanning pixel grid, endpoints inclusive."""
    n = len(cam)
    tv, tu = torch.meshgrid(torch.linspace(0.0, 1.0, h), torch.linspace(0.0, 1.0, w), indexing='ij')
    u = tu[None] * cam.vp_w().view(n, 1, 1) + cam.viewport[:, 0].view(n, 1, 1)
    v = tv[None] * cam.vp_h().view(n, 1, 1) + cam.viewport[:, 1].view(n, 1, 1)
    return u, v


def _depth_camera_coords(cam, depth):
    """geometry.py:533-545: back-project a depth map [n,1,H,W] to camera coordinates."""
    n = len(cam)
    u, v = _pixel_uv(cam, depth.shape[-2], depth.shape[-1])
    z = depth.reshape(u.shape)
    K = cam.intrinsic
    x = (u - K[:, 0, 2].view(n, 1, 1)) / K[:, 0, 0].view(n, 1, 1) * z
    y = (v - K[:, 1, 2].view(n, 1, 1)) / K[:, 1, 1].view(n, 1, 1) * z
    return x, y, z


def _apply(transform, pts):
    """three/core.py:40-55: homogenise, multiply, de-homogenise (divide by the LAST output row)."""
    ones = torch.ones_like(pts[..., :1])
    out = torch.cat((pts, ones), dim=-1) @ transform.transpose(1, 2)
    return out[..., :-1] / out[..., -1:]


def _normalize_depth(cam, depth, eps=0.01):
    """geometry.py:560-563."""
    zn = (cam.znear() - eps).view(-1, 1, 1, 1)
    zf = (cam.zfar() + eps).view(-1, 1, 1, 1)
    return ((depth - zn) / (zf - zn)).clamp(0, 1) * 2.0 - 1.0


def _repeat_interleave(cam, k):
    return Cam(cam.intrinsic.repeat_interleave(k, 0), cam.log_quaternion.repeat_interleave(k, 0),
               cam.translation.repeat_interleave(k, 0), cam.viewport.repeat_interleave(k, 0),
               cam.z_span, cam.width, cam.height)


def ibr_warp_field(cam_in, cam_out, depth_out):
    """ibr.py:11-52 (depth_to_warp_field) -> grid [V_o, V_i, H, W, 2]."""
    vo, vi = len(cam_out), len(cam_in)
    h, w = depth_out.shape[-2:]
    x, y, z = _depth_camera_coords(cam_out, cam_out.denormalize_depth(depth_out))
    cam_pts = torch.stack((x, y, z), dim=-1).view(vo, -1, 3)
    obj = _apply(cam_out.cam_to_obj(), cam_pts)                                     # [vo, HW, 3]
    obj = obj[:, None].expand(-1, vi, -1, -1).reshape(vo * vi, -1, 3)
    o2i = (cam_in.intrinsic @ cam_in.obj_to_cam())[None].expand(vo, -1, -1, -1).reshape(vo * vi, 3, 4)
    pix = _apply(o2i, obj)                                                          # [vo*vi, HW, 2]
    vp = cam_in.viewport.repeat(vo, 1)
    gw, gh = vp[:, 2] - vp[:, 0], vp[:, 3] - vp[:, 1]
    grid = torch.stack((((pix[..., 0] - vp[:, 0, None]) / gw[:, None]) * 2 - 1,
                        ((pix[..., 1] - vp[:, 1, None]) / gh[:, None]) * 2 - 1), dim=-1)
    return grid.view(vo, vi, h, w, 2)


def ibr_reproject_views(image_in, depth_in, depth_out, cam_in, cam_out):
    """ibr.py:55-93.  image_in [V_i,C,H,W], depth_in [V_i,1,H,W] (used as given — the reference passes the
    NORMALISED depth here without denormalising it, kept), depth_out [V_o,1,H,W] ->
    (image_reproj [V_o,V_i,C,H,W], depth_reproj [V_o,V_i,1,H,W])."""
    vo, vi = len(cam_out), len(cam_in)
    grid = ibr_warp_field(cam_in, cam_out, depth_out).reshape(vo * vi, *depth_out.shape[-2:], 2)
    img = image_in[None].expand(vo, -1, -1, -1, -1).reshape(vo * vi, *image_in.shape[1:])
    x, y, z = _depth_camera_coords(cam_in, depth_in)
    obj_in = _apply(cam_in.cam_to_obj(), torch.stack((x, y, z), dim=-1).view(vi, -1, 3))        # [vi, HW, 3]
    obj_in = obj_in[None].expand(vo, -1, -1, -1).reshape(vo * vi, -1, 3)
    cam_o = _repeat_interleave(cam_out, vi)
    z_tf = _apply(cam_o.obj_to_cam(), obj_in)[..., 2].view(vo * vi, 1, *depth_in.shape[-2:])
    d_tf = _normalize_depth(cam_o, z_tf)
    image_reproj = F.grid_sample(img, grid, mode='bilinear', align_corners=False)
    depth_reproj = F.grid_sample(d_tf, grid, mode='bilinear', align_corners=False)
    return (image_reproj.view(vo, vi, *image_reproj.shape[1:]), depth_reproj.view(vo, vi, *depth_reproj.shape[1:]))


def outer_cosine_distance(x1, x2, eps=1e-8):
    """distances.py:27-32."""
    w1, w2 = x1.norm(dim=1, keepdim=True), x2.norm(dim=1, keepdim=True)
    return 1.0 - (x1 @ x2.t()) / (w1 @ w2.t()).clamp(min=eps)


def quat_angular_distance(q1, q2, eps=1e-7):
    """three/quaternion.py:372-377."""
    q1, q2 = F.normalize(q1, dim=-1, eps=1e-12), F.normalize(q2, dim=-1, eps=1e-12)
    return 2 * torch.acos(torch.clamp((q1 @ q2.t()).abs(), min=-1.0 + eps, max=1.0 - eps))


def ibr_view_weights(cam_in, cam_out, weight_type, p=0.5, eps=1e-2, depth_reproj=None, depth_out=None):
    """ibr.py:196-222: softmax-normalised blending weights over the input views."""
    if weight_type == 'cam_dist':
        d = outer_cosine_distance(_cam_position(cam_out), _cam_position(cam_in), eps=eps) / 2.0
    elif weight_type == 'cam_angle':
        d = quat_angular_distance(qexp3(cam_out.log_quaternion), qexp3(cam_in.log_quaternion)) / math.pi
    elif weight_type == 'cam_hybrid':
        dt = outer_cosine_distance(_cam_position(cam_out), _cam_position(cam_in)) / 2.0
        dr = (quat_angular_distance(qexp3(cam_out.log_quaternion), qexp3(cam_in.log_quaternion)) / (math.pi / 8)).clamp(0.0, 1.0)
        d = 1.0 - (1.0 - dt) * (1.0 - dr)
    elif weight_type == 'depth':
        diff = (depth_reproj - depth_out.unsqueeze(1).expand_as(depth_reproj)).abs()
        return torch.softmax(1.0 / ((diff / diff.max()) ** p + eps), dim=1).squeeze(2)
    else:
        raise ValueError(f'Unknown weight_type {weight_type}')
    return torch.softmax(1.0 / (d.unsqueeze(-1).unsqueeze(-1) ** p).clamp(min=eps), dim=1)


def ibr_render(cam_in, cam_out, image_in, depth_in, depth_out, p=0.5, weight_type='cam_dist', eps=1e-2):
    """ibr.py:181-228 for one object: (image_ibr [V_o,C,H,W], image_reproj [V_o,V_i,C,H,W])."""
    image_reproj, depth_reproj = ibr_reproject_views(image_in, depth_in, depth_out, cam_in, cam_out)
    wts = ibr_view_weights(cam_in, cam_out, weight_type, p, eps, depth_reproj, depth_out)
    return (wts.unsqueeze(2) * image_reproj).sum(dim=1), image_reproj


def ibr_blend_logits(logits, image_reproj):
    """ibr.py:231-234."""
    wts = torch.softmax(logits, dim=1).unsqueeze(2)
    return (wts * image_reproj).sum(dim=1), wts


def ibr_warp_blend_logits(logits, image_reproj, flow_size):
    """ibr.py:237-249: per-view softmax weights + a bounded flow refinement of each reprojection."""
    b, vi = image_reproj.shape[:2]
    h, w = image_reproj.shape[-2:]
    bl, fxl, fyl = torch.split(logits, vi, dim=1)
    wts = torch.softmax(bl, dim=1).unsqueeze(2)
    dx = flow_size / w * torch.tanh(fxl)
    dy = flow_size / h * torch.tanh(fyl)
    gy, gx = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing='ij')
    grid = torch.stack((gx[None, None].expand_as(dx) + dx, gy[None, None].expand_as(dy) + dy), dim=-1).clamp(-1, 1)
    warped = F.grid_sample(image_reproj.reshape(b * vi, *image_reproj.shape[2:]), grid.reshape(b * vi, h, w, 2),
                           mode='bilinear', align_corners=False)
    warped = warped.view(b, vi, *warped.shape[1:])
    return (wts * warped).sum(dim=1), wts, dx, dy
