"""dev probe: with the REAL upstream gradient g_v0 (fp64 oracle), compare resample-path camera grads:
 (a) fp64 from params  (b) ATen fp32 from params  (c) ours (block kernel + bwd_cam + VJP)
 (d) ATen fp32 from OUR fp32 block -> our VJP     (e) ATen fp64 from OUR fp32 block -> our VJP"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from tests import parity_helpers as ph
from latentfusion_b200 import ops
from oracle import lf_oracle as O

dev = torch.device('cuda:0')
g, model, z_obj, target = ph.config_b_case(dev, smooth='--smooth' in sys.argv)
S, C = 64, 32
w = g.meta['weights']
sl3 = (slice(0, 3), slice(3, 6), slice(6, 10))


def oracle_gv(dtype):
    torch.set_default_device(dev); torch.set_default_dtype(dtype)
    old, oto = O.resample, O.object_to_camera
    O.resample = lambda vol, grid: F.grid_sample(vol.to(dtype), grid.to(dtype), padding_mode='border', align_corners=False)
    cap = {}

    def capture(obj_vol, cam, cube_size=1.0):
        out = oto(obj_vol, cam, cube_size); out.retain_grad(); cap['v0'] = out
        return out
    O.object_to_camera = capture
    try:
        d = {k: v.to(dev).to(dtype) for k, v in g.cam('hyp_cam').items()}
        cam = ph.oracle_camera(d, requires_grad=True)
        sd = {k: v.to(dev).to(dtype) for k, v in g.state_dict('photographer').items()}
        total, *_ = O.refine_iteration(sd, ph.oracle_arch(g.meta, 'photographer'), z_obj[0].to(dtype), cam,
                                       target.depth.to(dtype), target.mask.to(dtype), w)
        total.mean().backward()
        return cap['v0'].grad.detach()
    finally:
        O.resample, O.object_to_camera = old, oto
        torch.set_default_dtype(torch.float32); torch.set_default_device('cpu')


gv = oracle_gv(torch.float64)


def from_params(dt):
    torch.set_default_device(dev); torch.set_default_dtype(dt)
    old = O.resample
    O.resample = lambda vol, grid: F.grid_sample(vol.to(dt), grid.to(dt), padding_mode='border', align_corners=False)
    try:
        d = {k: v.to(dev).to(dt) for k, v in g.cam('hyp_cam').items()}
        c = ph.oracle_camera(d, requires_grad=True)
        v = O.object_to_camera(z_obj[0].to(dt), c)
        (v * gv.to(dt)).sum().backward()
        return torch.cat([c.log_quaternion.grad, c.translation.grad, c.viewport.grad], 1).double().cpu()
    finally:
        O.resample = old
        torch.set_default_dtype(torch.float32); torch.set_default_device('cpu')


def block_grad_torch(blk, dt):
    b = blk.detach().to(dt).clone().requires_grad_(True)
    n = b.shape[0]
    lin = torch.linspace(0.0, 1.0, S, device=dev, dtype=dt)
    zp, vp, up = torch.meshgrid(lin, lin, lin, indexing='ij')
    u = up[None] * b[:, 14].view(n, 1, 1, 1) + b[:, 12].view(n, 1, 1, 1)
    v = vp[None] * b[:, 15].view(n, 1, 1, 1) + b[:, 13].view(n, 1, 1, 1)
    z = zp[None] * b[:, 21].view(n, 1, 1, 1) + b[:, 20].view(n, 1, 1, 1)
    x = (u - b[:, 16].view(n, 1, 1, 1)) / b[:, 18].view(n, 1, 1, 1) * z
    y = (v - b[:, 17].view(n, 1, 1, 1)) / b[:, 19].view(n, 1, 1, 1) * z
    M = b[:, :12].view(n, 3, 4)
    pts = torch.stack((x, y, z, torch.ones_like(x)), dim=-1).view(n, -1, 4)
    grid = ((M @ pts.transpose(2, 1)).transpose(1, 2) / 0.5).view(n, S, S, S, 3)
    o = F.grid_sample(z_obj[0].to(dt).expand(n, -1, -1, -1, -1), grid, padding_mode='border', align_corners=False)
    (o * gv.to(dt)).sum().backward()
    return b.grad.detach()


def via_vjp(block_grad):
    cam = ph.product_camera(g.cam('hyp_cam'), dev, requires_grad=True)
    (cam.o2c_block(1.0) * block_grad.float()).sum().backward()
    return torch.cat([cam.log_quaternion.grad, cam.translation.grad, cam.viewport.grad], 1).double().cpu()


a = from_params(torch.float64)
b_ = from_params(torch.float32)
cam = ph.product_camera(g.cam('hyp_cam'), dev, requires_grad=True)
blk = cam.o2c_block(1.0)
v = ops.resample_o2c(z_obj[0], blk)
(v * gv.float()).sum().backward()
c_ = torch.cat([cam.log_quaternion.grad, cam.translation.grad, cam.viewport.grad], 1).double().cpu()
gb_ours = None
b1 = blk.detach().clone().requires_grad_(True)
(ops.resample_o2c(z_obj[0], b1) * gv.float()).sum().backward()
gb_ours = b1.grad.double()
gb32, gb64 = block_grad_torch(blk, torch.float32).double(), block_grad_torch(blk, torch.float64)
d_, e_ = via_vjp(gb32), via_vjp(gb64)


def rel(x):
    return ['%.2e' % float((x[:, s] - a[:, s]).abs().max() / a[:, s].abs().max()) for s in sl3]


print('(b) ATen fp32 from params        :', rel(b_))
print('(c) ours                         :', rel(c_))
print('(d) ATen fp32 from our block+VJP :', rel(d_))
print('(e) ATen fp64 from our block+VJP :', rel(e_))
idx = list(range(16)) + [20]
print('block-grad rel err ours  vs fp64-from-block:', ' '.join('%.1e' % float((gb_ours[:, i] - gb64[:, i]).abs().max() / gb64[:, i].abs().max()) for i in idx))
print('block-grad rel err aten32 vs fp64-from-block:', ' '.join('%.1e' % float((gb32[:, i] - gb64[:, i]).abs().max() / gb64[:, i].abs().max()) for i in idx))
print('exact', a)
