"""dev probe: lf_camera_o2c_bwd (fp32 analytic VJP) vs fp64 autograd of the torch camera algebra, fed with REALISTIC
block gradients (large, cancelling) from an fp64 evaluation of the resampler backward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tests import parity_helpers as ph
from latentfusion_b200 import ops

dev = torch.device('cuda:0')
g, model, z_obj, target = ph.config_b_case(dev, smooth=True)
S, C = 64, 32
cam = ph.product_camera(g.cam('hyp_cam'), dev, requires_grad=True)
blk = cam.o2c_block(1.0)
torch.manual_seed(3)
gv = F.avg_pool3d(F.pad(torch.randn(2, C, S, S, S, device=dev), (2,) * 6, mode='replicate'), 5, stride=1)
dt = torch.float64
b = blk.detach().to(dt).clone().requires_grad_(True)
n = 2
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
gB = b.grad.detach()            # fp64 block gradient [2, 40]

# ours: fp32 VJP kernel
(blk * gB.float()).sum().backward()
ours = torch.cat([cam.log_quaternion.grad, cam.translation.grad, cam.viewport.grad], 1).cpu().double()
# exact: torch chain on CPU in fp64
c64 = ph.product_camera({k: v.double() for k, v in g.cam('hyp_cam').items()}, 'cpu', requires_grad=True)
b64 = c64.o2c_block(1.0)
print('cpu block dtype', b64.dtype)
(b64 * gB.cpu().to(b64.dtype)).sum().backward()
ex = torch.cat([c64.log_quaternion.grad, c64.translation.grad, c64.viewport.grad], 1).double()
# fp32 torch chain on CPU
c32 = ph.product_camera(g.cam('hyp_cam'), 'cpu', requires_grad=True)
b32 = c32.o2c_block(1.0)
(b32 * gB.cpu().float()).sum().backward()
t32 = torch.cat([c32.log_quaternion.grad, c32.translation.grad, c32.viewport.grad], 1).double()
for name, a in (('lf_camera_o2c_bwd', ours), ('torch fp32 chain', t32)):
    print(name, ['%.2e' % float((a[:, s] - ex[:, s]).abs().max() / ex[:, s].abs().max()) for s in (slice(0, 3), slice(3, 6), slice(6, 10))])
print('exact grads', ex)
print('block grads', gB[:, :16], gB[:, 20])
