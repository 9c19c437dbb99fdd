"""dev probe: sampling-coordinate error (voxels) of block-based fp32 / param-based fp32 vs param-based fp64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import parity_helpers as ph
from oracle import lf_oracle as O

dev = torch.device('cuda:0')
g, model, z_obj, target = ph.config_b_case(dev, smooth=True)
S = 64
cam = ph.product_camera(g.cam('hyp_cam'), dev)
blk = cam.o2c_block(1.0)


def grid_from_block(b, dt):
    b = b.to(dt)
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
    return ((M @ pts.transpose(2, 1)).transpose(1, 2) / 0.5).view(n, S, S, S, 3)


def grid_from_params(dt):
    torch.set_default_device(dev); torch.set_default_dtype(dt)
    try:
        d = {k: v.to(dev).to(dt) for k, v in g.cam('hyp_cam').items()}
        return O.o2c_grid(ph.oracle_camera(d), S, 1.0)
    finally:
        torch.set_default_dtype(torch.float32); torch.set_default_device('cpu')


ref = grid_from_params(torch.float64)
to_vox = S / 2.0
for name, gr in (('params fp32 (oracle/ATen)', grid_from_params(torch.float32)), ('block fp32', grid_from_block(blk, torch.float32)),
                 ('block (fp32 entries) evaluated in fp64', grid_from_block(blk, torch.float64))):
    e = (gr.double() - ref) * to_vox
    print(f'{name:42s} coord err voxels: max {float(e.abs().max()):.2e} rms {float(e.pow(2).mean().sqrt()):.2e} mean(per axis) {[float(v) for v in e.mean(dim=(0,1,2,3))]}')
print('block entries vs fp64 cam_to_obj:')
torch.set_default_dtype(torch.float64)
d = {k: v.double() for k, v in g.cam('hyp_cam').items()}
oc = ph.oracle_camera(d)
M64 = oc.cam_to_obj()[:, :3].reshape(2, 12)
print((blk[:, :12].cpu().double() - M64).abs().max(dim=0).values)
print('znear err', (blk[:, 20].cpu().double() - oc.znear()).abs())
