"""dev probe: lf_resample_o2c_bwd_cam block gradients vs an fp64 torch evaluation from the SAME fp32 camera block,
and lf_camera_o2c_bwd vs fp64 autograd of the camera algebra."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tests import parity_helpers as ph
from latentfusion_b200 import ops

dev = torch.device('cuda:0')
smooth = '--smooth' in sys.argv
g, model, z_obj, target = ph.config_b_case(dev, smooth=smooth)
S, C = 64, 32
cam = ph.product_camera(g.cam('hyp_cam'), dev, requires_grad=True)
blk = cam.o2c_block(1.0)
torch.manual_seed(3)
# a smooth upstream gradient field (like the real one) and a white one
gsm = F.avg_pool3d(F.pad(torch.randn(2, C, S, S, S, device=dev), (2,) * 6, mode='replicate'), 5, stride=1)
for name, gv in (('smooth g', gsm), ('white g', torch.randn(2, C, S, S, S, device=dev))):
    b1 = blk.detach().clone().requires_grad_(True)
    out = ops.resample_o2c(z_obj[0], b1)
    (out * gv).sum().backward()
    ours = b1.grad.double()

    def torch_from_block(dt):
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
        obj = (M @ pts.transpose(2, 1)).transpose(1, 2) / 0.5
        grid = obj.view(n, S, S, S, 3)
        o = F.grid_sample(z_obj[0].to(dt).expand(n, -1, -1, -1, -1), grid, padding_mode='border', align_corners=False)
        (o * gv.to(dt)).sum().backward()
        return b.grad.double(), o.detach()
    g64, o64 = torch_from_block(torch.float64)
    g32, o32 = torch_from_block(torch.float32)
    print(name, ' fwd abs err ours/aten32 vs fp64:', float((out.double() - o64).abs().max()), float((o32.double() - o64).abs().max()))
    idx = list(range(16)) + [20]
    for label, gg in (('ours', ours), ('aten fp32', g32)):
        rel = [(float((gg[:, i] - g64[:, i]).abs().max() / g64[:, i].abs().max())) for i in idx]
        print(f'  {label:9s} block-grad rel err per term:', ' '.join('%.1e' % r for r in rel))
    print('  |g64| per term:', ' '.join('%.2e' % float(g64[:, i].abs().max()) for i in idx))

# camera algebra VJP
gb = torch.randn_like(blk)
c2 = ph.product_camera(g.cam('hyp_cam'), dev, requires_grad=True)
(c2.o2c_block(1.0) * gb).sum().backward()
c3 = ph.product_camera(g.cam('hyp_cam'), 'cpu', requires_grad=True)
for k in ('log_quaternion', 'translation', 'viewport', 'intrinsic'):
    getattr(c3, k).data = getattr(c3, k).data.double()
