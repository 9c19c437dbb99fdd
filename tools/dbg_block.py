import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latentfusion_b200 import ops
from latentfusion_b200.modules import EqualizedConv3d
from latentfusion_b200.modules.blocks import Block
dev = torch.device('cuda:0')
c, (n, d, h, w) = 32, (2, 9, 12, 10)
torch.manual_seed(c + d)
blk = Block(c, c, conv_module=EqualizedConv3d, scale_factor=1.0).to(dev)
for k, p in blk.named_parameters():
    p.requires_grad_(False)
x0 = torch.randn(n, c, d, h, w, device=dev)
g = torch.randn(n, c, d, h, w, device=dev)

def run(prec, fuse=True):
    ops.set_default_precision(prec)
    orig = ops._dz_shape
    if not fuse:
        ops._dz_shape = lambda *a: False
    x = x0.clone().requires_grad_(True)
    ops.KernelTrace.reset(True)
    y = blk(x); y.backward(g)
    names = [r[0] for r in ops.KernelTrace.records]
    ops.KernelTrace.reset(False)
    ops._dz_shape = orig
    return y.detach(), x.grad.detach(), names

y0, g0, _ = run(0)
y1, g1, n1 = run(1, True)
y2, g2, n2 = run(1, False)
print(n1); print(n2)
print('fused  vs exact: y', (y1 - y0).abs().max().item(), 'gx', (g1 - g0).abs().max().item())
print('unfused vs exact: y', (y2 - y0).abs().max().item(), 'gx', (g2 - g0).abs().max().item())
bad = ((g1 - g0).abs() > 1e-3).nonzero()
print(len(bad), bad[:10].tolist(), bad[-5:].tolist())
import collections
print(collections.Counter((b[0], b[2]) for b in bad.tolist()))
