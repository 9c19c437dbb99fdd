"""dev: one small launch of the resamplers (fwd, bwd_cam, bwd_vol), the depth-batched conv (fwd + fused bwd) and the
weight-gradient kernel — the workload for `compute-sanitizer --tool memcheck|racecheck` (profiles/r02_sanitizer.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latentfusion_b200 import ops
from latentfusion_b200.modules import EqualizedConv3d
from latentfusion_b200.modules.blocks import Block
from latentfusion_b200.modules.geometry import ObjectToCameraTransform, CameraToObjectTransform
from tests import parity_helpers as ph
dev = torch.device('cuda:0')
ops.set_default_precision(1)
S, C, N = 16, 16, 2
cams, _ = ph.synthetic_cameras(N, S, seed=3)
cam = cams.to(dev)
vol = torch.randn(1, C, S, S, S, device=dev, requires_grad=True)
blk = Block(C, C, conv_module=EqualizedConv3d, scale_factor=1.0).to(dev)
z = ObjectToCameraTransform(1.0)(vol, cam)
y = blk(z)
back = CameraToObjectTransform(1.0)(y, cam)
back.sum().backward()
torch.cuda.synchronize()
print('ok', float(vol.grad.abs().sum()), float(blk.conv1.module.weight.grad.abs().sum()))
