import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latentfusion_b200 import ops
dev = torch.device('cuda:0')
n, c, s = 8, 32, 64
prec = int(os.environ.get('PREC', 1))
x = torch.randn(n, c, s, s, s, device=dev).contiguous(memory_format=torch.channels_last_3d)
wt = torch.randn(c, c, 3, 3, 3, device=dev); b = torch.randn(c, device=dev) * 0.1
wf, _ = ops._pack_weight(wt, ops.KIND_CONV, 0)
wpk = ops._dz_pack(wf, (wt, id(wt), wt._version, 'b'))
xs = ops.split_pack(x)
for _ in range(3):
    ops.conv3d_dz(xs, wpk, b, c, math.sqrt(2.0 / (c * 27)), True, 0.2, True, prec, want_dense=os.environ.get('SPLIT', '0') == '0', want_split=os.environ.get('SPLIT', '0') == '1')
torch.cuda.synchronize()
