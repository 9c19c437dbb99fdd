import sys, os, math, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latentfusion_b200 import ops, _lib as L
dev = torch.device('cuda:0')
n, c, s = 8, 32, 64
prec = int(os.environ.get('PREC', 1))
x = torch.randn(n, c, s, s, s, device=dev).contiguous(memory_format=torch.channels_last_3d)
wt = torch.randn(c, c, 3, 3, 3, device=dev); b = torch.randn(c, device=dev) * 0.1
wf, _ = ops._pack_weight(wt, ops.KIND_CONV, 0)
wpk = ops._dz_pack(wf, (wt, id(wt), wt._version, 'b'))
xs = ops.split_pack(x)
he = math.sqrt(2.0 / (c * 27))
for _ in range(2):
    ops.conv3d_dz(xs, wpk, b, c, he, True, 0.2, True, prec)
desc = ops._desc(ops.KIND_CONV, 3, n, s, s, s, c, c, 3, he, True, 0.2, True, prec)
y = ops.empty_cl((n, c, s, s, s), dev); rn = torch.empty(n * s ** 3, device=dev)
st = torch.zeros(4 * 64 * 2, dtype=torch.int64, device=dev)
L.check(L.lib().lf_conv3d_dz_timeline(ctypes.byref(desc), ops._p(xs.buf), ops._p(wpk), ops._p(b), ops._p(y), None, ops._p(rn), ops._p(st), ops._stream()), 'tl')
torch.cuda.synchronize()
t = st.cpu().view(4, 64, 2)
t0 = int(t[0, 0, 0])
names = ['producer (empty-wait done -> copies issued)', 'issuer0 (slab ready -> issued)', 'issuer1', 'epilogue g0 (acc ready -> drained)']
for r in range(4):
    print(names[r])
    print('  ', ' '.join(f'{int(t[r, i, 0]) - t0}-{int(t[r, i, 1]) - t0}' for i in range(0, 24)))
