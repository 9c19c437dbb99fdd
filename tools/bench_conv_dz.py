"""dev: time the depth-batched conv kernel at config B (N=8, 32->32, 64^3) vs the per-tap tcgen05 kernel."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latentfusion_b200 import ops

dev = torch.device('cuda:0')
n, c, s = int(os.environ.get('N', 8)), 32, 64
torch.manual_seed(0)
x = torch.randn(n, c, s, s, s, device=dev).contiguous(memory_format=torch.channels_last_3d)
wt = torch.randn(c, c, 3, 3, 3, device=dev)
b = torch.randn(c, device=dev) * 0.1
he = math.sqrt(2.0 / (c * 27))
wf, _ = ops._pack_weight(wt, ops.KIND_CONV, 0)
wpk = ops._dz_pack(wf, (wt, id(wt), wt._version, 'b'))


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


xs = ops.split_pack(x)
flops = 2 * 27 * c * c * n * s ** 3
print('split_pack            %.3f ms' % timeit(lambda: ops.split_pack(x)))
for prec in (1, 2):
    for dense, split in ((True, False), (False, True), (True, True)):
        t = timeit(lambda: ops.conv3d_dz(xs, wpk, b, c, he, True, 0.2, True, prec, want_dense=dense, want_split=split))
        print(f'conv3d_dz prec {prec} dense={dense} split={split}: {t:.3f} ms  {flops / t / 1e9:.0f} TF/s algorithmic')
    t = timeit(lambda: ops.eq_conv(x, wt, b, act=True, norm=True, precision=prec))
    print(f'eq_conv (current path) prec {prec}: {t:.3f} ms')
