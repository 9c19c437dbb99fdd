"""dev: time the tcgen05 weight-gradient kernel (lf_conv3d_dw) at the training shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latentfusion_b200 import ops
dev = torch.device('cuda:0')
for n, c, s, prec in ((8, 32, 64, 1), (32, 32, 64, 1), (2, 32, 64, 1), (8, 32, 64, 2), (2, 16, 128, 1)):
    x = torch.randn(n, c, s, s, s, device=dev)
    du = torch.randn(n, c, s, s, s, device=dev)
    xs, ds = ops.split_pack(x), ops.split_pack(du)
    del x, du
    for _ in range(2):
        ops.conv3d_dw(xs, ds, prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        ops.conv3d_dw(xs, ds, prec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2 * 27 * c * c * n * s ** 3
    print(f'n={n} C={c} S={s} prec={prec}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TF/s algorithmic, {2 * (xs.buf.numel() + ds.buf.numel()) / ms / 1e6:.0f} GB/s of operands')
    del xs, ds
