import sys, os, math, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latentfusion_b200 import ops, _lib as L
dev = torch.device('cuda:0')
for (n, c, d, h, w) in ((2, 32, 9, 12, 10), (1, 32, 9, 12, 10), (2, 16, 9, 12, 10), (1, 32, 3, 12, 10), (1, 32, 9, 64, 64)):
    torch.manual_seed(0)
    du = torch.randn(n, c, d, h, w, device=dev)
    yprev = torch.randn(n, c, d, h, w, device=dev)
    rprev = torch.rand(n * d * h * w, device=dev) + 0.5
    wt = torch.randn(c, c, 3, 3, 3, device=dev)
    wf, wb = ops._pack_weight(wt, ops.KIND_CONV, 0)
    wpk = ops._dz_pack(wb, (wt, id(wt), wt._version, 'dbg'))
    he = 0.05
    dus, ys = ops.split_pack(du), ops.split_pack(yprev)
    bdesc = ops._desc(ops.KIND_CONV, 3, n, d, h, w, c, c, 3, he, 0, 0.0, 0, 1)
    gx = ops.empty_cl((n, c, d, h, w), dev); gxs = ops.SplitVol.empty(n, c, d, h, w, dev)
    L.check(L.lib().lf_conv3d_dz_bwd_epi(ctypes.byref(bdesc), ops._p(dus.buf), ops._p(wpk), ops._p(ys.buf), ops._p(rprev), 1, 0.2, 1,
                                         ops._p(gx), ops._p(gxs.buf), ops._stream()), 'epi')
    g, _, _ = ops.conv3d_dz(dus, wpk, None, c, he, False, 0.0, False, 1)
    # separate actnorm
    ref = torch.empty_like(g)
    L.check(L.lib().lf_actnorm_bwd(ops._p(g), ops._p(ops.to_cl(yprev)), ops._p(rprev), ops._p(ref), n * d * h * w, 1, 1, c, 1, 0.2, 1, ops._stream()), 'an')
    # split actnorm kernel
    s2 = ops.SplitVol.empty(n, c, d, h, w, dev); d2 = torch.empty_like(g)
    L.check(L.lib().lf_actnorm_bwd_split(ops._p(g), ops._p(ops.to_cl(yprev)), ops._p(rprev), ops._p(d2), ops._p(s2.buf), n, d, h, w, c, 1, 0.2, 1, ops._stream()), 'ans')
    torch.cuda.synchronize()
    e1 = (gx - ref).abs().max().item(); e2 = (gxs.to_dense() - ref).abs().max().item()
    e3 = (d2 - ref).abs().max().item(); e4 = (s2.to_dense() - ref).abs().max().item()
    bad = ((gx - ref).abs() > 1e-3).nonzero()
    print((n, c, d, h, w), 'epi dense', e1, 'epi split', e2, 'actnorm_split dense', e3, 'split', e4, 'ref max', ref.abs().max().item(), 'nbad', len(bad), bad[:5].tolist())
