"""GPU tests of the depth-batched tcgen05 3x3x3 convolution (csrc/conv3d_dz.cu) through the C ABI: against an fp64
torch evaluation of the reference op (modules/equalized.py:57-64 + blocks.py:152-158: conv * he + bias ->
LeakyReLU -> PixelNorm), at the fp32 parity tolerance for precision 1 (bf16x3) and a stated 3e-2 for precision 2."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def reference(x, w, b, act, norm, slope=0.2):
    he = math.sqrt(2.0 / (w.shape[1] * 27))
    y = F.conv3d(x.double(), w.double(), None, padding=1) * he
    if b is not None:
        y = y + b.double().view(1, -1, 1, 1, 1)
    if act:
        y = F.leaky_relu(y, slope)
    r = None
    if norm:
        r = torch.sqrt((y * y).mean(dim=1, keepdim=True) + 1e-8)
        y = y / r
    return y, r


SHAPES = [  # n, cin, cout, d, h, w
    (2, 32, 32, 6, 12, 12),
    (1, 16, 16, 9, 10, 14),
    (3, 16, 32, 5, 7, 9),
    (1, 32, 16, 4, 20, 6),
    (1, 8, 12, 3, 5, 5),           # channel padding (8 -> 16, 12 -> 16)
    (2, 32, 32, 64, 64, 64),       # BASELINE configs[1] extent: 17 tile columns, whole-depth march, ring wraps
    (1, 32, 32, 40, 33, 47),       # depth chunking (few columns -> several chunks per column), odd extents
    (1, 32, 32, 1, 9, 9),          # single plane
    (1, 32, 32, 2, 9, 9),
]


@pytest.mark.parametrize('n,cin,cout,d,h,w', SHAPES)
@pytest.mark.parametrize('precision', [1, 2])
def test_conv3d_dz_vs_fp64(dev, n, cin, cout, d, h, w, precision):
    from latentfusion_b200 import ops
    torch.manual_seed(n * 1000 + cin + d)
    x = torch.randn(n, cin, d, h, w, device=dev)
    wt = torch.randn(cout, cin, 3, 3, 3, device=dev)
    b = torch.randn(cout, device=dev) * 0.1
    desc = ops._desc(ops.KIND_CONV, 3, n, d, h, w, cin, cout, 3, 1.0, True, 0.2, True, precision)
    assert ops._dz_ok(desc)
    wf, wb = ops._pack_weight(wt, ops.KIND_CONV, 0)
    wpk = ops._dz_pack(wf, (wt, id(wt), wt._version, 'test'))
    xs = ops.split_pack(x)
    torch.testing.assert_close(xs.to_dense(), x, atol=2e-5, rtol=2e-5)          # hi + lo carries 16 mantissa bits
    tol = dict(atol=1e-4, rtol=1e-3) if precision == 1 else dict(atol=3e-2, rtol=3e-2)
    he = math.sqrt(2.0 / (cin * 27))
    for act, norm in ((True, True), (False, False)):
        ref, r = reference(x, wt, b, act, norm)
        y, ys, rn = ops.conv3d_dz(xs, wpk, b, cout, he, act, 0.2, norm, precision, want_dense=True, want_split=True)
        torch.testing.assert_close(y.double(), ref, **tol)
        torch.testing.assert_close(ys.to_dense().double(), ref, atol=tol['atol'] + 2e-5, rtol=tol['rtol'])
        if norm:
            torch.testing.assert_close(rn.view(n, 1, d, h, w).double(), r, **tol)
        # the split-planar output carries its zero halo (the next convolution's padding)
        cp = (cout + 15) // 16 * 16
        raw = ys.buf.view(torch.bfloat16).view(2, n, d, cp // 8, h + 2, w + 2, 8).float()
        assert float(raw[..., 0, :, :].abs().max()) == 0 and float(raw[..., -1, :, :].abs().max()) == 0
        assert float(raw[..., :, 0, :].abs().max()) == 0 and float(raw[..., :, -1, :].abs().max()) == 0
    # chaining: conv -> conv through the split-planar output only (what a Block does)
    if cin == cout:
        ref1, _ = reference(x, wt, b, True, True)
        ref2, _ = reference(ref1.float(), wt, b, True, True)
        _, ys1, _ = ops.conv3d_dz(xs, wpk, b, cout, he, True, 0.2, True, precision, want_dense=False, want_split=True)
        y2, _, _ = ops.conv3d_dz(ys1, wpk, b, cout, he, True, 0.2, True, precision)
        torch.testing.assert_close(y2.double(), ref2, atol=tol['atol'] * 3, rtol=tol['rtol'] * 3)
