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


@pytest.mark.parametrize('c,size', [(32, (2, 9, 12, 10)), (16, (1, 6, 7, 8))])
@pytest.mark.parametrize('train', [False, True])
def test_block_forward_backward_through_dz_path(dev, c, size, train):
    """Block(C, C) (reference modules/blocks.py:152-164) at precision 1: conv1 hands conv2 its split-planar output,
    conv2's bwd-data applies conv1's PixelNorm/LeakyReLU backward in its epilogue and hands conv1 the split-planar du.
    Checked against fp64 autograd of the same composition: output, d/dx and (train) the weight/bias gradients."""
    from latentfusion_b200 import ops
    from latentfusion_b200.modules import EqualizedConv3d
    from latentfusion_b200.modules.blocks import Block
    n, d, h, w = size
    torch.manual_seed(c + d)
    blk = Block(c, c, conv_module=EqualizedConv3d, scale_factor=1.0).to(dev)
    for k, p in blk.named_parameters():
        if k.endswith('bias'):
            p.data.normal_(0, 0.1)
        p.requires_grad_(train)
    x = torch.randn(n, c, d, h, w, device=dev, requires_grad=True)
    g = torch.randn(n, c, d, h, w, device=dev)
    old = ops.get_default_precision()
    ops.set_default_precision(1)
    try:
        ops.KernelTrace.reset(True)
        y = blk(x)
        y.backward(g)
        names = [r[0] for r in ops.KernelTrace.records]
    finally:
        ops.KernelTrace.reset(False)
        ops.set_default_precision(old)
    # only the block's input is packed (in training once more in the backward, for conv1's weight gradient: the
    # split-planar twin of a leaf input is not kept alive between forward and backward)
    assert names.count('lf_split_pack') == (2 if train else 1), names
    assert ('lf_conv3d_dw' in names) == train and 'lf_conv_bwd_weight' not in names
    assert 'lf_conv_bwd_data[conv3d_k3]' in names
    xr = x.detach().double().requires_grad_(True)
    ws = {k: v.detach().double().requires_grad_(True) for k, v in blk.named_parameters()}
    t = xr
    fragile = torch.zeros(n, 1, d, h, w, device=dev)
    for i in (1, 2):
        he = math.sqrt(2.0 / (c * 27))
        t = F.conv3d(t, ws[f'conv{i}.module.weight'], None, padding=1) * he + ws[f'conv{i}.bias'].view(1, -1, 1, 1, 1)
        # a pre-activation within rounding of zero may take the other LeakyReLU branch in fp32: its gradient then differs
        # by 0.8x on that channel, which the bwd-data convolutions spread over a 3^3 (layer 1) / 5^3 (layer 2) stencil
        near0 = (t.detach().abs() < 2e-4).any(dim=1, keepdim=True).float()
        fragile = torch.maximum(fragile, F.max_pool3d(near0, 2 * i + 1, stride=1, padding=i))
        t = F.leaky_relu(t, 0.2)
        t = t / torch.sqrt((t * t).mean(dim=1, keepdim=True) + 1e-8)
    t.backward(g.double())
    torch.testing.assert_close(y.detach().double(), t.detach(), atol=2e-4, rtol=2e-3)
    keep = (fragile == 0).expand_as(x)
    assert keep.float().mean() > 0.5
    torch.testing.assert_close(x.grad.double()[keep], xr.grad[keep], atol=2e-4, rtol=2e-3)
    if train:
        # (sums over all positions: a flipped LeakyReLU branch at a near-zero pre-activation moves them by one term)
        for k, p in blk.named_parameters():
            rel = float((p.grad.double() - ws[k].grad).norm() / ws[k].grad.norm())
            assert rel < 2e-2, f'{k}: relative L2 error {rel:.3g}'


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(2, 32, 32, 16, 16, 16, 1), (1, 8, 8, 10, 12, 14, 1), (3, 32, 16, 5, 20, 9, 1),
                                   (2, 32, 32, 12, 16, 16, 2), (1, 16, 32, 1, 24, 24, 1), (2, 12, 20, 7, 33, 17, 1),
                                   (1, 16, 16, 6, 7, 8, 1),      # (narrow plane: the rounded-up run ends past the buffer)
                                   (1, 64, 96, 4, 16, 16, 1), (1, 128, 60, 3, 12, 14, 1)])   # wide: 32x32 channel-group pairs
def test_conv3d_weight_gradient_on_tensor_cores_vs_fp64(shape):
    """lf_conv3d_dw (MN-major tcgen05 over the split-planar twins) against autograd of F.conv3d in fp64: weight gradient
    [27][Cin][Cout] and bias gradient; precision 1 = all four bf16x2 split products (fp32-grade), 2 = bf16 operands."""
    import torch.nn.functional as F
    from latentfusion_b200 import ops
    n, cin, cout, d, h, w, prec = shape
    dev = torch.device('cuda:0')
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, d, h, w, device=dev)
    du = torch.randn(n, cout, d, h, w, device=dev)
    gwp, gbp = ops.conv3d_dw(ops.split_pack(x), ops.split_pack(du), prec)
    wz = torch.zeros(cout, cin, 3, 3, 3, device=dev, dtype=torch.float64, requires_grad=True)
    F.conv3d(x.double(), wz, padding=1).backward(du.double())
    ref = wz.grad.permute(2, 3, 4, 1, 0).reshape(27, cin, cout)
    err = float((gwp.double() - ref).norm() / ref.norm())
    worst = float((gwp.double() - ref).abs().max() / ref.abs().max())
    tol = 2e-5 if prec == 1 else 1e-2
    assert err < tol and worst < 10 * tol, f'relative L2 {err:.3g}, worst/max {worst:.3g}'
    bref = du.double().sum(dim=(0, 2, 3, 4))
    assert float((gbp.double().reshape(-1) - bref).abs().max() / bref.abs().max()) < (1e-4 if prec == 1 else 1e-2)
    gwp2, _ = ops.conv3d_dw(ops.split_pack(x), ops.split_pack(du), prec)
    assert torch.equal(gwp, gwp2)                       # fixed-order reduction: bit-reproducible


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(3, 32, 64, 24, 20, 1), (2, 64, 64, 16, 16, 1), (4, 4, 32, 33, 17, 1), (2, 64, 32, 12, 12, 1),
                                   (2, 32, 32, 16, 16, 2), (2, 256, 128, 16, 16, 1), (1, 128, 196, 24, 20, 1)])
def test_conv2d_weight_gradient_on_tensor_cores_vs_fp64(shape):
    """the same kernel on 2-D maps (the U-Nets' 3x3 convolutions): one plane per image, centre dz only, up to 64 channels"""
    import torch.nn.functional as F
    from latentfusion_b200 import ops
    n, cin, cout, h, w, prec = shape
    dev = torch.device('cuda:0')
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w, device=dev)
    du = torch.randn(n, cout, h, w, device=dev)
    gwp, gbp = ops.conv3d_dw(ops.split_pack(x.unsqueeze(2)), ops.split_pack(du.unsqueeze(2)), prec, ndim=2)
    wz = torch.zeros(cout, cin, 3, 3, device=dev, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), wz, padding=1).backward(du.double())
    ref = wz.grad.permute(2, 3, 1, 0).reshape(9, cin, cout)
    err = float((gwp.double() - ref).norm() / ref.norm())
    assert err < (2e-5 if prec == 1 else 1e-2), f'relative L2 {err:.3g}'
    bref = du.double().sum(dim=(0, 2, 3))
    assert float((gbp.double().reshape(-1) - bref).abs().max() / bref.abs().max()) < (1e-4 if prec == 1 else 1e-2)


@pytest.mark.parametrize('shape', [(2, 64, 128, 6, 16, 16), (1, 128, 64, 5, 16, 16), (1, 256, 256, 3, 16, 16), (2, 64, 64, 4, 12, 10),
                                   (1, 96, 192, 3, 16, 16)])
@pytest.mark.parametrize('precision', [1, 2])
def test_wide_conv_weight_streaming_vs_fp64(dev, shape, precision):
    """Released network widths (reference tools/train/train.sh:37-46: 64/128/256-channel 3-D blocks on a 16^3 latent):
    eq_conv routes them to the weight-streaming tcgen05 kernel (csrc/conv3d_ws.cu), forward (scale, bias, LeakyReLU,
    PixelNorm over all channel chunks) and bwd-data; checked against fp64 autograd of the reference op."""
    from latentfusion_b200 import ops
    n, cin, cout, d, h, w = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, d, h, w, device=dev, requires_grad=True)
    wt = torch.randn(cout, cin, 3, 3, 3, device=dev, requires_grad=True)
    b = torch.randn(cout, device=dev) * 0.1
    g = torch.randn(n, cout, d, h, w, device=dev)
    ops.KernelTrace.reset(True)
    try:
        y = ops.eq_conv(x, wt, b, act=True, norm=True, precision=precision)
        y.backward(g)
        names = [r[0] for r in ops.KernelTrace.records]
    finally:
        ops.KernelTrace.reset(False)
    # forward, bwd-data and (channel-group pairs of) the weight gradient all on the tensor cores
    assert 'lf_conv_bwd_weight' not in names and 'lf_conv3d_dw' in names and names.count('lf_conv_fwd[conv3d_k3]') == 1, names
    x64 = x.detach().double().requires_grad_(True)
    w64 = wt.detach().double().requires_grad_(True)
    he = math.sqrt(2.0 / (cin * 27))
    u = F.conv3d(x64, w64, None, padding=1) * he + b.double().view(1, -1, 1, 1, 1)
    u = F.leaky_relu(u, 0.2)
    ref = u / torch.sqrt((u * u).mean(dim=1, keepdim=True) + 1e-8)
    ref.backward(g.double())
    tol = dict(atol=2e-4, rtol=2e-3) if precision == 1 else dict(atol=4e-2, rtol=4e-2)
    torch.testing.assert_close(y.double(), ref, **tol)
    gerr = float((x.grad.double() - x64.grad).norm() / x64.grad.norm())
    assert gerr < (1e-4 if precision == 1 else 5e-2), f'd/dx relative L2 {gerr:.3g}'     # (plain bf16 operands, K up to 27*256)
    werr = float((wt.grad.double() - w64.grad).norm() / w64.grad.norm())
    assert werr < (1e-4 if precision == 1 else 5e-2), f'd/dW relative L2 {werr:.3g}'


@pytest.mark.parametrize('shape', [(2, 512, 512, 8, 8), (1, 256, 512, 16, 16), (2, 128, 196, 64, 64), (1, 196, 128, 128, 128),
                                   (3, 64, 128, 33, 17)])
def test_wide_conv2d_weight_streaming_vs_fp64(dev, shape):
    """the U-Nets' wide 3x3 layers (released widths: 128..512 channels on 4^2..128^2 maps) on the same kernel: one plane
    per image, maps larger than one TMEM / shared-memory slab cut into tile groups.  The gradient is checked tightly
    without the LeakyReLU (a pre-activation within fp32 rounding of zero takes the other branch than in fp64, and with
    hundreds of channels per PixelNorm row one flip perturbs the whole row) and loosely with it."""
    from latentfusion_b200 import ops
    n, cin, cout, h, w = shape
    torch.manual_seed(sum(shape))
    wt = torch.randn(cout, cin, 3, 3, device=dev)
    b = torch.randn(cout, device=dev) * 0.1
    g = torch.randn(n, cout, h, w, device=dev)
    he = math.sqrt(2.0 / (cin * 9))
    for act, gtol in ((False, 1e-4), (True, 2e-2)):
        x = torch.randn(n, cin, h, w, device=dev, requires_grad=True)
        ops.KernelTrace.reset(True)
        try:
            y = ops.eq_conv(x, wt, b, act=act, norm=True, precision=1)
            y.backward(g)
            launched = {r[0] for r in ops.KernelTrace.records}
        finally:
            ops.KernelTrace.reset(False)
        x64 = x.detach().double().requires_grad_(True)
        u = F.conv2d(x64, wt.double(), None, padding=1) * he + b.double().view(1, -1, 1, 1)
        if act:
            u = F.leaky_relu(u, 0.2)
        ref = u / torch.sqrt((u * u).mean(dim=1, keepdim=True) + 1e-8)
        ref.backward(g.double())
        torch.testing.assert_close(y.double(), ref, atol=2e-4, rtol=2e-3)
        gerr = float((x.grad.double() - x64.grad).norm() / x64.grad.norm())
        assert gerr < gtol, f'd/dx relative L2 {gerr:.3g} (act={act})'
        assert 'lf_conv_fwd[conv2d_k3]' in launched and 'lf_conv_bwd_data[conv2d_k3]' in launched
