"""GPU: the tcgen05 implicit-GEMM convolution (precision 1 = bf16x3 split, 2 = bf16) against the exact
fp32 FFMA kernel of the same library (itself pinned to the reference by test_gpu_parity.py) and against
torch's fp64 convolution.  Tolerances: bf16x3 is fp32-parity grade (atol 1e-4 / rtol 1e-3 like every other
output); plain bf16 gets the looser stated bound 3e-2."""
import math

import pytest
import torch
import torch.nn.functional as F

from latentfusion_b200 import ops

pytestmark = pytest.mark.gpu

SHAPES = [
    # nd, n, cin, cout, size, k
    (3, 2, 32, 32, 16, 3),
    (3, 1, 32, 32, 64, 3),
    (3, 1, 16, 48, 12, 3),
    (3, 1, 32, 32, 20, 1),
    (2, 2, 32, 64, 64, 3),
    (2, 2, 64, 64, 32, 3),
    (2, 3, 64, 32, 50, 3),
    (2, 2, 32, 2, 128, 1),
    (2, 2, 4, 32, 40, 1),
]


def ref_conv(x, w, b, act, norm):
    conv = F.conv3d if x.dim() == 5 else F.conv2d
    he = math.sqrt(2.0 / w[0].numel())
    y = conv(x.double(), w.double(), None, padding=w.shape[-1] // 2) * he + b.double().view(1, -1, *([1] * (x.dim() - 2)))
    if act:
        y = F.leaky_relu(y, 0.2)
    if norm:
        y = y / torch.sqrt(torch.mean(y ** 2, dim=1, keepdim=True) + 1e-8)
    return y


@pytest.mark.parametrize('nd,n,cin,cout,size,k', SHAPES)
@pytest.mark.parametrize('precision', [1, 2])
def test_conv_tc_forward_backward(nd, n, cin, cout, size, k, precision):
    dev = torch.device('cuda:0')
    torch.manual_seed(nd * 1000 + cin + cout + size)
    shape = (n, cin) + (size,) * nd
    x = torch.randn(*shape, device=dev)
    w = torch.randn(cout, cin, *([k] * nd), device=dev)
    b = torch.randn(cout, device=dev) * 0.1
    norm = cout > 2
    xt = x.clone().requires_grad_(True)
    y = ops.eq_conv(xt, w, b, act=True, norm=norm, precision=precision)
    ref = ref_conv(x, w, b, True, norm)
    tol = dict(atol=1e-4, rtol=1e-3) if precision == 1 else dict(atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(y.double(), ref, **tol)
    # backward to the input goes through the same kernel with flipped/transposed weights.  It is checked on
    # the LINEAR layer (no LeakyReLU/PixelNorm): with the non-linearity, outputs within rounding distance of 0
    # legitimately flip the LeakyReLU gate between two precisions, which is not a kernel property.
    xl = x.clone().requires_grad_(True)
    yl = ops.eq_conv(xl, w, b, act=False, norm=False, precision=precision)
    g = torch.randn_like(yl)
    yl.backward(g)
    xr = x.double().clone().requires_grad_(True)
    conv = F.conv3d if nd == 3 else F.conv2d
    (conv(xr, w.double(), None, padding=k // 2) * math.sqrt(2.0 / w[0].numel())).backward(g.double())
    gtol = dict(atol=2e-4, rtol=2e-3) if precision == 1 else dict(atol=5e-2, rtol=5e-2)
    torch.testing.assert_close(xl.grad.double(), xr.grad, **gtol)


def test_conv_tc_unsupported_shape_falls_back_to_exact_kernel():
    dev = torch.device('cuda:0')
    x = torch.randn(1, 67, 6, 6, 6, device=dev)          # GRU gate: Cin = 2C+3 is not a multiple of 4
    w = torch.randn(32, 67, 3, 3, 3, device=dev)
    y1 = ops.eq_conv(x, w, None, precision=1)
    y0 = ops.eq_conv(x, w, None, precision=0)
    assert torch.equal(y1, y0)


def test_fused_backward_staging_matches_unfused():
    """lf_conv_bwd_data_fused (PixelNorm/LeakyReLU backward folded into the tcgen05 operand staging) against
    lf_actnorm_bwd + lf_conv_fwd on the same tensors."""
    dev = torch.device('cuda:0')
    torch.manual_seed(3)
    x = torch.randn(2, 32, 16, 16, 16, device=dev)
    w = torch.randn(32, 32, 3, 3, 3, device=dev)
    b = torch.randn(32, device=dev) * 0.1
    grads = []
    for fuse in (False, True):
        ops._FUSE_BWD = fuse
        try:
            xt = x.clone().requires_grad_(True)
            y = ops.eq_conv(xt, w, b, act=True, norm=True, precision=2)
            torch.manual_seed(4)
            y.backward(torch.randn_like(y))
            grads.append(xt.grad)
        finally:
            ops._FUSE_BWD = False
    torch.testing.assert_close(grads[1], grads[0], atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize('precision', [1, 2])
def test_gru_cell_channel_group_split_matches_oracle(precision):
    """ConvGRUCell on the tensor-core path evaluates each gate as conv_x(x) + conv_h(h) (modules/gru.py); the
    oracle concatenates like the reference (modules/gru.py:36-43).  Includes the backward to x and h."""
    from oracle import lf_oracle as O
    from latentfusion_b200.modules.gru import ConvGRUCell
    dev = torch.device('cuda:0')
    torch.manual_seed(11)
    C = 32
    cell = ConvGRUCell(C + 3, C, kernel_size=3).to(dev)
    for p in cell.parameters():
        p.requires_grad_(False)
    for g in (cell.update_gate, cell.reset_gate, cell.out_gate):
        g.bias.normal_(0, 0.1)
    x = torch.randn(1, C + 3, 12, 12, 12, device=dev)
    h = torch.randn(1, C, 12, 12, 12, device=dev)
    sd = {'gru.' + k: v.double().cpu() for k, v in cell.state_dict().items()}
    xr, hr = x.double().cpu().requires_grad_(True), h.double().cpu().requires_grad_(True)
    ref = O.gru_cell(sd, 'gru', xr, hr)
    g = torch.randn_like(ref)
    ref.backward(g)
    old = ops.get_default_precision()
    ops.set_default_precision(precision)
    try:
        xt, ht = x.clone().requires_grad_(True), h.clone().requires_grad_(True)
        ops.KernelTrace.reset(True)
        out = cell(xt, ht)
        names = [r[0] for r in ops.KernelTrace.records]
        out.backward(g.float().to(dev))
    finally:
        ops.KernelTrace.reset(False)
        ops.set_default_precision(old)
    # every one of the 3 gates x 2 channel groups went through the tcgen05 kernel (its weights were packed for it)
    # (per-tap kernel or, where the shape fits, the depth-batched one)
    assert names.count('lf_conv_tc_pack_weights') + names.count('lf_conv3d_dz_pack_weights') == 6, names
    tol = dict(atol=2e-4, rtol=2e-3) if precision == 1 else dict(atol=5e-2, rtol=5e-2)
    torch.testing.assert_close(out.double().cpu(), ref.detach(), **tol)
    torch.testing.assert_close(xt.grad.double().cpu(), xr.grad, **tol)
    torch.testing.assert_close(ht.grad.double().cpu(), hr.grad, **tol)



@pytest.mark.parametrize('n,c,cout,depth,hw', [(2, 8, 8, 16, 16), (1, 32, 32, 64, 64), (2, 12, 20, 12, 9), (1, 16, 32, 6, 20),
                                                (3, 32, 16, 8, 7)])
def test_depth_projections_vs_torch_fp64(n, c, cout, depth, hw):
    """FactorProjection3d2d / 2d3d (geometry.py:724-749) as the dedicated kernels: the cluster split-depth collapse
    (taps % 8 / 4 / 2 == 0 or the generic kernel otherwise) and the multi-slice expand, forward and backward, against
    torch's fp64 1x1 convolution over the C*D channel view.  Exact-fp32 kernels -> the fp32 tolerance."""
    dev = torch.device('cuda:0')
    torch.manual_seed(n * 100 + c)
    x = torch.randn(n, c, depth, hw, hw, device=dev)
    w = torch.randn(cout, c * depth, 1, 1, device=dev)
    b = torch.randn(cout, device=dev) * 0.1
    y = ops.eq_conv(x, w, b, act=True, norm=True, kind=ops.KIND_COLLAPSE, depth=depth, precision=0)
    torch.testing.assert_close(y.double(), ref_conv(x.reshape(n, c * depth, hw, hw), w, b, True, True), atol=1e-4, rtol=1e-3)
    # backward on the linear layer (a LeakyReLU gate may legitimately flip for pre-activations within rounding of 0)
    xt = x.clone().requires_grad_(True)
    yl = ops.eq_conv(xt, w, b, kind=ops.KIND_COLLAPSE, depth=depth, precision=0)
    xr = x.double().requires_grad_(True)
    ref = ref_conv(xr.reshape(n, c * depth, hw, hw), w, b, False, False)
    torch.testing.assert_close(yl.double(), ref, atol=1e-4, rtol=1e-3)
    g = torch.randn_like(yl)
    yl.backward(g)
    ref.backward(g.double())
    torch.testing.assert_close(xt.grad.double(), xr.grad, atol=1e-4, rtol=1e-3)
    # expand: [n, cin, h, w] -> [n, c, depth, h, w]
    x2 = torch.randn(n, cout, hw, hw, device=dev)
    w2 = torch.randn(c * depth, cout, 1, 1, device=dev)
    b2 = torch.randn(c * depth, device=dev) * 0.1
    y2 = ops.eq_conv(x2, w2, b2, act=True, norm=False, kind=ops.KIND_EXPAND, depth=depth, precision=0)
    torch.testing.assert_close(y2.double(), ref_conv(x2, w2, b2, True, False).reshape(n, c, depth, hw, hw), atol=1e-4, rtol=1e-3)
    x2t = x2.clone().requires_grad_(True)
    y2l = ops.eq_conv(x2t, w2, b2, kind=ops.KIND_EXPAND, depth=depth, precision=0)
    x2r = x2.double().requires_grad_(True)
    ref2 = ref_conv(x2r, w2, b2, False, False).reshape(n, c, depth, hw, hw)
    g2 = torch.randn_like(y2l)
    y2l.backward(g2)
    ref2.backward(g2.double())
    torch.testing.assert_close(x2t.grad.double(), x2r.grad, atol=1e-4, rtol=1e-3)


@pytest.mark.parametrize('precision', [1, 2])
def test_epilogue_fused_actnorm_backward_matches_separate_kernels(precision):
    """Block (conv1 -> conv2, both LeakyReLU + PixelNorm) followed by the depth collapse: with LFB200_FUSE_EPI=1 (opt-in:
    measured slower than the separate kernels, see ops._FUSE_EPI) the bwd-data kernels of conv2 and of the collapse apply the activation/norm backward of their producer in the epilogue
    (lf_conv_bwd_data_epi) and the standalone lf_actnorm_bwd passes over the 3-D tensors disappear."""
    from latentfusion_b200.modules.blocks import Block
    from latentfusion_b200.modules import EqualizedConv3d
    dev = torch.device('cuda:0')
    torch.manual_seed(21)
    C, S = 32, 16
    block = Block(C, C, conv_module=EqualizedConv3d).to(dev)
    for p_ in block.parameters():
        p_.requires_grad_(False)
    wc = torch.randn(C, C * S, 1, 1, device=dev)
    x = torch.randn(2, C, S, S, S, device=dev)
    g = torch.randn(2, C, S, S, device=dev)
    results = []
    old, old_fuse = ops.get_default_precision(), ops._FUSE_EPI
    ops.set_default_precision(precision)
    try:
        for fuse in (False, True):
            ops._FUSE_EPI = fuse
            xt = x.clone().requires_grad_(True)
            z = block(xt)
            z = ops.eq_conv(ops.mark_single_consumer(z), wc, None, act=True, norm=True, kind=ops.KIND_COLLAPSE, depth=S)
            ops.KernelTrace.reset(True)
            z.backward(g)
            names = [r[0] for r in ops.KernelTrace.records]
            ops.KernelTrace.reset(False)
            results.append((xt.grad.clone(), names))
    finally:
        ops._FUSE_EPI = old_fuse
        ops.set_default_precision(old)
    (g_sep, n_sep), (g_fused, n_fused) = results
    # (the 3x3x3 convolutions run on the depth-batched kernel, whose bwd-data always carries conv1's activation backward
    # in its epilogue; what LFB200_FUSE_EPI adds is the same fusion inside the depth-expand of the collapse's backward)
    assert n_sep.count('lf_actnorm_bwd') == 2 and not any('epi' in n for n in n_sep), n_sep
    assert n_fused.count('lf_actnorm_bwd') == 1 and sum('bwd_data_epi' in n for n in n_fused) == 1, n_fused
    tol = dict(atol=2e-5, rtol=1e-4) if precision == 1 else dict(atol=2e-3, rtol=2e-3)
    torch.testing.assert_close(g_fused, g_sep, **tol)


@pytest.mark.parametrize('c,heads', [(32, (1, 1)), (64, (3, 1, 1)), (8, (1,)), (128, (1, 1))])
def test_fused_output_heads_match_separate_1x1_convs(c, heads):
    """lf_heads_fwd/bwd (all 1x1 output heads of the decoder in one pass) against the per-head exact-fp32 convolutions
    concatenated along channels, forward, backward to the features and the training gradients of weights and biases."""
    dev = torch.device('cuda:0')
    torch.manual_seed(c + len(heads))
    x = torch.randn(3, c, 20, 17, device=dev)
    ws = [torch.randn(h, c, 1, 1, device=dev, requires_grad=True) for h in heads]
    bs = [(torch.randn(h, device=dev) * 0.1).requires_grad_(True) for h in heads]
    xa = x.clone().requires_grad_(True)
    ya = ops.fused_heads(xa, ws, bs)
    xb = x.clone().requires_grad_(True)
    yb = torch.cat([ops.eq_conv(xb, w, b, precision=0) for w, b in zip(ws, bs)], dim=1)
    torch.testing.assert_close(ya, yb, atol=1e-5, rtol=1e-5)
    g = torch.randn_like(ya)
    ga = torch.autograd.grad(ya, [xa] + ws + bs, g)
    gb = torch.autograd.grad(yb, [xb] + ws + bs, g)
    for a, b in zip(ga, gb):
        torch.testing.assert_close(a, b, atol=2e-4, rtol=1e-4)
