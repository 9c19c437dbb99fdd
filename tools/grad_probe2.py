"""dev probe: where does the config-B camera-gradient deviation from fp64 come from?
Runs the oracle (reference algorithm in torch ops) on the GPU in fp32 and fp64, captures the gradient w.r.t. the
resampled volume, and swaps pieces against the lfb200 kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from tests import parity_helpers as ph
from latentfusion_b200 import ops
from oracle import lf_oracle as O

dev = torch.device('cuda:0')
g, model, z_obj, target = ph.config_b_case(dev, smooth='--smooth' in sys.argv)
names = ('ov_depth', 'depth', 'iou', 'mask')
w = g.meta['weights']
g64 = torch.cat([g['grad64.log_quaternion'], g['grad64.translation'], g['grad64.viewport']], 1)
sl3 = (slice(0, 3), slice(3, 6), slice(6, 10))


def rel(a, b=g64):
    return ['%.2e' % float((a.double().cpu()[:, s] - b[:, s]).abs().max() / b[:, s].abs().max()) for s in sl3]


def oracle_run(dtype):
    torch.set_default_device(dev)
    torch.set_default_dtype(dtype)
    old = O.resample
    O.resample = lambda vol, grid: F.grid_sample(vol.to(dtype), grid.to(dtype), padding_mode='border', align_corners=False)
    cap = {}
    oto = O.object_to_camera

    def capture(obj_vol, cam, cube_size=1.0):
        out = oto(obj_vol, cam, cube_size)
        out.retain_grad()
        cap['v0'] = out
        return out
    O.object_to_camera = capture
    try:
        d = {k: v.to(dev).to(dtype) for k, v in g.cam('hyp_cam').items()}
        cam = ph.oracle_camera(d, requires_grad=True)
        sd = {k: v.to(dev).to(dtype) for k, v in g.state_dict('photographer').items()}
        arch = ph.oracle_arch(g.meta, 'photographer')
        total, losses, y, latent = O.refine_iteration(sd, arch, z_obj[0].to(dtype), cam, target.depth.to(dtype), target.mask.to(dtype), w)
        total.mean().backward()
        grads = torch.cat([cam.log_quaternion.grad, cam.translation.grad, cam.viewport.grad], 1)
        return grads.detach(), cap['v0'].grad.detach(), cap['v0'].detach()
    finally:
        O.resample, O.object_to_camera = old, oto
        torch.set_default_dtype(torch.float32)
        torch.set_default_device('cpu')


G64, gv64, v64 = oracle_run(torch.float64)
print('oracle fp64 on GPU vs reference fp64 golden:', rel(G64))
G32, gv32, v32 = oracle_run(torch.float32)
print('oracle fp32 on GPU (ATen/cuDNN, TF32 off) vs fp64:', rel(G32))
print('  its g_v0 rel err (max-norm):', float((gv32.double() - gv64).abs().max() / gv64.abs().max()),
      ' l2:', float((gv32.double() - gv64).norm() / gv64.norm()))

# ours, capturing g_v0
ops.set_default_precision(0)
for prec in (0, 1):
    ops.set_default_precision(prec)
    cap = {}
    tb = model.photographer.transform_block
    orig_fwd = tb.forward

    def fwd(z, camera):
        out = orig_fwd(z, camera)
        out.retain_grad()
        cap['v0'] = out
        return out
    tb.forward = fwd
    cam = ph.product_camera(g.cam('hyp_cam'), dev, requires_grad=True)
    y, latent = model.render_latent_object(z_obj, cam, return_latent=True, apply_mask=True)
    terms = ops.pose_loss_terms(y['depth_logits'].squeeze(0)[:, 0], y['mask_logits'].squeeze(0)[:, 0], cam.viewport,
                                cam.translation[:, 2], target.depth, target.mask, cam.z_span, 0.01, cam.width, cam.height)
    sum(w[k] * terms[:, i] for i, k in enumerate(names)).mean().backward()
    tb.forward = orig_fwd
    ours = torch.cat([cam.log_quaternion.grad, cam.translation.grad, cam.viewport.grad], 1)
    gv = cap['v0'].grad
    print(f'ours precision {prec}: total grads vs fp64', rel(ours))
    print('  g_v0 rel err max-norm:', float((gv.double() - gv64).abs().max() / gv64.abs().max()),
          ' l2:', float((gv.double() - gv64).norm() / gv64.norm()))
    print('  v0 (resample fwd) abs err:', float((cap['v0'].double() - v64).abs().max()))
    # resample-path gradient only: (a) our bwd_cam with OUR g_v0, (b) fp64 autograd with OUR g_v0, (c) our bwd_cam with fp64-exact g_v0
    def ours_rs(gvol):
        c = ph.product_camera(g.cam('hyp_cam'), dev, requires_grad=True)
        v = tb(z_obj[0], c)
        (v * gvol.float()).sum().backward()
        return torch.cat([c.log_quaternion.grad, c.translation.grad, c.viewport.grad], 1)

    def exact_rs(gvol, dt=torch.float64):
        torch.set_default_device(dev); torch.set_default_dtype(dt)
        old = O.resample
        O.resample = lambda vol, grid: F.grid_sample(vol.to(dt), grid.to(dt), padding_mode='border', align_corners=False)
        try:
            d = {k: v.to(dev).to(dt) for k, v in g.cam('hyp_cam').items()}
            c = ph.oracle_camera(d, requires_grad=True)
            v = O.object_to_camera(z_obj[0].to(dt), c)
            (v * gvol.to(dt)).sum().backward()
            return torch.cat([c.log_quaternion.grad, c.translation.grad, c.viewport.grad], 1)
        finally:
            O.resample = old
            torch.set_default_dtype(torch.float32); torch.set_default_device('cpu')
    a, b, c_ = ours_rs(gv), exact_rs(gv), ours_rs(gv64)
    e = exact_rs(gv64)
    scale = [float(g64[:, s].abs().max()) for s in sl3]
    def relto(x, y):
        return ['%.2e' % (float((x.double().cpu()[:, s] - y.double().cpu()[:, s]).abs().max()) / scale[i]) for i, s in enumerate(sl3)]
    print('  resample-path: ours(bwd_cam, our g) vs exact(our g):', relto(a, b))
    print('  resample-path: exact(our g) vs exact(exact g)     :', relto(b, e))
    print('  resample-path: ours(exact g) vs exact(exact g)     :', relto(c_, e))
    print('  resample-path: ATen fp32(exact g) vs exact(exact g):', relto(exact_rs(gv64, torch.float32), e))
    print('  resample-path magnitude / total scale:', ['%.2f' % (float(e.cpu()[:, s].abs().max()) / scale[i]) for i, s in enumerate(sl3)])
