"""Training iteration (BASELINE configs[3]): ReconTrainStep on the lfb200 kernels against the golden of
the unmodified reference's modules + losses (oracle/make_golden_train.py), and the host-side collectives (differentiable
view all-gather, flat gradient all-reduce) under gloo with two processes."""
import json
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden', 'train_s16_c8.npz')


def _load():
    z = np.load(GOLD)
    return z, json.loads(str(z['meta']))


def _build(z, meta, dev):
    from latentfusion_b200.modules.geometry import Camera
    from latentfusion_b200.recon import fusion, models

    def sd(prefix):
        return {k[len(prefix) + 1:]: torch.from_numpy(np.array(z[k])) for k in z.files if k.startswith(prefix + '/')}

    def cam(prefix):
        t = {k: torch.from_numpy(np.array(z[f'{prefix}.{k}'])) for k in ('intrinsic', 'log_quaternion', 'translation', 'viewport')}
        return Camera(t['intrinsic'], None, 0.5, t['viewport'], width=640, height=480,
                      log_quaternion=t['log_quaternion'], translation=t['translation']).to(dev)
    sculptor = models.Sculptor(**meta['arch_sculptor'])
    sculptor.load_state_dict(sd('sculptor'), strict=True)
    photographer = models.Photographer(**meta['arch_photographer'])
    photographer.load_state_dict(sd('photographer'), strict=True)
    fuser = fusion.get_fuser('gru', in_channels=meta['C'], cube_size=1.0)
    fuser.load_state_dict(sd('fuser'), strict=True)
    t = lambda k: torch.from_numpy(np.array(z[k])).to(dev)      # noqa: E731
    batch = {'in': {'camera': cam('cam_in'), 'image': t('in.image'), 'mask': t('in.mask')},
             'out_gt': {'camera': cam('cam_out'), 'depth': t('gt.depth'), 'mask': t('gt.mask')}}
    return sculptor.to(dev), fuser.to(dev), photographer.to(dev), batch


@pytest.mark.gpu
@pytest.mark.parametrize('precision', [0, 1])
def test_train_iteration_vs_reference_golden(precision):
    """losses, a spread of weight/bias gradients over all three networks, and the parameters after one Adam step.
    precision 0 = exact FFMA kernels; 1 = tcgen05 bf16x3 forward/backward-data (weight gradients stay exact fp32)."""
    from latentfusion_b200 import ops
    from latentfusion_b200.train import ReconTrainStep
    z, meta = _load()
    dev = torch.device('cuda:0')
    old = ops.get_default_precision()
    ops.set_default_precision(precision)
    try:
        sculptor, fuser, photographer, batch = _build(z, meta, dev)
        cfg = meta['cfg']
        step = ReconTrainStep(sculptor, fuser, photographer, lr=cfg['lr'], depth_weight=cfg['depth_weight'],
                              mask_weight=cfg['mask_weight'], beta_weight=cfg['beta_weight'], beta_param=cfg['beta_param'],
                              depth_k=cfg['depth_k'])
        named = {f'{n}/{k}': p for n, m in (('sculptor', sculptor), ('fuser', fuser), ('photographer', photographer))
                 for k, p in m.named_parameters()}
        with torch.no_grad():
            d0, m0, ml0 = step.forward(batch)
        torch.testing.assert_close(d0.cpu(), torch.from_numpy(np.array(z['fwd.depth'])), atol=2e-4, rtol=2e-3)
        torch.testing.assert_close(ml0.cpu(), torch.from_numpy(np.array(z['fwd.mask_logits'])), atol=2e-4, rtol=2e-3)
        out = step.run_iteration(batch, train=True, is_step=False)
        for k in ('depth', 'mask', 'beta', 'total'):
            torch.testing.assert_close(out[k].cpu(), torch.from_numpy(np.array(z[f'loss.{k}'])), atol=2e-4, rtol=1e-3)
        gn = float(sum((p.grad.double() ** 2).sum() for p in step.parameters).sqrt())
        assert abs(gn - float(z['gradnorm'][0])) <= 5e-3 * float(z['gradnorm'][0])
        for k in meta['pick']:
            ref = torch.from_numpy(np.array(z[f'grad/{k}']))
            err = float((named[k].grad.cpu() - ref).norm() / ref.norm().clamp(min=1e-6))
            assert err < 5e-3, f'{k}: relative L2 gradient error {err:.3g}'
        step.optimizer.step()
        for k in meta['pick']:
            # Adam's first step moves every weight by ~lr * sign(grad): compare the UPDATE, not the weight
            before = torch.from_numpy(np.array(z[k]))
            upd_ref = torch.from_numpy(np.array(z[f'after/{k}'])) - before
            upd = named[k].detach().cpu() - before
            agree = float(((upd - upd_ref).abs() < 0.2 * cfg['lr']).float().mean())
            assert agree > 0.97, f'{k}: only {agree:.3f} of the Adam updates agree'
    finally:
        ops.set_default_precision(old)


def _worker(rank, world, port, result_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from latentfusion_b200 import dist as lfdist
    torch.manual_seed(0)
    B, V, F = 2, 4, 3
    full = torch.randn(B, V, F)
    w = torch.randn(V * F, 5)
    lin = torch.nn.Linear(5, 1)
    ok = True
    # single-process reference: loss = mean over the V "output views" of a function of the gathered tensor
    ref_in = full.clone().requires_grad_(True)
    ref_lin = torch.nn.Linear(5, 1)
    ref_lin.load_state_dict(lin.state_dict())
    h = ref_in.reshape(B, V * F) @ w
    per_view = [(ref_lin(torch.tanh(h * (v + 1)))).sum() for v in range(V)]
    (sum(per_view) / V).backward()
    # sharded: each rank owns V/world input views and V/world of the loss terms
    lo, hi = rank * V // world, (rank + 1) * V // world
    local = full[:, lo:hi].clone().requires_grad_(True)
    gathered = lfdist.all_gather_views(local, dim=1)
    ok &= torch.equal(gathered.detach(), full)
    h = gathered.reshape(B, V * F) @ w
    share = sum((lin(torch.tanh(h * (v + 1)))).sum() for v in range(lo, hi)) / V
    share.backward()
    ok &= torch.allclose(local.grad, ref_in.grad[:, lo:hi], atol=1e-5)
    lfdist.allreduce_gradients(lin.parameters())
    ok &= torch.allclose(lin.weight.grad, ref_lin.weight.grad, atol=1e-5) and torch.allclose(lin.bias.grad, ref_lin.bias.grad, atol=1e-5)
    open(os.path.join(result_dir, f'ok{rank}'), 'w').write(str(bool(ok)))
    dist.destroy_process_group()


def test_differentiable_view_gather_and_gradient_allreduce_world2(tmp_path):
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f'ok{r}').read() for r in range(2)] == ['True', 'True']
