"""Coarse pose search on the device path (SURVEY §8 'configs[4]' row): CrossEntropyPoseEstimator scoring through the
fused loss head + latent cosine, S=128 render parity, and the host logic of sample / hypothesis sharding under gloo."""
import json
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CE_LATENT = {'type': 'cross_entropy',          # reference configs/cross_entropy_latent.toml
             'args': dict(num_samples=24, num_elites=4, num_iters=2, num_gmm_components=2, learning_rate=0.5,
                          sample_flipped=True, ranking_size=4),
             'loss_weights': dict(depth=1.0, ov_depth=0.3, iou=0.0, mask=0.1, latent=1.0)}


@pytest.mark.gpu
def test_cross_entropy_scores_device_path_equals_torch_composition():
    """the fused forward-only loss head + latent cosine vs default_pose_loss on uncropped full frames (ATen ops), for a
    population of 24 sampled cameras x 4 flips, latent weight 1.0"""
    from tests import parity_helpers as ph
    from latentfusion_b200.observation import Observation
    from latentfusion_b200.pose import estimation, utils as pu
    from latentfusion_b200.recon.inference import LatentFusionModel
    g = ph.Golden()
    dev = torch.device('cuda:0')
    sculptor, fuser, photographer = ph.build_product_models(g, dev)
    model = LatentFusionModel(sculptor, fuser, photographer, g.meta['camera_dist'], dev)
    est = estimation.load_from_config(json.loads(json.dumps(CE_LATENT)), model)
    gt = ph.product_camera(g.cam('ref_cam_full'), 'cpu')[0:1]
    torch.manual_seed(3)
    target = Observation(torch.rand(1, 3, 480, 640), g['target.depth'], g['target.mask'], gt).to(dev)
    cams = pu.sample_cameras_with_estimate(n=24, camera_est=gt).to(dev)
    z_obj = g['z_obj_gru'].to(dev)
    with torch.no_grad():
        code = model.compute_latent_code(target, cams[0])
        fused = est._score(z_obj, target, cams, code)
        est.loss_func = lambda *a, **k: estimation.default_pose_loss(*a, **k)        # not `is default` -> torch path
        plain = est._score(z_obj, target, cams, code)
    torch.testing.assert_close(fused, plain, atol=2e-4, rtol=2e-3)
    assert torch.equal(torch.argsort(fused)[:4], torch.argsort(plain)[:4])


@pytest.mark.gpu
def test_render_at_s128_matches_oracle():
    """BASELINE configs[4] extent: 128^3 latent volume (C=16), one camera, precision 1 — the depth-batched convolution
    plans 130-wide padded planes (one tile column per plane row pair), resample + collapse at S=128."""
    from oracle import lf_oracle as O
    from tests import parity_helpers as ph
    from latentfusion_b200 import ops
    from latentfusion_b200.recon.inference import LatentFusionModel
    dev = torch.device('cuda:0')
    S, C = 128, 16
    sculptor, fuser, photographer, arch, sds = ph.random_lfsynth(S, C, seed=4, device=dev)
    cams, dist_ = ph.synthetic_cameras(1, S, seed=6)
    model = LatentFusionModel(sculptor, fuser, photographer, dist_, dev)
    import torch.nn.functional as F
    torch.manual_seed(8)
    z = F.avg_pool3d(F.pad(torch.randn(1, C, S, S, S), (1,) * 6, mode='replicate'), 3, stride=1)
    z = z / z.std() * 0.5
    old = ops.get_default_precision()
    ops.set_default_precision(1)
    try:
        with torch.no_grad():
            y, latent = model.render_latent_object(z.unsqueeze(0).to(dev), cams.to(dev), return_latent=True)
    finally:
        ops.set_default_precision(old)
    with torch.no_grad():
        logits, olat = O.photographer_forward(sds['photographer'], arch['photographer'], z, ph.oracle_camera(ph.cam_to_dict(cams)))
    torch.testing.assert_close(y['depth_logits'].cpu()[0], logits[:, 0:1], atol=5e-4, rtol=2e-3)
    torch.testing.assert_close(y['mask_logits'].cpu()[0], logits[:, 1:2], atol=5e-4, rtol=2e-3)
    torch.testing.assert_close(latent.cpu(), olat, atol=5e-4, rtol=2e-3)


def _worker(rank, world, port, result_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from latentfusion_b200.modules.geometry import Camera
    from latentfusion_b200.pose import estimation
    from latentfusion_b200 import consts

    class _Model:
        device = 'cpu'
    ok = True
    ce = estimation.CrossEntropyPoseEstimator(model=_Model(), num_samples=7, num_elites=2, num_iters=1,
                                              num_gmm_components=1, learning_rate=0.5, ranking_size=2, loss_weights={})
    ce.group = dist.group.WORLD
    lo, hi = ce._shard(7)
    full = torch.arange(7, dtype=torch.float32) * 1.5
    ok &= torch.equal(ce._gather_scores(full[lo:hi].clone(), 7), full)
    gr = estimation.GradientPoseEstimator(model=_Model(), learning_rate=0.01, num_samples=4, num_iters=1,
                                          converge_threshold=1e-6, converge_patience=10, ranking_size=3, loss_weights={})
    gr.group = dist.group.WORLD
    n = 2
    K = torch.tensor(consts.INTRINSIC).unsqueeze(0).expand(n, -1, -1).contiguous()
    local = Camera(K, None, 0.5, None, width=640, height=480, log_quaternion=torch.full((n, 3), 0.1 * (rank + 1)),
                   translation=torch.tensor([[0.0, 0.0, 1.0 + rank], [0.0, 0.0, 3.0 + rank]]))
    losses = torch.tensor([0.5, 2.0]) if rank == 0 else torch.tensor([1.0, 0.1])
    merged = gr._merge_ranked(local, losses)
    ok &= len(merged) == 3 and merged.translation[:, 2].tolist() == [4.0, 1.0, 2.0]      # losses 0.1, 0.5, 1.0
    open(os.path.join(result_dir, f'ok{rank}'), 'w').write(str(bool(ok)))
    dist.destroy_process_group()


def test_sample_and_hypothesis_sharding_host_logic_world2(tmp_path):
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f'ok{r}').read() for r in range(2)] == ['True', 'True']
