"""BASELINE INFRASTRUCTURE — runs the UNMODIFIED reference's own pose-refinement loop for bench.py's reference arm.

`GradientPoseEstimator.estimate()` of the reference (latentfusion/pose/estimation.py:532-677, configs/adam_quick.toml)
on the synthetic configs[1] workload (LF-synth(64, 32), N hypotheses, 128^2 render): every iteration is the real
thing — Camera.cat -> render_latent_object -> default_pose_loss -> backward -> N Adam steps + N plateau schedulers ->
ranking.  Imported from /root/reference in the authoring container, from the verbatim copy oracle/_ref/ on the GPU
box (oracle/stage_ref.py).  Nothing here is product code; only bench.py's `--impl reference` arm and its
`reference_cuda` context block call it."""
import math
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_import  # noqa: E402


def available():
    return ref_import.available()


def build_case(device, n_hyp, S=64, C=32, num_iters=1):
    """(estimator, z_obj, target, hypothesis cameras) built from the reference's own classes, random weights (seed 0)."""
    ref_import.install()
    import torch
    from latentfusion import consts, three
    from latentfusion.modules.geometry import Camera
    from latentfusion.observation import Observation
    from latentfusion.pose import estimation as ref_estimation
    from latentfusion.pose import utils as ref_pu
    from latentfusion.recon import fusion as ref_fusion
    from latentfusion.recon import models as ref_models
    from latentfusion.recon.inference import LatentFusionModel
    from latentfusion.recon.utils import optimal_camera_dist
    from latentfusion.utils import parse_block_config as pbc

    torch.manual_seed(0)
    arch_p = dict(in_size=S, image_config=pbc(f"{C},D,{2*C}:{2*C},U,{2*C},U,{C}"), camera_config=pbc(f"{C},{C}"),
                  object_config=[], projection_type='factor', predict_depth=True, predict_mask=True,
                  predict_color=False, scale_mode='nearest')
    arch_s = dict(in_size=2 * S, image_config=pbc(f"{C},D,{2*C}:{2*C},{2*C}"), camera_config=pbc(f"{C},{C}"),
                  object_config=pbc(f"{C},{C}"), projection_type='factor', input_color=True, input_depth=False,
                  input_mask=True, scale_mode='nearest')
    photographer = ref_models.Photographer(**arch_p)
    sculptor = ref_models.Sculptor(**arch_s)
    fuser = ref_fusion.get_fuser('pool:mean', C, 1.0)
    dist = optimal_camera_dist(615.4991, 2 * S, 0.5, slack=128 / (2 * S))
    model = LatentFusionModel(sculptor, fuser, photographer, dist, device)
    cfg = {'type': 'gradient',
           'args': dict(optimizer='adam', num_iters=num_iters, num_samples=n_hyp, ranking_size=n_hyp, learning_rate=0.01,
                        lr_reduce_patience=10, lr_reduce_threshold=1e-4, converge_threshold=1e-6,
                        converge_patience=10 ** 9),
           'loss_weights': dict(depth=1.0, ov_depth=0.3, iou=0.0, mask=0.0, latent=0.0)}      # configs/adam_quick.toml
    est = ref_estimation.load_from_config(cfg, model)
    torch.manual_seed(2)
    K = torch.tensor(consts.INTRINSIC).unsqueeze(0)
    gt = Camera(K, three.to_extrinsic_matrix(torch.tensor([[0.0, 0.0, dist]]), three.orientation.evenly_distributed_quats(1)),
                z_span=0.5, width=640, height=480)
    torch.manual_seed(7)
    hyp = Camera.cat([ref_pu.perturb_camera(gt, 0.01, 10.0 / 180.0 * math.pi) for _ in range(n_hyp)])
    yy, xx = torch.meshgrid(torch.arange(480, dtype=torch.float32), torch.arange(640, dtype=torch.float32), indexing='ij')
    tmask = (((yy - 251.5) ** 2 + (xx - 315.4) ** 2) <= 45.0 ** 2).float().view(1, 1, 480, 640)
    target = Observation(torch.zeros(1, 3, 480, 640), tmask * dist, tmask, gt)
    torch.manual_seed(5)
    z_obj = (torch.randn(1, 1, C, S, S, S) * 0.5).to(device)       # the cube's values do not change the cost
    return est, z_obj, target, hyp


def time_iterations(device, n_hyp, warmup, steps, tf32=False, threads=None):
    """seconds per iteration of the reference's own estimate() loop (num_iters = steps), after `warmup` iterations"""
    import torch
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    if threads:
        torch.set_num_threads(threads)
    est, z_obj, target, hyp = build_case(device, n_hyp)
    dev = torch.device(device)

    def run(iters):
        est.num_iters = iters
        if dev.type == 'cuda':
            torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        est.estimate(z_obj, target, camera=hyp.clone())
        if dev.type == 'cuda':
            torch.cuda.synchronize(dev)
        return time.perf_counter() - t0

    if warmup > 0:
        run(warmup)
    return run(steps) / steps
