"""Pose estimators driving the render path.  API mirror of reference
``latentfusion/pose/estimation.py`` (load_from_config :29-59, default_pose_loss :70-118,
PoseEstimator :129-216, CrossEntropyPoseEstimator :298-497, GradientPoseEstimator :500-713); the
TOML configs under the reference's ``configs/`` load unchanged.

The host loops stay Python, like the reference.  Per iteration the heavy work is
``model.render_latent_object`` (fused sm_100a kernels) and its backward to the 10 camera floats.
"""
import copy
import math
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch
from torch import optim
from torch.nn import functional as F

from .. import three, utils
from ..modules.geometry import Camera
from ..observation import Observation  # noqa: F401  (re-exported for API parity)
from . import initialization
from . import utils as pu

DEFAULT_TRANSLATION_STD = 0.01
DEFAULT_QUATERION_STD = 10.0 / 180.0 * math.pi


def load_from_config(config, model, **kwargs):
    if isinstance(config, (Path, str)):
        import toml
        config = toml.load(config)
    params = dict(config['args'])
    params.update(kwargs)
    kind = config['type']
    if kind == 'cross_entropy':
        return CrossEntropyPoseEstimator(model=model, **params, loss_weights=config['loss_weights'])
    if kind == 'gradient':
        schedules = {k: load_schedules_from_config(v) for k, v in config.get('loss_schedules', {}).items()}
        return GradientPoseEstimator(model=model, **params, loss_weights=config['loss_weights'],
                                     loss_schedules=schedules)
    if kind == 'metropolis':
        raise NotImplementedError("the Metropolis estimator is not part of the accelerated path")
    raise ValueError(f"Unknown estimator type {kind}")


def load_schedules_from_config(config):
    config = dict(config)
    kind = config.pop('type')
    if kind == 'exponential':
        return utils.ExponentialScheduler(**config)
    if kind == 'linear':
        return utils.LinearScheduler(**config)
    raise ValueError(f"Unknown schedule type {kind}")


def cosine_distance(x1, x2, dim=1, eps=1e-8):
    return 1.0 - torch.cosine_similarity(x1, x2, 0 if x1.dim() == 1 else dim, eps)


def default_pose_loss(target, z_pred_depth, z_pred_mask_logits, z_pred_camera, z_pred_latent=None,
                      z_target_latent=None):
    """Per-hypothesis fitness terms of rendered crops against ONE full-frame target observation.

    The crops are pasted back into the sensor frame first (depth: nearest, mask logits: bilinear; border-clamped).
    Pixels the target mask claims but the depth sensor missed (depth == 0) are excluded from the depth and IoU terms.
    Returns a dict of [N] tensors: 'depth' (mean L1 over the frame), 'ov_depth' (L1 averaged over the overlap of the
    predicted and target masks), 'iou', 'mask' (BCE with logits) and, when both latents are given, 'latent'
    (cosine distance of the flattened pre-decoder features).  Same arithmetic as reference estimation.py:70-118.
    """
    frame_depth = z_pred_camera.uncrop(z_pred_depth, scale_mode='nearest')[0]
    frame_logits = z_pred_camera.uncrop(z_pred_mask_logits, scale_mode='bilinear')[0]
    frame_mask = torch.sigmoid(frame_logits)
    frame_depth = frame_depth * frame_mask
    sensor_hole = (target.depth == 0) & (target.mask > 0.1)
    target = target.prepare()
    per_sample = (1, 2, 3)

    abs_err = F.l1_loss(frame_depth, target.depth.expand_as(frame_depth), reduction='none')
    abs_err = pu.zero_invalid_pixels(abs_err, sensor_hole)
    terms = {
        'ov_depth': pu.reduce_loss_mask(abs_err, frame_mask * target.mask),
        'depth': abs_err.mean(dim=per_sample),
        'iou': pu.iou_loss(frame_mask, pu.zero_invalid_pixels(target.mask, sensor_hole)),
        'mask': F.binary_cross_entropy_with_logits(frame_logits, target.mask.expand_as(frame_mask),
                                                   reduction='none').mean(dim=per_sample),
    }
    if z_pred_latent is not None and z_target_latent is not None:
        rendered = z_pred_latent.flatten(1)
        terms['latent'] = cosine_distance(rendered, z_target_latent.flatten(1).expand_as(rendered))
    return terms


def weigh_losses(loss_dict, weight_dict):
    return {k: weight_dict.get(k, 0.0) * v for k, v in loss_dict.items()}


class PoseEstimator:

    def __init__(self, *, model, ranking_size, loss_weights, loss_func=None, return_camera_history=False,
                 verbose=False):
        self.model, self.ranking_size = model, ranking_size
        self.loss_func = loss_func or default_pose_loss
        self.loss_weights = defaultdict(float, loss_weights)          # unspecified terms weigh 0
        self.return_camera_history, self.verbose = return_camera_history, verbose

    device = property(lambda self: self.model.device)
    group = None

    @classmethod
    def initial_pose(cls, target_obs):
        """Centroid-of-the-masked-depth initial guess (pose/initialization.py)."""
        cam = target_obs.camera
        return initialization.estimate_initial_pose(target_obs.depth, target_obs.mask, cam.intrinsic, cam.width,
                                                    cam.height)

    def estimate(self, z_obj, target_obs, **kwargs):
        """`group=` (a torch.distributed process group, or True for the default one) shards the hypotheses / samples
        over its ranks: the cross-entropy search scores a slice of every generation per rank and all-gathers the
        scores; the gradient refinement optimises a slice of the hypotheses per rank and merges the rankings at the
        end (dist.merge_rankings).  Every rank returns the same cameras."""
        if len(target_obs) > 1:
            raise ValueError("The pose can only be estiamted for one observation at a time.")
        group = kwargs.pop('group', None)
        if group is True:
            import torch.distributed as dist
            group = dist.group.WORLD if dist.is_initialized() else None
        self.group = group
        try:
            return self._estimate(z_obj, target_obs, **kwargs)
        finally:
            self.group = None

    def _estimate(self, z_obj, target_obs, **kwargs):
        raise NotImplementedError()

    def _track_best_items(self, ranking, step, items, loss):
        """Merge this step's (item, loss) pairs into the running top-`ranking_size` list; returns how
        much the best loss improved."""
        before = ranking[0][1] if ranking else float('inf')
        for item, value in zip(items, loss.detach().cpu().tolist()):
            ranking.append((item, value, step))
        ranking.sort(key=lambda entry: entry[1])
        del ranking[self.ranking_size:]
        return max(before - ranking[0][1], 0.0)

    def _render_observation(self, z_obj, camera, **kwargs):
        """Render full-frame cameras through their zoomed crops -> (masked metric depth, mask logits, latent, crop camera)."""
        crop = camera.zoom(None, self.model.input_size, self.model.camera_dist)
        with torch.set_grad_enabled(bool(kwargs.get('grad_enabled', False))):
            out, latent = self.model.render_latent_object(z_obj, crop.to(self.device), return_latent=True)
            mask, mask_logits, depth = (out[k].squeeze(0) for k in ('mask', 'mask_logits', 'depth'))
            metric_depth = camera.denormalize_depth(depth) * mask
        return metric_depth, mask_logits, latent, crop


class CrossEntropyPoseEstimator(PoseEstimator):
    """Cross-entropy method over (translation, log-quaternion) with a diagonal GMM proposal
    (sklearn, host side); rendering the samples is forward-only."""

    def __init__(self, *, num_samples, num_elites, num_iters, num_gmm_components, learning_rate,
                 sample_flipped=False, init_hemisphere=False, init_upright=False,
                 translation_std=DEFAULT_TRANSLATION_STD, quaternion_std=DEFAULT_QUATERION_STD, **kwargs):
        super().__init__(**kwargs)
        self.num_samples, self.num_elites, self.num_iters = num_samples, num_elites, num_iters
        self.num_gmm_components, self.learning_rate = num_gmm_components, learning_rate
        self.sample_flipped, self.init_upright, self.init_hemisphere = sample_flipped, init_upright, init_hemisphere
        self.translation_std, self.quaternion_std = translation_std, quaternion_std
        self.elite_sched = utils.ExponentialScheduler(num_samples, num_elites, num_iters)

    _FLIP_AXES = ((0.0, 0.0, 1.0), (0.0, 1.0, 0.0), (1.0, 0.0, 0.0))

    def _seed_population(self, target_obs, cameras):
        """(reference camera for intrinsics/frame, initial population)"""
        if cameras:
            return cameras[0], cameras
        guess = self.initial_pose(target_obs)
        spread = pu.sample_cameras_with_estimate(n=self.num_gmm_components * self.num_samples, camera_est=guess,
                                                 upright=self.init_upright, hemisphere=self.init_hemisphere)
        return guess, spread

    def _fit(self, cameras):
        return self._create_gmm(self._camera_to_params(cameras).cpu())

    def _estimate(self, z_obj, target_obs, **kwargs):
        camera_init, population = self._seed_population(target_obs, kwargs.get('cameras'))
        target_obs = target_obs.to(self.device)
        proposal, previous = self._fit(population), None
        ranking, history = [], []
        for step in utils.trange(self.num_iters):
            keep = int(self.elite_sched.get(step))
            elites, losses = self._refine_pose(z_obj, target_obs, previous, proposal, num_elites=keep,
                                               camera_init=camera_init)
            previous, proposal = proposal, self._fit(elites)
            improved = self._track_best_items(ranking, step, elites, losses)
            if improved > 0:
                history.append((losses, Camera.cat([entry[0] for entry in ranking])))
        winners = Camera.cat([entry[0] for entry in ranking])
        return (winners, history) if self.return_camera_history else winners

    def _refine_pose(self, z_obj, target_obs, prev_gmm, gmm, num_elites, camera_init):
        """One CE generation: sample from the (blended) proposal, score every sample with a forward render, keep the
        `num_elites` best."""
        mixture = gmm if prev_gmm is None else self._combined_gmm(prev_gmm, gmm, self.learning_rate)
        draws = self.num_samples // (1 + len(self._FLIP_AXES)) if self.sample_flipped else self.num_samples
        cameras = self._params_to_camera(self._sample_poses(mixture, draws), camera_init=camera_init, device=self.device)
        if self.sample_flipped:                  # symmetric objects: also try each sample turned by pi about x, y, z
            cameras = Camera.cat([cameras, *(pu.flip_camera(cameras, axis=axis) for axis in self._FLIP_AXES)])
        target_code = None
        if self.loss_weights.get('latent', 0.0) > 0.0:
            with torch.no_grad():
                target_code = self.model.compute_latent_code(target_obs, cameras[0])
        # samples shard contiguously over the ranks of `group` (every rank drew the same population: same seed /
        # broadcast GMM); the [n] scores are all-gathered and every rank picks the same elites
        lo, hi = self._shard(len(cameras))
        score = self._score(z_obj, target_obs, cameras[lo:hi] if (lo, hi) != (0, len(cameras)) else cameras, target_code)
        score = self._gather_scores(score, len(cameras))
        best = torch.argsort(score)[:num_elites]
        return cameras[best], score[best]

    group = None            # torch.distributed process group for sample sharding (None: not sharded)

    def _shard(self, n):
        import torch.distributed as dist
        if self.group is None or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return 0, n
        from .. import dist as lfdist
        return lfdist.shard_range(n, dist.get_rank(self.group), dist.get_world_size(self.group))

    def _gather_scores(self, score, n):
        import torch.distributed as dist
        if score.shape[0] == n:
            return score
        from .. import dist as lfdist
        world = dist.get_world_size(self.group)
        sizes = [lfdist.shard_range(n, r, world)[1] - lfdist.shard_range(n, r, world)[0] for r in range(world)]
        return lfdist.all_gather_ragged(score.reshape(-1, 1), sizes, self.group).reshape(-1)

    def _score(self, z_obj, target_obs, cameras, target_code):
        """weighted fitness of each sample, forward only.  On the device path the crops never become full frames:
        the raw head outputs go through the fused loss head (csrc/pose_loss.cu, the kernels the graphed refiner
        uses), the latent term is a cosine distance of the projected features."""
        ph = self.model.photographer
        on_device = (self.loss_func is default_pose_loss and target_obs.depth.is_cuda and ph.predict_depth
                     and ph.predict_mask and not ph.predict_color)
        if not on_device:
            depth, mask_logits, code, crop = self._render_observation(z_obj, cameras)
            terms = self.loss_func(target_obs, depth, mask_logits, crop, z_pred_latent=code, z_target_latent=target_code)
            return sum(weigh_losses(terms, self.loss_weights).values())
        from .. import ops
        crops = cameras.zoom(None, self.model.input_size, self.model.camera_dist).to(self.device)
        scores = []
        with torch.no_grad():
            for lo in range(0, len(crops), self.render_chunk):          # bounds the frustum-volume working set
                crop = crops[lo:lo + self.render_chunk]
                logits, latent, _ = ph.decode(z_obj, crop, interpret_logits=False, return_latent=True)
                t = ops.pose_search_terms(logits[:, 0], logits[:, 1], crop.viewport, crop.translation[:, 2],
                                          target_obs.depth, target_obs.mask, crop.z_span, 0.01, crop.width, crop.height)
                terms = {'ov_depth': t[:, 0], 'depth': t[:, 1], 'iou': t[:, 2], 'mask': t[:, 3]}
                if target_code is not None:
                    rendered = latent.squeeze(0).flatten(1)
                    terms['latent'] = cosine_distance(rendered, target_code.flatten(1).expand_as(rendered))
                scores.append(sum(weigh_losses(terms, self.loss_weights).values()))
        return torch.cat(scores)

    render_chunk = 32

    def _sample_poses(self, gmm, n):
        """n draws (translation | log-quaternion) from the proposal, jittered so elites never collapse to a point"""
        drawn = torch.as_tensor(gmm.sample(n)[0], dtype=torch.float32, device=self.device)
        jitter = torch.cat((torch.full((3,), float(self.translation_std)), torch.full((3,), float(self.quaternion_std))))
        return drawn + torch.randn_like(drawn) * jitter.to(drawn.device)

    def _create_gmm(self, params=None):
        import sklearn.mixture
        gmm = sklearn.mixture.GaussianMixture(covariance_type='diag', n_components=self.num_gmm_components,
                                              reg_covar=1e-5)
        if params is not None:
            if torch.is_tensor(params):
                params = params.detach().cpu().numpy()
            gmm.fit(np.asarray(params, dtype=np.float64))     # float64: current sklearn rejects degenerate f32 fits
        return gmm

    def _combined_gmm(self, old_gmm, new_gmm, alpha):
        if not 0.0 <= alpha <= 1.0:
            raise ValueError("alpha must be between 0.0 and 1.0")
        out = self._create_gmm()
        out.weights_ = np.concatenate([(1.0 - alpha) * old_gmm.weights_, alpha * new_gmm.weights_], axis=0)
        for name in ('means_', 'covariances_', 'precisions_cholesky_'):
            setattr(out, name, np.concatenate([getattr(old_gmm, name), getattr(new_gmm, name)], axis=0))
        return out

    @classmethod
    def _camera_to_params(cls, camera):
        return torch.cat([camera.translation, camera.log_quaternion], dim=-1)

    @classmethod
    def _params_to_camera(cls, params, camera_init, device='cpu'):
        if params.dim() == 1:
            params = params.unsqueeze(0)
        return Camera(intrinsic=camera_init.intrinsic.expand(params.shape[0], -1, -1).to(device), extrinsic=None,
                      translation=params[:, :3].to(device), log_quaternion=params[:, 3:].to(device),
                      width=camera_init.width, height=camera_init.height, z_span=camera_init.z_span).to(device)


class GradientPoseEstimator(PoseEstimator):
    """First-order refinement of N pose hypotheses: each hypothesis has its own optimiser and
    ReduceLROnPlateau schedule over (log-quaternion, translation, viewport)."""

    def __init__(self, *, learning_rate, num_samples, num_iters, converge_threshold, converge_patience,
                 lr_reduce_patience=25, lr_reduce_threshold=1e-5, lr_reduce_factor=0.5, track_stats=False,
                 loss_schedules=None, optimizer='adamw', cuda_graph=True, graph_chunk=16, fused_loss=True, **kwargs):
        super().__init__(**kwargs)
        # B200 path: capture one loop body as a CUDA graph (pose/refine_graph.py).  Falls back to the eager
        # loop for configurations the graphed step does not cover (non-Adam optimisers, latent loss, custom
        # loss functions, CPU tensors).
        self.cuda_graph, self.graph_chunk, self.fused_loss = cuda_graph, graph_chunk, fused_loss
        self.learning_rate, self.num_samples, self.num_iters = learning_rate, num_samples, num_iters
        self.optimizer = optimizer
        self.lr_reduce_patience, self.lr_reduce_threshold = lr_reduce_patience, lr_reduce_threshold
        self.lr_reduce_factor = lr_reduce_factor
        self.converge_threshold, self.converge_patience = converge_threshold, converge_patience
        self.loss_schedules = dict(loss_schedules or {})
        self.track_stats = track_stats

    def _estimate(self, z_obj, target_obs, **kwargs):
        if 'camera' in kwargs:
            camera = kwargs['camera']
        else:
            camera = pu.sample_cameras_with_estimate(n=self.num_samples, camera_est=self.initial_pose(target_obs))
        target_obs = target_obs.to(self.device)
        # the *zoomed* camera (viewport box around the object) is what gets optimised
        camera = camera.zoom(None, self.model.input_size, self.model.camera_dist).to(self.device)
        sharded = self._sharded()
        if sharded:
            from .. import dist as lfdist
            lo, hi = lfdist.shard_range(len(camera), sharded[0], sharded[1])
            camera = camera[lo:hi]
        ranking = []
        stats, history = self._optimize_camera(z_obj, target_obs, camera, iters=self.num_iters, ranking=ranking)
        best = Camera.cat([c for c, _, _ in ranking])
        if sharded:
            best = self._merge_ranked(best, torch.tensor([loss for _, loss, _ in ranking]))
        if self.track_stats and self.return_camera_history:
            return best, stats, history
        if self.track_stats:
            return best, stats
        if self.return_camera_history:
            return best, history
        return best

    def _sharded(self):
        import torch.distributed as dist
        if self.group is None or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return None
        return dist.get_rank(self.group), dist.get_world_size(self.group)

    def _merge_ranked(self, cameras, losses):
        """global top-`ranking_size` over all ranks' local rankings (tiny all-gather of losses + the 6 pose floats)"""
        from .. import dist as lfdist
        dev = self.device
        params = torch.cat([cameras.translation, cameras.log_quaternion], dim=-1).to(dev)
        top_l, top_p, _ = lfdist.merge_rankings(losses.to(dev), params, self.ranking_size, self.group)
        top_p = top_p.cpu()
        k = top_p.shape[0]
        return Camera(cameras.intrinsic[:1].cpu().expand(k, -1, -1).contiguous(), None, cameras.z_span, None,
                      width=cameras.width, height=cameras.height, log_quaternion=top_p[:, 3:].contiguous(),
                      translation=top_p[:, :3].contiguous())

    @classmethod
    def get_optimizer(cls, name, *args, **kwargs):
        table = {'adamw': optim.AdamW, 'adam': optim.Adam, 'sgd': optim.SGD, 'adagrad': optim.Adagrad}
        if name not in table:
            raise ValueError(f"Unknow optimizer {name!r}")
        return table[name](*args, **kwargs)

    def _can_graph(self, cameras):
        return (self.cuda_graph and self.optimizer == 'adam' and self.loss_func is default_pose_loss
                and cameras.device.type == 'cuda')

    def _optimize_camera_graphed(self, z_obj, target_obs, cameras, iters, ranking):
        from .refine_graph import GraphedRefiner
        refiner = getattr(self, '_refiner', None)
        probe = GraphedRefiner.make_signature(self, z_obj, target_obs, cameras, self.graph_chunk)
        if refiner is not None and refiner.signature() == probe:
            refiner.reset(z_obj, target_obs, cameras)           # same shapes: reuse the captured graph
        else:
            refiner = GraphedRefiner(self, z_obj, target_obs, cameras, chunk=self.graph_chunk)
            refiner.capture()
            self._refiner = refiner
        stats, history, converge_count = {}, [], 0
        gt_cam = target_obs.camera
        gt_quat, gt_trans = gt_cam.quaternion.detach().cpu(), gt_cam.translation.detach().cpu()
        step, done = 0, False
        while step < iters and not done:
            count = min(refiner.chunk, iters - step)
            hist = refiner.run_chunk(step, count)
            for i in range(count):
                rank_loss = hist['rank'][i]
                snapshot = refiner.camera_at(hist['lq'][i], hist['tr'][i])
                if self.return_camera_history:
                    history.append((rank_loss, snapshot))
                delta = self._track_best_items(ranking, step, items=snapshot, loss=rank_loss)
                if self.track_stats:
                    angle = three.quaternion.angular_distance(snapshot.quaternion, gt_quat).squeeze()
                    trans = torch.norm(snapshot.translation - gt_trans, dim=1).squeeze()
                    weights = copy.copy(self.loss_weights)
                    weights.update({k: s.get(step) for k, s in self.loss_schedules.items()})
                    self._record_stat_dict(stats, {
                        **{f'{k}_loss': hist['terms'][i][j] for j, k in enumerate(refiner.TERMS)},
                        **{f'{k}_weight': v for k, v in weights.items()},
                        'delta': delta, 'converge_count': converge_count, 'angle_dist': angle,
                        'trans_dist': trans, 'optim_loss': hist['optim'][i], 'rank_loss': rank_loss})
                if delta < self.converge_threshold:
                    converge_count += 1
                elif delta > self.converge_threshold:
                    converge_count = 0
                step += 1
                if converge_count >= self.converge_patience:
                    done = True          # iterations already replayed past this point are discarded
                    break
        return stats, history

    def _optimize_camera(self, z_obj, target_obs, cameras, iters, ranking):
        if self._can_graph(cameras):
            return self._optimize_camera_graphed(z_obj, target_obs, cameras, iters, ranking)
        params = [pu.parameterize_camera(c, optimize_viewport=True) for c in cameras]
        optimizers, schedulers = [], []
        for cam in params:
            opt = self.get_optimizer(self.optimizer, [cam.log_quaternion, cam.translation, cam.viewport],
                                     lr=self.learning_rate)
            optimizers.append(opt)
            schedulers.append(optim.lr_scheduler.ReduceLROnPlateau(
                opt, patience=self.lr_reduce_patience, threshold=self.lr_reduce_threshold,
                factor=self.lr_reduce_factor))
        stats, history, converge_count = {}, [], 0
        for step in utils.trange(iters):
            for opt in optimizers:
                opt.zero_grad()
            cameras = Camera.cat(params)
            z_target_latent = None
            if self.loss_weights.get('latent', 0.0) > 0.0:
                with torch.no_grad():
                    z_target_latent = self.model.compute_latent_code(target_obs, cameras)
            z_depth, z_mask, z_mask_logits, z_latent = self._render_observation(z_obj, cameras)
            weights = copy.copy(self.loss_weights)
            weights.update({k: s.get(step) for k, s in self.loss_schedules.items()})
            loss_dict = self.loss_func(target_obs, z_depth, z_mask_logits, cameras, z_pred_latent=z_latent,
                                       z_target_latent=z_target_latent)
            optim_loss = sum(weigh_losses(loss_dict, weights).values())
            optim_loss.mean().backward()
            rank_loss = sum(weigh_losses(loss_dict, self.loss_weights).values()).detach()

            snapshot = pu.deparameterize_camera(cameras.uncrop()).clone()
            if self.track_stats:     # before snapshot.cpu(): nn.Module.cpu() moves the camera in place
                angle = three.quaternion.angular_distance(snapshot.quaternion, target_obs.camera.quaternion).squeeze()
                trans = torch.norm(snapshot.translation - target_obs.camera.translation, dim=1).squeeze()
            if self.return_camera_history:
                history.append((rank_loss.cpu(), snapshot.cpu()))
            delta = self._track_best_items(ranking, step, items=snapshot.cpu(), loss=rank_loss)

            if self.track_stats:
                self._record_stat_dict(stats, {
                    **{f'{k}_loss': v.detach().cpu() for k, v in loss_dict.items()},
                    **{f'{k}_weight': v for k, v in weights.items()},
                    'delta': delta, 'converge_count': converge_count, 'angle_dist': angle.cpu(),
                    'trans_dist': trans.cpu(), 'optim_loss': optim_loss.detach().cpu(),
                    'rank_loss': rank_loss.cpu()})

            rank_host = rank_loss.cpu()
            for i, (opt, sched) in enumerate(zip(optimizers, schedulers)):
                opt.step()
                sched.step(rank_host[i])

            if delta < self.converge_threshold:
                converge_count += 1
            elif delta > self.converge_threshold:
                converge_count = 0
            if converge_count >= self.converge_patience:
                break
        return stats, history

    @classmethod
    def _record_stat(cls, history, key, value):
        value = value.detach().cpu() if torch.is_tensor(value) else torch.tensor(value)
        value = value.squeeze().unsqueeze(0)
        if value.dim() > 2:
            for i in range(value.shape[-1]):
                cls._record_stat(history, f'{key}[{i}]', value[..., i])
        else:
            history[key] = torch.cat((history[key], value), dim=0) if key in history else value

    @classmethod
    def _record_stat_dict(cls, history, entries):
        for name in entries:
            cls._record_stat(history, name, entries[name])

    def _render_observation(self, z_obj, camera, **kwargs):
        """The optimised camera is already the zoomed one, so render it as is."""
        out, latent = self.model.render_latent_object(z_obj, camera.to(self.model.device), return_latent=True)
        mask, mask_logits, depth = (out[k].squeeze(0) for k in ('mask', 'mask_logits', 'depth'))
        return camera.denormalize_depth(depth), mask, mask_logits, latent
