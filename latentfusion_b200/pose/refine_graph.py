"""CUDA-graph execution of the pose-refinement iteration.

One body of the reference's ``GradientPoseEstimator._optimize_camera`` loop (pose/estimation.py:601-677)
— camera assembly, ``render_latent_object`` forward, ``default_pose_loss``, backward to the 10 camera
floats of each hypothesis, Adam step, ReduceLROnPlateau step — is expressed as device-only work (no
``.item()``/``.cpu()`` inside), captured once into a ``torch.cuda.CUDAGraph`` and replayed.  The
reference issues ~300 kernel launches and ~10 host synchronisations per iteration from Python; a replay
is one ``cudaGraphLaunch``.

Semantics kept identical to the reference's N independent optimisers:
* Adam is elementwise, so N ``optim.Adam`` instances over ([3],[3],[4]) tensors equal one update over
  ([N,3],[N,3],[N,4]) with a per-row learning rate (``torch/optim/adam.py`` single-tensor maths);
* ``ReduceLROnPlateau(mode='min', threshold_mode='rel', cooldown=0, min_lr=0, eps=1e-8)`` is evaluated per
  row on device with the same update order (optimizer step, then scheduler step on this iteration's loss);
* per-iteration snapshots (ranking losses, loss terms, detached camera parameters) are written to device
  history buffers and drained every ``chunk`` iterations, where the host replays the reference's ranking /
  convergence bookkeeping iteration by iteration (extra iterations past convergence are discarded).
"""
import copy

import torch

from .. import ops
from ..modules.geometry import Camera


class _BatchedAdamPlateau:
    """Adam + ReduceLROnPlateau for N independent hypotheses, state on device, capturable."""

    def __init__(self, params, n, lr, patience, threshold, factor, betas=(0.9, 0.999), eps=1e-8):
        dev = params[0].device
        self.params = params
        self.b1, self.b2, self.eps = betas[0], betas[1], eps
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.step_count = torch.zeros(1, device=dev)
        self.lr = torch.full((n, 1), float(lr), device=dev)
        self.best = torch.full((n,), float('inf'), device=dev)
        self.num_bad = torch.zeros(n, device=dev)
        self.patience, self.threshold, self.factor = float(patience), float(threshold), float(factor)

    @torch.no_grad()
    def step(self, rank_loss, counted=False):
        """counted: the step counter was already advanced on the device (ops.refine_record_)"""
        if self.params[0].is_cuda:
            if not counted:
                self.step_count += 1
            for p, m, v in zip(self.params, self.m, self.v):
                ops.adam_step_(p, p.grad, m, v, self.step_count, self.lr, self.b1, self.b2, self.eps)
            ops.plateau_step_(rank_loss, self.lr, self.best, self.num_bad, self.threshold, self.patience, self.factor)
            return
        self.step_count += 1
        bc1 = 1 - self.b1 ** self.step_count
        bc2_sqrt = (1 - self.b2 ** self.step_count).sqrt()
        for p, m, v in zip(self.params, self.m, self.v):
            g = p.grad
            m.lerp_(g, 1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (v.sqrt() / bc2_sqrt).add_(self.eps)
            p.addcdiv_(m, denom * (bc1 / self.lr), value=-1.0)      # p -= (lr / bc1) * m / denom
        # ReduceLROnPlateau, mode='min', threshold_mode='rel'
        better = rank_loss < self.best * (1.0 - self.threshold)
        self.best.copy_(torch.where(better, rank_loss, self.best))
        self.num_bad.copy_(torch.where(better, torch.zeros_like(self.num_bad), self.num_bad + 1))
        reduce = self.num_bad > self.patience
        new_lr = self.lr * self.factor
        apply = reduce.unsqueeze(1) & ((self.lr - new_lr) > 1e-8)
        self.lr.copy_(torch.where(apply, new_lr, self.lr))
        self.num_bad.copy_(torch.where(reduce, torch.zeros_like(self.num_bad), self.num_bad))


class GraphedRefiner:
    """Owns the static tensors, the captured graph and the device-side history of one refinement run."""

    TERMS = ('ov_depth', 'depth', 'iou', 'mask')            # + 'latent' on instances whose loss weights use it

    def __init__(self, estimator, z_obj, target_obs, cameras, chunk=16):
        self.est = estimator
        self.model = estimator.model
        dev = self.model.device
        self.n = n = len(cameras)
        self.chunk = chunk
        # static copies: the captured graph is bound to these addresses; reset() refills them for a new run
        self.z_obj = z_obj.detach().clone()
        self.target = target_obs.clone()
        self.template = cameras.detach().clone()                  # intrinsics / sizes
        self.lq = cameras.log_quaternion.detach().clone().requires_grad_(True)
        self.tr = cameras.translation.detach().clone().requires_grad_(True)
        self.vp = cameras.viewport.detach().clone().requires_grad_(True)
        self.launches_per_iteration = 0
        self.opt = _BatchedAdamPlateau([self.lq, self.tr, self.vp], n, estimator.learning_rate,
                                       estimator.lr_reduce_patience, estimator.lr_reduce_threshold,
                                       estimator.lr_reduce_factor)
        self.weights = dict(estimator.loss_weights)
        # latent loss (configs/adam_latent.toml; reference estimation.py:605-609): the target code is the autoencoding of
        # the PREPROCESSED target observation through the current hypothesis cameras, recomputed every iteration without
        # grad.  The preprocessing depends on the observation only, so it is hoisted out of the captured loop.
        self.use_latent = float(self.weights.get('latent', 0.0)) > 0.0
        self.target_inputs = None
        if self.use_latent:
            self.TERMS = GraphedRefiner.TERMS + ('latent',)
            self._set_target_inputs(target_obs, n)
        # the four fused-head terms with static / scheduled weights: combined, differentiated and recorded by one kernel
        # (ops.refine_record_).  Scheduled weights are 0-dim views into w_opt, refilled from the host before a replay.
        ph = self.model.photographer
        self.fast = bool(estimator.fused_loss and ph.predict_depth and ph.predict_mask and not ph.predict_color
                         and not self.use_latent and z_obj.is_cuda)
        self.w_rank = torch.tensor([float(self.weights.get(k, 0.0)) for k in self.TERMS[:4]], device=dev)
        self.w_opt = self.w_rank.clone()
        self.rank_buf = torch.zeros(n, device=dev)
        self.gterms = torch.zeros(n, 4, device=dev)
        if self.fast:
            self.sched_w = {k: self.w_opt[self.TERMS.index(k)] for k in estimator.loss_schedules if k in self.TERMS[:4]}
        else:
            self.sched_w = {k: torch.tensor(float(self.weights.get(k, 0.0)), device=dev) for k in estimator.loss_schedules}
        # device history (one chunk)
        self.h_rank = torch.zeros(chunk, n, device=dev)
        self.h_optim = torch.zeros(chunk, n, device=dev)
        self.h_terms = torch.zeros(chunk, len(self.TERMS), n, device=dev)
        self.h_lq = torch.zeros(chunk, n, 3, device=dev)
        self.h_tr = torch.zeros(chunk, n, 3, device=dev)
        self.slot = torch.zeros(1, dtype=torch.long, device=dev)
        self.graph = None
        self._signature = self.make_signature(estimator, z_obj, target_obs, cameras, chunk)

    def _set_target_inputs(self, target_obs, n):
        obs = self.model.preprocess_observation(target_obs).to(self.model.device)
        if len(obs) == 1:
            obs = obs.expand(n)
        fresh = {k: v.detach().clone() for k, v in self.model._inputs(obs, 1).items()}
        if self.target_inputs is None:
            self.target_inputs = fresh
        else:                                   # keep the addresses the captured graph reads
            for k, v in fresh.items():
                self.target_inputs[k].copy_(v)

    def _target_code(self, cam):
        from ..recon import models
        with torch.no_grad():
            return models.autoencode(self.model.sculptor, self.model.fuser, self.model.photographer, camera=cam.detach(),
                                     **self.target_inputs)[1]

    # ---- one iteration, device only ----
    def _camera(self):
        t = self.template
        return Camera(t.intrinsic, None, t.z_span, self.vp, width=t.width, height=t.height,
                      log_quaternion=self.lq, translation=self.tr)

    def _iteration(self):
        est = self.est
        for p in (self.lq, self.tr, self.vp):
            p.grad = None
        cam = self._camera()
        ph = self.model.photographer
        if self.fast:
            logits, _, _ = ph.decode(self.z_obj, cam, interpret_logits=False, return_latent=False)
            terms = ops.pose_loss_terms_packed(logits, cam.viewport, cam.translation, self.target.depth, self.target.mask,
                                               cam.z_span, 0.01, cam.width, cam.height)
            with torch.no_grad():
                # rank / optim, d mean(optim)/d terms, the history snapshot (BEFORE the update: the reference ranks the
                # cameras that produced this loss) and the optimiser's step counter
                ops.refine_record_(terms, self.w_rank, self.w_opt, self.lq, self.tr, self.rank_buf, self.gterms,
                                   self.h_rank, self.h_optim, self.h_terms, self.h_lq, self.h_tr, self.slot, self.chunk,
                                   self.opt.step_count)
            terms.backward(self.gterms)
            self.opt.step(self.rank_buf, counted=True)
            return
        if est.fused_loss and ph.predict_depth and ph.predict_mask and not ph.predict_color:
            # raw head outputs -> fused loss head (csrc/pose_loss.cu): no full-frame intermediates
            logits, latent, _ = ph.decode(self.z_obj, cam, interpret_logits=False, return_latent=self.use_latent)
            terms = ops.pose_loss_terms(logits[:, 0], logits[:, 1], cam.viewport, cam.translation[:, 2],
                                        self.target.depth, self.target.mask, cam.z_span, 0.01, cam.width, cam.height)
            losses = {k: terms[:, i] for i, k in enumerate(self.TERMS[:4])}
            if self.use_latent:
                from .estimation import cosine_distance
                rendered = latent.squeeze(0).flatten(1)
                losses['latent'] = cosine_distance(rendered, self._target_code(cam).flatten(1).expand_as(rendered))
        else:
            z_depth, _, z_mask_logits, z_latent = est._render_observation(self.z_obj, cam)
            code = self._target_code(cam) if self.use_latent else None
            losses = est.loss_func(self.target, z_depth, z_mask_logits, cam, z_pred_latent=z_latent, z_target_latent=code)
        rank = sum(self.weights.get(k, 0.0) * v for k, v in losses.items())
        optim = rank
        if self.sched_w:
            optim = sum((self.sched_w[k] if k in self.sched_w else self.weights.get(k, 0.0)) * v
                        for k, v in losses.items())
        optim.mean().backward()
        with torch.no_grad():
            # snapshot BEFORE the update: the reference ranks the cameras that produced this loss
            self.h_rank.index_copy_(0, self.slot, rank.detach().unsqueeze(0))
            self.h_optim.index_copy_(0, self.slot, optim.detach().unsqueeze(0))
            self.h_terms.index_copy_(0, self.slot, torch.stack([losses[k].detach() for k in self.TERMS]).unsqueeze(0))
            self.h_lq.index_copy_(0, self.slot, self.lq.detach().unsqueeze(0))
            self.h_tr.index_copy_(0, self.slot, self.tr.detach().unsqueeze(0))
            self.slot.add_(1).remainder_(self.chunk)
        self.opt.step(rank.detach())

    def set_schedule_weights(self, step):
        for k, sched in self.est.loss_schedules.items():
            if k in self.sched_w:
                self.sched_w[k].fill_(float(sched.get(step)))

    @staticmethod
    def make_signature(est, z_obj, target_obs, cameras, chunk):
        """Everything the captured graph bakes in: shapes, the identity AND version of every network parameter (the
        graph holds raw pointers into the packed-weight caches, which are keyed by parameter version), the convolution
        precision, the loss weights / schedules and the optimiser hyper-parameters.  A mismatch means re-capture."""
        model = est.model
        nets = (model.photographer,)
        if float(est.loss_weights.get('latent', 0.0)) > 0.0:       # the target code runs the encoder side too
            nets += (model.sculptor, model.fuser)
        params = tuple((id(p), p._version) for net in nets for p in net.parameters())
        return (len(cameras), tuple(z_obj.shape), tuple(target_obs.depth.shape), cameras.width, cameras.height,
                cameras.z_span, params, ops.get_default_precision(), tuple(sorted(est.loss_weights.items())),
                tuple(sorted(est.loss_schedules)), est.learning_rate, est.lr_reduce_patience, est.lr_reduce_threshold,
                est.lr_reduce_factor, bool(est.fused_loss), int(chunk))

    def signature(self):
        return self._signature

    @torch.no_grad()
    def reset(self, z_obj, target_obs, cameras):
        """Re-arm the captured graph for another run with same-shaped inputs (no re-capture)."""
        self.z_obj.copy_(z_obj)
        self.target.color.copy_(target_obs.color); self.target.depth.copy_(target_obs.depth)
        self.target.mask.copy_(target_obs.mask)
        if self.use_latent:
            self._set_target_inputs(target_obs, self.n)
        self.template.intrinsic.copy_(cameras.intrinsic)
        self.lq.copy_(cameras.log_quaternion); self.tr.copy_(cameras.translation); self.vp.copy_(cameras.viewport)
        o = self.opt
        for t in (*o.m, *o.v, o.step_count, o.num_bad):
            t.zero_()
        o.lr.fill_(float(self.est.learning_rate)); o.best.fill_(float('inf'))
        self.slot.zero_()

    def capture(self):
        trace_was = ops.KernelTrace.enabled
        ops.KernelTrace.enabled = False               # no CUDA events between captured nodes
        try:
            self._capture()
        finally:
            ops.KernelTrace.enabled = trace_was

    def _capture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        state = self._save_state()
        with torch.cuda.stream(side):
            for i in range(2):                      # warm-up on a side stream (allocator, packed-weight caches)
                before = ops.KernelTrace.launches
                self._iteration()
                self.launches_per_iteration = ops.KernelTrace.launches - before
        torch.cuda.current_stream().wait_stream(side)
        self._load_state(state)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iteration()
        self._load_state(state)                     # capture does not execute, but keep state pristine

    def _save_state(self):
        o = self.opt
        return [t.detach().clone() for t in (self.lq, self.tr, self.vp, *o.m, *o.v, o.step_count, o.lr, o.best,
                                             o.num_bad, self.slot)]

    def _load_state(self, state):
        o = self.opt
        with torch.no_grad():
            for dst, src in zip((self.lq, self.tr, self.vp, *o.m, *o.v, o.step_count, o.lr, o.best, o.num_bad,
                                 self.slot), state):
                dst.copy_(src)

    def run_chunk(self, first_step, count):
        """Replay `count` (<= chunk) iterations; returns host copies of this chunk's history."""
        for i in range(count):
            if self.est.loss_schedules:
                self.set_schedule_weights(first_step + i)
            self.graph.replay()
        ops.KernelTrace.launches += count * self.launches_per_iteration   # lfb200 kernels inside the replays
        out = {k: getattr(self, 'h_' + k)[:count].cpu() for k in ('rank', 'optim', 'terms', 'lq', 'tr')}
        return out

    def camera_at(self, lq, tr):
        """Full-frame (uncropped) detached camera for one iteration's snapshot, on the host."""
        t = self.template
        return Camera(t.intrinsic.detach().cpu(), None, t.z_span, None, width=t.width, height=t.height,
                      log_quaternion=lq.clone(), translation=tr.clone())
