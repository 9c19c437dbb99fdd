"""The generator half of one reconstruction training iteration on the lfb200 kernels.

Mirror of ``ReconTrainer.run_iteration`` (reference tools/train/train_reconstruct.py:421-534) for the released recipe
(tools/train/train.sh: ``--no-discriminator``, hard smooth-L1 depth + BCE mask losses, Adam betas (0, 0.99)):

    z_obj = Sculptor.encode(fuser, in cameras, in colour / mask)                 :460-465
    depth, mask, mask_logits = Photographer.decode(z_obj, recon cameras)         :466-467, :643-655
    loss_g = w_d * HardPixelLoss(SmoothL1, k)(depth, gt) + w_m * BCE(mask_logits, gt) [+ w_b * beta prior]   :491-521
    loss_g.backward(); optimizer.step()                                          :523-534

Batches arrive already processed (zoomed / normalised tensors + zoomed cameras — the dataset side, recon/utils.py:107-128,
is outside the hot path).  Multi-GPU: the views of every object are sharded over the ranks — each rank
encodes its input views, the per-view cubes are all-gathered with a differentiable collective, the fuser runs
replicated, each rank decodes its share of the reconstruction cameras, takes its share of the loss, and the weight
gradients are summed over the ranks in one flat all-reduce before the (replicated) optimiser step."""
import math

import torch
import torch.nn.functional as F

from . import dist as lfdist
from .recon.models import gan_normalize
from .three.batchview import b2bv, bv2b


def hard_pixel_loss(x, y, k, base='smooth_l1'):
    """reference losses.py:33-57: per-image mean over channels, top-k hardest pixels, mean (summed form returned too)."""
    x = x.reshape(-1, *x.shape[-3:])
    y = y.reshape(-1, *y.shape[-3:])
    loss = (F.smooth_l1_loss if base == 'smooth_l1' else F.l1_loss)(x, y, reduction='none')
    loss = loss.mean(dim=1).reshape(x.shape[0], -1)
    k = min(int(k), loss.shape[1])
    top, _ = torch.topk(loss, k=k, dim=1, largest=True)
    return top.mean(dim=1)                      # [images]; the reference then takes the mean over images


def beta_prior_loss(t, alpha, beta, eps=1e-4):
    """reference losses.py:88-99 (per element)."""
    log_beta = math.lgamma(alpha) + math.lgamma(beta) - math.lgamma(alpha + beta)
    loss = (alpha - 1.0) * torch.log(t.clamp(min=eps)) + (beta - 1.0) * torch.log((1.0 - t).clamp(min=eps)) - log_beta
    return (-loss).clamp(min=0)


class ReconTrainStep:
    def __init__(self, sculptor, fuser, photographer, lr=0.00075, depth_weight=25.0, mask_weight=25.0, beta_weight=0.0,
                 beta_param=0.01, depth_loss='hard_smooth_l1', depth_k=16384, batch_groups=1, optimizer='adam',
                 group=None):
        self.sculptor, self.fuser, self.photographer = sculptor, fuser, photographer
        self.depth_weight, self.mask_weight, self.beta_weight, self.beta_param = depth_weight, mask_weight, beta_weight, beta_param
        self.depth_loss, self.depth_k, self.batch_groups, self.group = depth_loss, depth_k, batch_groups, group
        self.parameters = [p for m in (sculptor, fuser, photographer) for p in m.parameters()]
        for m in (sculptor, fuser, photographer):
            m.train(True)
            m.requires_grad_(True)
        opt = {'adam': torch.optim.Adam, 'adamw': torch.optim.AdamW}[optimizer]
        self.optimizer = opt(self.parameters, lr=lr, betas=(0.0, 0.99))       # trainutils.py:103-107

    # ------------------------------------------------------------------------------------------------
    def forward(self, batch):
        """-> (depth, mask, mask_logits) for this rank's reconstruction cameras, [B, v_out_local, 1, P, P] each."""
        sc = self.sculptor
        cam_in, cam_out = batch['in']['camera'], batch['out_gt']['camera']
        color, mask = batch['in']['image'], batch['in']['mask']
        views = color.shape[1]
        planes = []
        if sc.input_color:
            planes.append(bv2b(color))
        if sc.input_depth:
            planes.append(bv2b(batch['in']['depth']))
        if sc.input_mask:
            planes.append(gan_normalize(bv2b(mask)))
        z_views, _, _ = sc(torch.cat(planes, dim=1), cam_in)                       # [B*v_local, C, S, S, S]
        z_views = lfdist.all_gather_views(b2bv(z_views, views), dim=1, group=self.group)
        z_obj, _ = self.fuser(z_views, [], [], None)                               # replicated on every rank
        y, _, _ = self.photographer.decode(z_obj, cam_out, interpret_logits=True)
        return y['depth'], y['mask'], y['mask_logits']

    def losses(self, depth, mask, mask_logits, gt_depth, gt_mask, total_images):
        """this rank's SHARE of the reference's losses (means over all B*V_out images of the global batch)"""
        n_local = depth.shape[0] * depth.shape[1]
        if self.depth_loss.startswith('hard_'):
            ld = hard_pixel_loss(depth, gt_depth, self.depth_k, self.depth_loss[5:]).sum() / total_images
        else:
            fn = F.smooth_l1_loss if self.depth_loss == 'smooth_l1' else F.l1_loss
            ld = fn(depth, gt_depth, reduction='sum') / (total_images * depth[0, 0].numel())
        lm = F.binary_cross_entropy_with_logits(mask_logits, gt_mask, reduction='sum') / (total_images * mask_logits[0, 0].numel())
        lb = beta_prior_loss(mask, self.beta_param, self.beta_param).sum() / (total_images * mask[0, 0].numel())
        total = (self.depth_weight * ld + self.mask_weight * lm + self.beta_weight * lb) / self.batch_groups
        return dict(depth=ld, mask=lm, beta=lb, total=total, images=n_local)

    def run_iteration(self, batch, train=True, is_step=True):
        world = torch.distributed.get_world_size(self.group) if torch.distributed.is_initialized() else 1
        gt = batch['out_gt']
        total_images = gt['depth'].shape[0] * gt['depth'].shape[1] * world
        with torch.set_grad_enabled(train):
            depth, mask, mask_logits = self.forward(batch)
            out = self.losses(depth, mask, mask_logits, gt['depth'], gt['mask'], total_images)
        if train:
            out['total'].backward()
            if is_step:
                lfdist.allreduce_gradients(self.parameters, self.group)
                self.optimizer.step()
                self.optimizer.zero_grad(set_to_none=True)
        report = {k: v.detach() for k, v in out.items() if torch.is_tensor(v)}
        if world > 1:
            for k in report:                     # the global figures are the sums of the ranks' shares
                torch.distributed.all_reduce(report[k], group=self.group)
        return report
