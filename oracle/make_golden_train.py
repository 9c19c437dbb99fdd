"""TEST INFRASTRUCTURE — golden of ONE generator training iteration of the UNMODIFIED reference modules.

    python oracle/make_golden_train.py           # authoring container only (needs /root/reference)

The body of ReconTrainer.run_iteration (tools/train/train_reconstruct.py:456-534) with the released recipe's switches
(--no-discriminator, hard_smooth_l1 depth, binary_cross_entropy mask, Adam betas (0, 0.99); tools/train/train.sh)
is executed on the reference's own Sculptor / GRUFuser / Photographer and its own loss classes (latentfusion/losses.py,
trainutils.get_recon_criterion / get_optimizer) at a tiny size: B=2 objects, 3 input + 2 output views, S=16, C=8.
(The trainer class itself drags in the dataset / tensorboard stack; its iteration body is restated here statement
for statement.)  Stored: inputs, state_dicts, the three loss values, a spread of weight gradients, and a few
parameters after the optimiser step."""
import json
import math
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
warnings.filterwarnings('ignore')

from oracle import ref_import  # noqa: E402

ref_import.install()

import torch  # noqa: E402

torch.set_num_threads(1)

from latentfusion import consts, three  # noqa: E402
from latentfusion.losses import HardPixelLoss, beta_prior_loss, reduce_loss  # noqa: E402
from latentfusion.modules.geometry import Camera  # noqa: E402
from latentfusion.recon import fusion as ref_fusion  # noqa: E402
from latentfusion.recon import models as ref_models  # noqa: E402
from latentfusion.recon.utils import optimal_camera_dist  # noqa: E402
from latentfusion.utils import parse_block_config as pbc  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'train_s16_c8.npz')
S, C, B, VI, VO = 16, 8, 2, 3, 2
CFG = dict(lr=0.00075, depth_weight=25.0, mask_weight=25.0, beta_weight=1.0, beta_param=0.01, depth_k=300)
PICK = ('sculptor/image_encoder.input_block.conv.module.weight', 'sculptor/camera_blocks.0.conv2.module.weight',
        'sculptor/projection_block.conv.module.weight', 'sculptor/output_block.conv.bias',
        'fuser/gru.update_gate.module.weight', 'fuser/gru.out_gate.bias',
        'photographer/camera_blocks.0.conv1.module.weight', 'photographer/camera_blocks.0.conv2.bias',
        'photographer/projection_block.conv.module.weight', 'photographer/output_blocks.0.conv.module.weight',
        'photographer/output_blocks.1.conv.bias', 'photographer/image_decoder.up_blocks.0.conv1.module.weight')


def npy(t):
    return t.detach().cpu().numpy().copy()      # a copy: optimizer.step() later updates the parameters in place


def cams(n, dist, seed):
    torch.manual_seed(seed)
    K = torch.tensor(consts.INTRINSIC).unsqueeze(0).expand(n, -1, -1).contiguous()
    q = three.orientation.evenly_distributed_quats(n)
    t = torch.tensor([[0.0, 0.0, dist]]).expand(n, -1).contiguous()
    return Camera(K, three.to_extrinsic_matrix(t, q), z_span=0.5, width=640, height=480).zoom(None, 2 * S, dist)


def main():
    torch.manual_seed(0)
    arch_s = dict(in_size=2 * S, image_config=pbc(f"{C},D,{2*C}:{2*C},{2*C}"), camera_config=pbc(f"{C},{C}"),
                  object_config=pbc(f"{C},{C}"), projection_type='factor', input_color=True, input_depth=False,
                  input_mask=True, scale_mode='nearest')
    arch_p = dict(in_size=S, image_config=pbc(f"{C},D,{2*C}:{2*C},U,{2*C},U,{C}"), camera_config=pbc(f"{C},{C}"),
                  object_config=[], projection_type='factor', predict_depth=True, predict_mask=True,
                  predict_color=False, scale_mode='nearest')
    sculptor = ref_models.Sculptor(**arch_s)
    fuser = ref_fusion.get_fuser('gru', in_channels=C, cube_size=1.0)
    photographer = ref_models.Photographer(**arch_p)
    mods = dict(sculptor=sculptor, fuser=fuser, photographer=photographer)
    for m in mods.values():
        for k, p in m.named_parameters():
            if k.endswith('bias'):
                p.data.normal_(0, 0.1)
        m.train()
    dist = optimal_camera_dist(615.4991, 2 * S, 0.5, slack=128 / (2 * S))
    g = {'meta': np.array(json.dumps(dict(S=S, C=C, B=B, VI=VI, VO=VO, camera_dist=dist, arch_sculptor=arch_s,
                                          arch_photographer=arch_p, cfg=CFG, pick=PICK, torch=torch.__version__)))}
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            g[f'{name}/{k}'] = npy(v)
    cam_in, cam_out = cams(B * VI, dist, 31), cams(B * VO, dist, 32)
    for pre, c in (('cam_in', cam_in), ('cam_out', cam_out)):
        for k in ('intrinsic', 'log_quaternion', 'translation', 'viewport'):
            g[f'{pre}.{k}'] = npy(getattr(c, k))
    torch.manual_seed(33)
    P = 2 * S
    image = torch.rand(B, VI, 3, P, P) * 2 - 1
    mask_in = (torch.rand(B, VI, 1, P, P) > 0.4).float()
    gt_depth = torch.rand(B, VO, 1, P, P) * 2 - 1
    gt_mask = (torch.rand(B, VO, 1, P, P) > 0.5).float()
    g['in.image'], g['in.mask'], g['gt.depth'], g['gt.mask'] = npy(image), npy(mask_in), npy(gt_depth), npy(gt_mask)

    # ---- run_iteration body (train_reconstruct.py:456-534), generator only
    params = [p for m in mods.values() for p in m.parameters()]
    # (latentfusion.trainutils itself imports the tensorboard/torchnet stack; these three lines are what its
    #  get_optimizer('adam') :103-105 and get_recon_criterion :114-130 return)
    optimizer = torch.optim.Adam(params, lr=CFG['lr'], betas=(0.0, 0.99))
    depth_criterion = HardPixelLoss(torch.nn.SmoothL1Loss, k=CFG['depth_k'])
    mask_criterion = torch.nn.BCEWithLogitsLoss(reduction='none')
    z_obj, z_extra = sculptor.encode(fuser, camera=cam_in, color=image, depth=None, mask=mask_in)      # :460-465
    y, z, _ = photographer.decode(z_obj, cam_out, interpret_logits=True)                               # :643-646
    g['fwd.z_obj'], g['fwd.depth'], g['fwd.mask_logits'] = npy(z_obj), npy(y['depth']), npy(y['mask_logits'])
    loss_depth = reduce_loss(depth_criterion(y['depth'], gt_depth))                                     # :497-498
    loss_mask = reduce_loss(mask_criterion(y['mask_logits'], gt_mask))                                  # :502-507
    loss_beta = beta_prior_loss(y['mask'], alpha=CFG['beta_param'], beta=CFG['beta_param'])             # :508-510
    loss_g = (CFG['depth_weight'] * loss_depth + CFG['mask_weight'] * loss_mask + CFG['beta_weight'] * loss_beta) / 1
    loss_g.backward()                                                                                   # :527
    g['loss.depth'], g['loss.mask'], g['loss.beta'], g['loss.total'] = (npy(v) for v in (loss_depth, loss_mask, loss_beta, loss_g))
    named = {f'{n}/{k}': p for n, m in mods.items() for k, p in m.named_parameters()}
    for k in PICK:
        g[f'grad/{k}'] = npy(named[k].grad)
    g['gradnorm'] = np.array([float(sum((p.grad.double() ** 2).sum() for p in params).sqrt())])
    optimizer.step()                                                                                    # :534
    for k in PICK:
        g[f'after/{k}'] = npy(named[k])
    np.savez_compressed(OUT, **g)
    print('wrote', OUT, f'{os.path.getsize(OUT) / 1e6:.2f} MB', len(g), 'arrays; losses', float(loss_depth), float(loss_mask), float(loss_beta))


if __name__ == '__main__':
    main()
