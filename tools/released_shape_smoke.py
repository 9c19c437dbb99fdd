"""Functional check at the RELEASED network widths (reference tools/train/train.sh:37-46: 256^2 input, 16^3 latent,
camera blocks 64/128/256, decoder up to 512 channels): reconstruction from 8 views, render of 4 hypotheses, backward
to the cameras.  Random weights (no checkpoint is available offline).  Prints timings and which conv shapes took the
tcgen05 path.
    python tools/released_shape_smoke.py [--precision 1]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', type=int, default=1)
    ap.add_argument('--views', type=int, default=8)
    ap.add_argument('--hyp', type=int, default=4)
    args = ap.parse_args()
    from tests import parity_helpers as ph
    from latentfusion_b200 import ops
    from latentfusion_b200.recon import models, fusion
    from latentfusion_b200.recon.inference import LatentFusionModel
    from latentfusion_b200.pose import utils as pu
    from latentfusion_b200.utils import parse_block_config as pbc
    dev = torch.device('cuda:0')
    ops.set_default_precision(args.precision)
    torch.manual_seed(0)
    sculptor = models.Sculptor(in_size=256, image_config=pbc("64,D,128,D,196,D,256,D,512,D,512,D,512:512,U,512,U,256"),
                               camera_config=pbc("64,128,256"), object_config=pbc("256,256"), projection_type='factor',
                               input_color=True, input_depth=False, input_mask=True, scale_mode='nearest')
    photographer = models.Photographer(in_size=sculptor.out_size,
                                       image_config=pbc("256,D,512,D,512:512,U,512,U,512,U,256,U,196,U,128,U,64"),
                                       camera_config=pbc("256,256"), object_config=[], projection_type='factor',
                                       predict_depth=True, predict_mask=True, predict_color=False, scale_mode='nearest')
    fuser = fusion.get_fuser('gru', in_channels=256, cube_size=1.0)
    print('sculptor out', sculptor.out_size, sculptor.out_channels, ' photographer out', photographer.out_size)
    S = sculptor.out_size
    cams, dist = ph.synthetic_cameras(args.views, 128, seed=1, perturb=False)       # zoomed to 256^2 crops
    model = LatentFusionModel(sculptor, fuser, photographer, dist, dev)
    for m in (sculptor, fuser, photographer):
        for p in m.parameters():
            p.requires_grad_(False)
    color = torch.rand(1, args.views, 3, 256, 256, device=dev) * 2 - 1
    mask = (torch.rand(1, args.views, 1, 256, 256, device=dev) > 0.3).float()
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.no_grad():
            z_obj, _ = sculptor.encode(fuser, cams.to(dev), color, mask=mask)
        torch.cuda.synchronize()
        print(f'recon call {it}: {1e3 * (time.perf_counter() - t0):.1f} ms  z_obj {tuple(z_obj.shape)}')
    hyp, _ = ph.synthetic_cameras(args.hyp, 128, seed=2)
    for it in range(2):
        cam = pu.parameterize_camera(hyp.to(dev), optimize_viewport=True)
        ops.KernelTrace.reset(it == 1)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y, z = model.render_latent_object(z_obj, cam, return_latent=True, apply_mask=True)
        loss = (y['depth_logits'] ** 2).mean() + (y['mask_logits'] ** 2).mean()
        loss.backward()
        torch.cuda.synchronize()
        print(f'render+backward call {it}: {1e3 * (time.perf_counter() - t0):.1f} ms  depth {tuple(y["depth"].shape)} '
              f'grad finite {bool(torch.isfinite(cam.log_quaternion.grad).all() and torch.isfinite(cam.translation.grad).all())}')
    rows = sorted(ops.KernelTrace.summary().items(), key=lambda kv: -kv[1]['ms_total'])
    for name, d in rows[:12]:
        tfs = d['flops'] / d['ms_total'] / 1e9 if d['flops'] else 0
        print(f"{name:36s} calls {d['calls']:3d} total {d['ms_total']:8.3f} ms  {tfs:7.2f} TF/s")
    ops.KernelTrace.reset(False)


if __name__ == '__main__':
    main()
