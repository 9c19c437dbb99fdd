"""TEST INFRASTRUCTURE — golden vectors of the IBR colour branch from the UNMODIFIED reference.

    python oracle/make_golden_ibr.py          # authoring container only (needs /root/reference)

Calls ``latentfusion.ibr`` directly on small synthetic inputs (no networks involved) and writes
``tests/golden/ibr_p24.npz``: cameras, inputs, ``reproject_views`` / ``render_ibr`` (all four weight types) /
``blend_logits`` / ``warp_blend_logits`` outputs.
"""
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

from latentfusion import ibr as ref_ibr  # noqa: E402
from latentfusion.modules.geometry import Camera  # noqa: E402
from latentfusion.pose import utils as ref_pu  # noqa: E402
from latentfusion.recon.utils import optimal_camera_dist  # noqa: E402

from oracle.make_golden import OUT, cam_dict, npy, reference_cameras  # noqa: E402


def smooth_depth(n, p, seed):
    """normalised depth maps in (-1, 1): a smooth bump plus a little noise."""
    torch.manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, p), torch.linspace(-1, 1, p), indexing='ij')
    base = 0.6 - 0.9 * (xx ** 2 + yy ** 2)
    return (base[None, None] + 0.05 * torch.randn(n, 1, p, p)).clamp(-0.95, 0.95)


def main():
    P, VI, VO, C = 24, 3, 2, 3
    dist = optimal_camera_dist(615.4991, P, 0.5, slack=128 / P)
    cam_in_full = reference_cameras(VI, P, dist, seed=20)
    cam_in = cam_in_full.zoom(None, P, dist)
    torch.manual_seed(21)
    cam_out = Camera.cat([ref_pu.perturb_camera(cam_in_full[k:k + 1], 0.02, 25.0 / 180.0 * math.pi) for k in range(VO)])
    cam_out = cam_out.zoom(None, P, dist)
    torch.manual_seed(22)
    image_in = torch.rand(VI, C, P, P) * 2 - 1
    depth_in = smooth_depth(VI, P, 23)
    depth_out = smooth_depth(VO, P, 24)

    g = {'meta': np.array(json.dumps(dict(P=P, VI=VI, VO=VO, C=C, camera_dist=dist, torch=torch.__version__)))}
    g.update(cam_dict('cam_in', cam_in))
    g.update(cam_dict('cam_out', cam_out))
    g['image_in'], g['depth_in'], g['depth_out'] = npy(image_in), npy(depth_in), npy(depth_out)
    with torch.no_grad():
        g['warp_field'] = npy(ref_ibr.depth_to_warp_field(cam_in, cam_out, depth_out))
        img_r, dep_r = ref_ibr.reproject_views(image_in, depth_in, depth_out, cam_in, cam_out)
        g['image_reproj'], g['depth_reproj'] = npy(img_r), npy(dep_r)
        for wt in ('cam_dist', 'cam_angle', 'cam_hybrid', 'depth'):
            fake, reproj = ref_ibr.render_ibr(cam_in, cam_out, image_in[None], depth_in[None], depth_out[None],
                                              p=0.5, weight_type=wt, eps=1e-2)
            g[f'render_ibr.{wt}'] = npy(fake)
        torch.manual_seed(25)
        logits = torch.randn(VO, 3 * VI, P, P)
        out, wts = ref_ibr.blend_logits(logits[:, :VI], img_r)
        g['logits'], g['blend.image'], g['blend.weights'] = npy(logits), npy(out), npy(wts)
        out, wts, dx, dy = ref_ibr.warp_blend_logits(logits, img_r, 5)
        g['warp_blend.image'], g['warp_blend.weights'] = npy(out), npy(wts)
        g['warp_blend.dx'], g['warp_blend.dy'] = npy(dx), npy(dy)
    path = os.path.join(OUT, 'ibr_p24.npz')
    np.savez_compressed(path, **g)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
