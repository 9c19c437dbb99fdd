"""TEST INFRASTRUCTURE — golden vectors for the fusers that lfsynth_s16_c8.npz does not cover (LSTM, concat),
from the UNMODIFIED reference (recon/fusion.py:87-92, :204-246; modules/lstm.py:41-56).

    python oracle/make_golden_fusers.py       # authoring container only (needs /root/reference)
"""
import json
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

from latentfusion.recon import fusion as ref_fusion  # noqa: E402
from latentfusion.recon.utils import optimal_camera_dist  # noqa: E402
from latentfusion.utils import parse_block_config as pbc  # noqa: E402

from oracle.make_golden import OUT, cam_dict, npy, reference_cameras, sd_dict  # noqa: E402


def main():
    C, S, V = 8, 10, 4
    torch.manual_seed(30)
    z_obj = torch.randn(1, V, C, S, S, S)
    lstm = ref_fusion.get_fuser('lstm', in_channels=C, cube_size=1.0)
    for k, p in lstm.named_parameters():
        if k.endswith('bias'):
            p.data.normal_(0, 0.1)
    g = {'meta': np.array(json.dumps(dict(C=C, S=S, V=V, torch=torch.__version__))), 'z_obj': npy(z_obj)}
    g.update(sd_dict('lstm', lstm))
    with torch.no_grad():
        g['fused.lstm'] = npy(lstm(z_obj, None, None, None)[0])
        g['fused.concat'] = npy(ref_fusion.get_fuser('concat', in_channels=C, cube_size=1.0)(z_obj, None, None, None)[0])
    # gradient of the LSTM fusion w.r.t. the per-view cubes (training path)
    zt = z_obj.clone().requires_grad_(True)
    torch.manual_seed(31)
    w = torch.randn(1, 1, C, S, S, S)
    (lstm(zt, None, None, None)[0] * w).sum().backward()
    g['lstm.w'], g['lstm.grad_z'] = npy(w), npy(zt.grad)
    # blend fuser (fusion.py:95-149): UNet3d on [z_cam, depth coordinate] -> camera->object resample -> softmax over views
    S2 = 8
    dist = optimal_camera_dist(615.4991, 2 * S2, 0.5, slack=128 / (2 * S2))
    cam = reference_cameras(V, 2 * S2, dist, seed=32).zoom(None, 2 * S2, dist)
    torch.manual_seed(33)
    blend_cfg = "8,D,8:8,U,8"
    blend = ref_fusion.get_fuser('blend', in_channels=C, cube_size=1.0, block_config=pbc(blend_cfg))
    for k, p in blend.named_parameters():
        if k.endswith('bias'):
            p.data.normal_(0, 0.1)
    z_cam = torch.randn(1, V, C, S2, S2, S2)
    z_obj2 = torch.randn(1, V, C, S2, S2, S2)
    g.update(sd_dict('blend', blend))
    g.update(cam_dict('blend.cam', cam))
    g['blend.cfg'] = np.array(blend_cfg)
    g['blend.z_cam'], g['blend.z_obj'] = npy(z_cam), npy(z_obj2)
    with torch.no_grad():
        fused, extra = blend(z_obj2, [z_cam], None, cam)
    g['fused.blend'], g['blend.weights'] = npy(fused), npy(extra['blend_weights'])
    path = os.path.join(OUT, 'fusers_c8_s10.npz')
    np.savez_compressed(path, **g)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
