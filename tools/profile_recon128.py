"""dev: per-kernel breakdown of the 128^3 reconstruction of bench.py's coarse-search block"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests import parity_helpers as ph
from latentfusion_b200 import ops, dist as lfdist
from latentfusion_b200.recon.inference import LatentFusionModel
dev = torch.device('cuda:0')
ops.set_default_precision(1)
S4, C4, V = 128, 16, 16
sculptor, fuser, photographer, _, _ = ph.random_lfsynth(S4, C4, seed=0, device=dev)
ref_cams, dist_ = ph.synthetic_cameras(V, S4, seed=31, perturb=False)
model = LatentFusionModel(sculptor, fuser, photographer, dist_, dev)
g = torch.Generator().manual_seed(33)
color = torch.rand(1, V, 3, 2 * S4, 2 * S4, generator=g) * 2 - 1
mask = (torch.rand(1, V, 1, 2 * S4, 2 * S4, generator=g) > 0.3).float()
for it in range(4):
    ops.KernelTrace.reset(it == 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        lfdist.build_latent_object_sharded(model, ref_cams, color, mask, 0, 1)
    e1.record(); torch.cuda.synchronize()
    print(f'recon {it}: {e0.elapsed_time(e1):.1f} ms')
rows = sorted(ops.KernelTrace.summary().items(), key=lambda kv: -kv[1]['ms_total'])
print('kernel total', sum(d['ms_total'] for _, d in rows))
for k, d in rows[:10]:
    print(f"  {k:30s} calls {d['calls']:4d} total {d['ms_total']:8.2f} avg {d['ms_avg']:.3f}")
