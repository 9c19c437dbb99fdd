#!/usr/bin/env python
"""Kernel-level breakdown of one refine iteration (eager) with torch.profiler: which launches are lfb200
kernels and which are torch glue (camera algebra, Adam, bookkeeping)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from tests import parity_helpers as ph
from latentfusion_b200 import ops
from latentfusion_b200.observation import Observation
from latentfusion_b200.pose import estimation
from latentfusion_b200.recon.inference import LatentFusionModel

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device('cuda:0')
ops.set_default_precision(prec)
inp = bench.synthetic_inputs()
sculptor, fuser, photographer, arch, sds = ph.random_lfsynth(bench.S, bench.C, seed=0, device=dev)
model = LatentFusionModel(sculptor, fuser, photographer, inp['dist'], dev)
z_obj = torch.randn(1, 1, bench.C, bench.S, bench.S, bench.S, device=dev) * 0.5
gt_full = inp['gt'].uncrop()
target = Observation(torch.zeros(1, 3, 480, 640), inp['tdepth'], inp['tmask'], gt_full).to(dev)
cfg = {'type': 'gradient', 'args': dict(bench.EST_ARGS, num_iters=3), 'loss_weights': bench.LOSS_WEIGHTS}
est = estimation.load_from_config(cfg, model)
est.estimate(z_obj, target, camera=bench.hypothesis_cameras(gt_full, bench.N_HYP, 7).to(dev))
r = est._refiner
for _ in range(2):
    r._iteration()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    r._iteration()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:60]][0] += 1
        agg[e.name[:60]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if 'lf::' in k or 'tc::' in k)
n_ours = sum(v[0] for k, v in agg.items() if 'lf::' in k or 'tc::' in k)
n_all = sum(v[0] for v in agg.values())
print(f'total kernel time {tot/1e3:.3f} ms over {n_all} launches; lfb200 {ours/1e3:.3f} ms over {n_ours}; torch glue {(tot-ours)/1e3:.3f} ms over {n_all-n_ours}')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f'{v[1]/1e3:8.3f} ms  n={v[0]:4d}  {k}')
