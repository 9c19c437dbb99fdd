#!/usr/bin/env python
"""dev: one reconstruction training iteration (latentfusion_b200.train.ReconTrainStep) at BASELINE configs[3] extents
per GPU — B objects x V_in input views, V_out reconstruction views, LF-synth(64, 32) — with the per-entry-point
CUDA-event breakdown.   python tools/bench_train.py [-B 2] [--vin 16] [--vout 8] [--precision 1] [--steps 3]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentfusion_b200 import ops
from latentfusion_b200.modules.geometry import Camera
from latentfusion_b200.train import ReconTrainStep
from tests import parity_helpers as ph


def synthetic_batch(B, vin, vout, S, dev, seed=0):
    P = 2 * S
    cin, dist = ph.synthetic_cameras(B * vin, S, seed=seed + 1, perturb=False)
    cout, _ = ph.synthetic_cameras(B * vout, S, seed=seed + 2, perturb=False)
    torch.manual_seed(seed + 3)
    return {'in': {'camera': cin.to(dev), 'image': torch.rand(B, vin, 3, P, P, device=dev) * 2 - 1,
                   'mask': (torch.rand(B, vin, 1, P, P, device=dev) > 0.4).float()},
            'out_gt': {'camera': cout.to(dev), 'depth': torch.rand(B, vout, 1, P, P, device=dev) * 2 - 1,
                       'mask': (torch.rand(B, vout, 1, P, P, device=dev) > 0.5).float()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('-B', type=int, default=2)
    ap.add_argument('--vin', type=int, default=16)
    ap.add_argument('--vout', type=int, default=8)
    ap.add_argument('-S', type=int, default=64)
    ap.add_argument('-C', type=int, default=32)
    ap.add_argument('--precision', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    ops.set_default_precision(a.precision)
    sculptor, fuser, photographer, arch, sds = ph.random_lfsynth(a.S, a.C, seed=0, device=dev)
    step = ReconTrainStep(sculptor, fuser, photographer, depth_k=4096)
    batch = synthetic_batch(a.B, a.vin, a.vout, a.S, dev)
    for _ in range(2):
        step.run_iteration(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = step.run_iteration(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    ops.KernelTrace.reset(True)
    step.run_iteration(batch)
    torch.cuda.synchronize()
    summ = ops.KernelTrace.summary()
    ops.KernelTrace.reset(False)
    tot = sum(d['ms_total'] for d in summ.values())
    print(json.dumps(dict(ms_per_step=ms, views_per_s=a.B * (a.vin + a.vout) / ms * 1e3, loss=float(out['total']),
                          traced_ms=tot, mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)))
    for k, d in sorted(summ.items(), key=lambda kv: -kv[1]['ms_total'])[:18]:
        print(f"  {k:34s} calls {d['calls']:4d}  total {d['ms_total']:8.3f} ms  avg {d['ms_avg']:.3f}")


if __name__ == '__main__':
    main()
