"""Per-kernel CUDA-event breakdown of one reconstruction (build_latent_object, config B: 16 views, GRU fuser).
    python tools/profile_recon.py [--precision 1]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', type=int, default=1)
    args = ap.parse_args()
    import bench
    import parity_helpers as ph
    from latentfusion_b200 import ops, dist as lfdist
    from latentfusion_b200.recon.inference import LatentFusionModel
    dev = torch.device('cuda:0')
    ops.set_default_precision(args.precision)
    inp = bench.synthetic_inputs()
    sculptor, fuser, photographer, arch, sds = ph.random_lfsynth(bench.S, bench.C, seed=0, device=dev)
    model = LatentFusionModel(sculptor, fuser, photographer, inp['dist'], dev)
    with torch.no_grad():
        for it in range(3):
            ops.KernelTrace.reset(it == 2)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            lfdist.build_latent_object_sharded(model, inp['ref_cams'], inp['color'], inp['mask'], 0, 1)
            t1.record()
            torch.cuda.synchronize()
            print(f"recon call {it}: {t0.elapsed_time(t1):.2f} ms")
    rows = sorted(ops.KernelTrace.summary().items(), key=lambda kv: -kv[1]['ms_total'])
    total = sum(d['ms_total'] for _, d in rows)
    print(f"kernel time through the C ABI: {total:.2f} ms")
    for name, d in rows:
        gbs = d['bytes'] / d['ms_total'] / 1e6 if d['bytes'] else 0
        tfs = d['flops'] / d['ms_total'] / 1e9 if d['flops'] else 0
        print(f"{name:36s} calls {d['calls']:4d}  total {d['ms_total']:8.3f} ms  avg {d['ms_avg']:7.4f} ms  {gbs:7.1f} GB/s  {tfs:6.2f} TF/s")
    ops.KernelTrace.reset(False)


if __name__ == '__main__':
    main()
