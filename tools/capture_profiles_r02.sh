#!/bin/bash
# round-2 ncu captures behind profiles/r02_*: the depth-batched tcgen05 conv (fwd + bwd-data inside one refine iteration),
# the resamplers, the depth collapse, and the launch list of a short bench run.  Run under gpurun, 1 GPU.
set -x
NCU="ncu --set full --clock-control none --import-source on -f"
# one eager refine iteration at precision 1 (tools/profile_iteration.py): conv3d_dz launches 3..6 are one iteration's 2 fwd + 2 bwd
$NCU -k regex:conv3d_dz_kernel -s 4 -c 4 -o gpurun_out/r02_conv3d_dz python tools/profile_iteration.py 1 > /dev/null 2>&1
$NCU -k regex:resample_march -c 1 -o gpurun_out/r02_resample_march python tools/kbench.py --only resample --iters 1 > /dev/null 2>&1
$NCU -k regex:bwd_cam_kernel -c 1 -o gpurun_out/r02_bwd_cam python tools/kbench.py --only resample --iters 1 > /dev/null 2>&1
$NCU -k regex:"collapse_cluster|expand_multi|actnorm_bwd" -s 6 -c 6 -o gpurun_out/r02_misc python tools/profile_iteration.py 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-events --no-strong > gpurun_out/r02_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -8
