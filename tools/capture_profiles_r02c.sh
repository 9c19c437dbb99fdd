#!/bin/bash
# round-2 (late) captures behind profiles/r02c_*: the depth-marching camera-gradient kernel, the one-pass 2-D tcgen05
# convolution, the loss head, and the launch list of a short bench run.  Run under gpurun, 1 GPU.
set -x
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:bwd_cam_march -c 1 -o gpurun_out/r02c_bwd_cam_march python tools/kbench.py --only bwdcam --iters 1 > /dev/null 2>&1
$NCU -k regex:"pose_loss_bwd|pose_loss_sums|conv_tc_kernel" -s 20 -c 14 -o gpurun_out/r02c_iter_small python tools/profile_iteration.py 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-events --no-strong --no-train --no-search > gpurun_out/r02c_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -6
