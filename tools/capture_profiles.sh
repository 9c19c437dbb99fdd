#!/bin/bash
# ncu captures behind profiles/: one launch of each hot kernel (--set full) + the launch list of a short bench run.
set -x
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:resample_march -c 1 -o gpurun_out/r01b_resample_march python tools/kbench.py --only resample --iters 1 > /dev/null 2>&1
$NCU -k regex:bwd_cam_kernel -c 1 -o gpurun_out/r01b_bwd_cam python tools/kbench.py --only resample --iters 1 > /dev/null 2>&1
$NCU -k regex:"conv_tc_kernel|collapse_cluster" -c 3 -o gpurun_out/r01b_conv python tools/kbench.py --only conv --precision 1 --iters 1 > /dev/null 2>&1
$NCU -k regex:ibr_reproject -c 1 -o gpurun_out/r01b_ibr python tools/kbench.py --only ibr --iters 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-events > gpurun_out/r01b_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -8
