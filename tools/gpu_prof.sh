#!/bin/bash
# ncu captures of the generic one-group kernel and of the wide (cluster) kernel on their benchmark shapes.
set -u
mkdir -p gpurun_out
WF_WIDE_R=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_fused -s 3 -c 1 -o gpurun_out/prof_gen8192 \
    python tools/bench_shapes.py "--only=N=8192 mono" --iters=1 > gpurun_out/ncu_gen8192.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_wide -s 3 -c 1 -o gpurun_out/prof_wide_c5 \
    python tools/bench_shapes.py "--only=c5 N=16384" --iters=1 > gpurun_out/ncu_wide_c5.log 2>&1
tail -3 gpurun_out/ncu_gen8192.log gpurun_out/ncu_wide_c5.log
ls -la gpurun_out/*.ncu-rep
