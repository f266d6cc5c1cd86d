#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -q --timeout 240 -k "zero_copy" 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_v3 -s 3 -c 1 -o gpurun_out/r02_v3_8192 \
    python tools/bench_shapes.py "--only=N=8192 mono" --iters=1 > gpurun_out/r02_ncu_v3_8192.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_v3 -s 3 -c 1 -o gpurun_out/r02_v3_4096 \
    python tools/bench_shapes.py "--only=N=4096 mono" --iters=1 > gpurun_out/r02_ncu_v3_4096.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:parity -s 2 -c 1 -o gpurun_out/r02_par16384 \
    python tools/bench_shapes.py "--only=c5 full" --iters=1 > gpurun_out/r02_ncu_par16384.log 2>&1
ls -la gpurun_out/r02_v3_*.ncu-rep gpurun_out/r02_par16384.ncu-rep
