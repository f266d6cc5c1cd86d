#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -x -q --timeout 240 -k "pair4096 or other_kernels" 2>&1 | tail -4 | tee gpurun_out/r02k_pytest.txt
timeout 300 python tools/bench_shapes.py "--only=N=4096 mono" 2>&1 | tee gpurun_out/r02k_shapes.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair -s 3 -c 1 -o gpurun_out/r02k_pair4096 \
    python tools/bench_shapes.py "--only=N=4096 mono" --iters=1 > gpurun_out/r02k_ncu.log 2>&1
