#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair -s 3 -c 1 -o gpurun_out/r02j_pair4096 \
    python tools/bench_shapes.py "--only=N=4096 mono" --iters=1 > gpurun_out/r02j_ncu.log 2>&1
tail -3 gpurun_out/r02j_ncu.log
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q --timeout 240 -k "pair4096 or per_tick or warp2" 2>&1 | tail -5 | tee gpurun_out/r02j_pytest.txt
