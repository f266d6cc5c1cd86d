#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 420 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -8
echo "== shapes default"; timeout 200 python tools/bench_shapes.py 2>&1 | tee gpurun_out/shapes_default.txt
echo "== c5 with R=2 / R=8"
WF_WIDE_R=2 timeout 100 python tools/bench_shapes.py --only=c5 2>&1 | tail -2
WF_WIDE_R=8 timeout 100 python tools/bench_shapes.py --only=c5 2>&1 | tail -2
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stft_v3 -s 3 -c 1 -o gpurun_out/prof_v3c_c5 \
    python tools/bench_shapes.py "--only=c5 N=16384" --iters=1 > gpurun_out/ncu_v3c_c5.log 2>&1
ls -la gpurun_out/prof_v3c_c5.ncu-rep
