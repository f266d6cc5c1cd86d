#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 240 2>&1 | tail -12 | tee gpurun_out/r02i_pytest.txt
timeout 200 python tools/bench_meter.py 2>&1 | head -3 | tee gpurun_out/r02i_meter.txt
timeout 300 python tools/bench_shapes.py --only=N=4096 2>&1 | tee gpurun_out/r02i_shapes.txt
WF_PAIR_MIN_STREAMS=1000000000 timeout 300 python tools/bench_shapes.py --only=N=4096 2>&1 | tee -a gpurun_out/r02i_shapes.txt
