#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q --timeout 240 -k "par16384 or other_kernels" 2>&1 | tail -5 | tee gpurun_out/r02q_pytest.txt
timeout 300 python tools/bench_shapes.py --only=c5 2>&1 | tee gpurun_out/r02q_shapes.txt
