#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q --timeout 240 -k "par16384 or warp2 or other_kernels or 16384 or wide" 2>&1 | tail -8 | tee gpurun_out/r02p_pytest.txt
timeout 300 python tools/bench_shapes.py --only=c5 --only=N=800 --only=N=1920 --only=N=1600 2>&1 | tee gpurun_out/r02p_shapes.txt
WF_PAR16384=0 timeout 300 python tools/bench_shapes.py --only=c5 2>&1 | tee -a gpurun_out/r02p_shapes.txt
