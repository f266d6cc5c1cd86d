#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_seam.py -m gpu -x -q --timeout 180 2>&1 | tail -15 | tee gpurun_out/r02f_pytest.txt
timeout 300 python tools/bench_next_rows.py 2>&1 | tail -12 | tee gpurun_out/r02f_next_rows.txt
