#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 240 2>&1 | tail -8 | tee gpurun_out/r02l_pytest.txt
timeout 400 python tools/bench_shapes.py 2>&1 | tee gpurun_out/r02l_shapes.txt
