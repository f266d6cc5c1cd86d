#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 240 2>&1 | tail -15 | tee gpurun_out/r02g_pytest.txt
timeout 300 python tools/bench_next_rows.py 2>&1 | tail -22 | tee gpurun_out/r02g_next_rows.txt
