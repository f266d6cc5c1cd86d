#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/debug_v3.py 2>&1 | tail -12 | tee gpurun_out/r02n_debug.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -12 | tee gpurun_out/r02m_pytest.txt
timeout 200 python tools/bench_meter.py 2>&1 | tee gpurun_out/r02m_meter.txt
