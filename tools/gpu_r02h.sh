#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 240 2>&1 | tail -12 | tee gpurun_out/r02h_pytest.txt
timeout 200 python tools/bench_meter.py 2>&1 | tee gpurun_out/r02h_meter.txt
WF_METER_FUSED=0 timeout 200 python tools/bench_meter.py 2>&1 | head -3 | tee gpurun_out/r02h_meter_unfused.txt
