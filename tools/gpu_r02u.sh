#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -q --timeout 240 -k "par16384 or other_kernels" 2>&1 | tail -3 | tee gpurun_out/r02u_pytest.txt
timeout 300 python tools/bench_shapes.py --only=c5 2>&1 | grep -v "^env" | tee gpurun_out/r02u_shapes.txt
echo "== live tick, N=2048 on the warp-per-stream kernel (default) vs the CTA-per-tick kernel (WF_FORCE_GENERIC=1)"
timeout 200 python tools/bench_next_rows.py 2>&1 | grep "wf_process alone" | tee gpurun_out/r02u_live.txt
WF_FORCE_GENERIC=1 timeout 200 python tools/bench_next_rows.py 2>&1 | grep "wf_process alone" | tee -a gpurun_out/r02u_live.txt
