#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 500 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
echo "== memcheck (R=4 + meter + wave)"
WF_WIDE_R=4 PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py 2>&1 | tail -3 | tee gpurun_out/sanitizer.txt
echo "== shapes default"; timeout 200 python tools/bench_shapes.py 2>&1 | tee gpurun_out/shapes_default.txt
echo "== c5 R=2"; WF_WIDE_R=2 timeout 100 python tools/bench_shapes.py --only=c5 2>&1 | tail -2
echo "== meter + wave"; timeout 200 python tools/bench_meter.py 2>&1 | tee gpurun_out/meter.txt
