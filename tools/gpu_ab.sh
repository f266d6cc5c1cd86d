#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 500 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -12
echo "== memcheck R=4 (single-barrier rounds)"
WF_WIDE_R=4 PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py 2>&1 | tail -3
echo "== shapes default"; timeout 200 python tools/bench_shapes.py 2>&1 | tee gpurun_out/shapes_default.txt
echo "== meter + wave"; timeout 200 python tools/bench_meter.py 2>&1 | tee gpurun_out/meter.txt
