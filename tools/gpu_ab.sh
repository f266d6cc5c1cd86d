#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 500 python -m pytest tests -m gpu -q --timeout 90 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
echo "== shapes default"; timeout 200 python tools/bench_shapes.py 2>&1 | tee gpurun_out/shapes_default.txt
