#!/bin/bash
# A/B evidence run (one gpurun call): correctness first, then shapes under each variant.
set -u
mkdir -p gpurun_out
echo "== pytest gpu (default build)"
timeout 420 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -6
echo "== memcheck, v3 kernel forced R=4"
WF_WIDE_R=4 PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py 2>&1 | tail -4
for v in default r1 r2 r8 old; do
  echo "== shapes: $v"
  case $v in
    default) timeout 200 python tools/bench_shapes.py ;;
    r1) WF_WIDE_R=1 timeout 200 python tools/bench_shapes.py ;;
    r2) WF_WIDE_R=2 timeout 200 python tools/bench_shapes.py --only=c --only=N=4096 --only=N=8192;;
    r8) WF_WIDE_R=8 timeout 200 python tools/bench_shapes.py --only=c --only=N=4096 --only=N=8192;;
    old) WF_V3=0 timeout 200 python tools/bench_shapes.py --only=c --only=N=4096 --only=N=8192;;
  esac 2>&1 | tee gpurun_out/shapes_$v.txt
done
echo "== meter bench"
timeout 200 python tools/bench_meter.py 2>&1 | tee gpurun_out/meter.txt
