#!/bin/bash
# A/B evidence run for the generic + wide kernels (one gpurun call): correctness first, then shapes under each variant.
set -u
mkdir -p gpurun_out
echo "== pytest gpu (default build, automatic cluster size)"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== pytest gpu parity with the wide kernel disabled (one-group kernel + prefetch)"
WF_WIDE_R=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spectrum_parity or golden" 2>&1 | tail -3
echo "== memcheck, wide kernel forced R=4"
WF_WIDE_R=4 PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py 2>&1 | tail -6
for v in default wide1 t384 t256; do
  echo "== shapes: $v"
  case $v in
    default) timeout 200 python tools/bench_shapes.py ;;
    wide1) WF_WIDE_R=1 timeout 200 python tools/bench_shapes.py ;;
    t384) WF_LIB_PATH=$PWD/waveform_b200/lib/variants/t384/libwfstft.so timeout 200 python tools/bench_shapes.py ;;
    t256) WF_LIB_PATH=$PWD/waveform_b200/lib/variants/t256/libwfstft.so WF_WIDE_R=1 timeout 200 python tools/bench_shapes.py ;;
  esac 2>&1 | tee gpurun_out/shapes_$v.txt
done
