#!/bin/bash
# validation + measurements (one gpurun call)
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 420 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -12
echo "== memcheck R=4"
WF_WIDE_R=4 PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py 2>&1 | tail -3
echo "== shapes default"; timeout 200 python tools/bench_shapes.py 2>&1 | tee gpurun_out/shapes_default.txt
echo "== shapes old kernels"; WF_V3=0 timeout 200 python tools/bench_shapes.py --only=c1 --only=generic 2>&1 | tee gpurun_out/shapes_old.txt
echo "== meter"; timeout 200 python tools/bench_meter.py 2>&1 | tee gpurun_out/meter.txt
WF_WIDE_R=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:stft_v3 -s 3 -c 1 -o gpurun_out/prof_v3b_8192 \
    python tools/bench_shapes.py "--only=N=8192 mono" --iters=1 > gpurun_out/ncu_v3b_8192.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stft_v3 -s 3 -c 1 -o gpurun_out/prof_v3b_c1 \
    python tools/bench_shapes.py "--only=c1 N=1024" --iters=1 > gpurun_out/ncu_v3b_c1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
