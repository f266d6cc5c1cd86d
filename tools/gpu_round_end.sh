#!/bin/bash
# One-shot evidence run on a B200 box: tests, smoke, bench (both arms), ncu launch list + full capture of the headline
# kernel, layout sweep, the other BASELINE shapes, level meter.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err
timeout 300 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_ours.json 2>gpurun_out/bench_ours.err
tail -c 1500 gpurun_out/bench_ours.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stft|fill_kernel|peak_normalize" -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft2048 -s 4 -c 1 -o gpurun_out/prof_final \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
lay() { timeout 120 python bench.py --streams $1 --frames $2 --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print(f\"  {d['value']/1e6:8.1f} M spectra/s  kernel {r['kernel_ms']*1e3:7.1f} us  frac_of_measured_hbm {r['frac']:.3f}\")"; }
for l in "4096 16" "2048 32" "8192 8" "16384 4" "65536 1" "1024 64" "512 128" "256 256"; do set -- $l
  echo "layout streams=$1 frames=$2 (warp-per-stream kernel)"; lay $1 $2
done | tee gpurun_out/layouts.txt
for l in "1024 64" "512 128" "256 256"; do set -- $l
  echo "layout streams=$1 frames=$2 (cluster kernel, WF_FAST_MIN_STREAMS=1000000)"; WF_FAST_MIN_STREAMS=1000000 lay $1 $2
done | tee -a gpurun_out/layouts.txt
timeout 200 python tools/bench_shapes.py 2>&1 | tee gpurun_out/shapes.txt
WF_WIDE_R=4 PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py 2>&1 | tail -4 | tee gpurun_out/sanitizer.txt
timeout 100 python tools/bench_meter.py 2>&1 | tee gpurun_out/meter.txt
