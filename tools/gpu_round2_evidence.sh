#!/bin/bash
# Round-2 evidence run on one B200: tests, smoke, both bench arms, ncu launch list + full captures, shapes, next rows, meter, sanitizer.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_smoke.txt
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_reference.json 2>gpurun_out/r02_bench_reference.err
timeout 400 python bench.py --steps 200 --warmup 5 > gpurun_out/r02_bench_ours.json 2>gpurun_out/r02_bench_ours.err
tail -c 600 gpurun_out/r02_bench_ours.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stft|fill_kernel|peak_normalize|materialize" -c 60 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-layouts --no-c5 --e2e-steps 1 > gpurun_out/r02_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft2048_fast -s 4 -c 1 -o gpurun_out/r02_fast2048 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-layouts --no-c5 --e2e-steps 1 > gpurun_out/r02_ncu_fast.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft2048_team -s 3 -c 1 -o gpurun_out/r02_team2048 \
    python bench.py --streams 256 --frames 256 --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-layouts --no-c5 --e2e-steps 1 > gpurun_out/r02_ncu_team.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_warp2 -s 3 -c 1 -o gpurun_out/r02_warp2_800 \
    python tools/bench_shapes.py "--only=N=800 (auto" --iters=1 > gpurun_out/r02_ncu_warp2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_warp2 -s 3 -c 1 -o gpurun_out/r02_warp2_display_800 \
    python tools/bench_shapes.py "--only=disp N=800" --iters=1 > gpurun_out/r02_ncu_warp2d.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:wave_chunk -s 3 -c 1 -o gpurun_out/r02_wave_chunk \
    python tools/bench_meter.py > gpurun_out/r02_ncu_wave.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:meter_fused -s 3 -c 1 -o gpurun_out/r02_meter_fused \
    python tools/bench_meter.py > gpurun_out/r02_ncu_meter.log 2>&1
timeout 400 python tools/bench_shapes.py 2>&1 | tee gpurun_out/r02_shapes.txt
timeout 300 python tools/bench_next_rows.py 2>&1 | tee gpurun_out/r02_next_rows.txt
timeout 200 python tools/bench_meter.py 2>&1 | tee gpurun_out/r02_meter.txt
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 400 compute-sanitizer --tool memcheck python tools/sanitize.py 2>&1 | tail -6 | tee gpurun_out/r02_sanitizer.txt
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize.py 2>&1 | tail -3 | tee -a gpurun_out/r02_sanitizer.txt
ls -la gpurun_out/*.ncu-rep
