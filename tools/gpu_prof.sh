#!/bin/bash
# correctness of the touched paths, then ncu captures of the v3 kernel and the meter launch list
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_meter.py tests/test_gpu_parity.py -m gpu -x -q --timeout 90 -k "meter or wide or rms_feed or parity_vs_oracle" 2>&1 | tail -15
WF_WIDE_R=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:stft_v3 -s 3 -c 1 -o gpurun_out/prof_v3_8192 \
    python tools/bench_shapes.py "--only=N=8192 mono" --iters=1 > gpurun_out/ncu_v3_8192.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stft_v3 -s 3 -c 1 -o gpurun_out/prof_v3_c5 \
    python tools/bench_shapes.py "--only=c5 N=16384" --iters=1 > gpurun_out/ncu_v3_c5.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stft_v3 -s 3 -c 1 -o gpurun_out/prof_v3_c2 \
    python tools/bench_shapes.py "--only=c2 stereo" --iters=1 > gpurun_out/ncu_v3_c2.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:meter -c 40 --csv --log-file gpurun_out/meter_launches.csv \
    python tools/bench_meter.py > gpurun_out/ncu_meter.log 2>&1
python tools/bench_shapes.py --only=N=4096 --only=N=8192 --only=c 2>&1 | tee gpurun_out/shapes_new.txt
ls -la gpurun_out/*.ncu-rep
