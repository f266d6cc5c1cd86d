#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 420 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -8
echo "== shapes default"; timeout 200 python tools/bench_shapes.py 2>&1 | tee gpurun_out/shapes_default.txt
echo "== shapes t768 (80 registers, 3 CTAs of 256 threads per SM)"
WF_LIB_PATH=$PWD/waveform_b200/lib/variants/t768/libwfstft.so timeout 200 python tools/bench_shapes.py 2>&1 | tee gpurun_out/shapes_t768.txt
echo "== meter"; timeout 200 python tools/bench_meter.py 2>&1 | tee gpurun_out/meter.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:meter -c 24 --csv --log-file gpurun_out/meter_launches.csv \
    python tools/bench_meter.py > gpurun_out/ncu_meter.log 2>&1
