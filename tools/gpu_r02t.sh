#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q --timeout 240 -k "not warp2" 2>&1 | tail -5 | tee gpurun_out/r02t_pytest.txt
echo "== TMA variant, 768 threads/SM (80 regs)"; timeout 300 python tools/bench_shapes.py --only=N=4096 --only=N=8192 "--only=c4'" 2>&1 | grep -v "^env" | tee gpurun_out/r02t_shapes.txt
echo "== TMA variant, 640 threads/SM (102 regs)"; WF_LIB_PATH=$PWD/waveform_b200/lib_b/libwfstft.so timeout 300 python tools/bench_shapes.py --only=N=4096 --only=N=8192 2>&1 | grep -v "^env" | tee -a gpurun_out/r02t_shapes.txt
echo "== register prefetch (WF_V3_TMA=0)"; WF_V3_TMA=0 timeout 300 python tools/bench_shapes.py --only=N=4096 --only=N=8192 2>&1 | grep -v "^env" | tee -a gpurun_out/r02t_shapes.txt
