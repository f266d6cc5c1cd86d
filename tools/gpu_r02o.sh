#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_seam.py tests/test_meter.py -m gpu -q --timeout 240 2>&1 | grep -v "^  \|^$" | tail -120 > gpurun_out/r02o_pytest.txt
tail -5 gpurun_out/r02o_pytest.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 240 2>&1 | tail -4
