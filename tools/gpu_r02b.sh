#!/bin/bash
# team kernel: tests + layouts
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q --timeout 180 2>&1 | tail -25 | tee gpurun_out/r02b_pytest.txt
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -m gpu -x -q --timeout 180 2>&1 | tail -5 | tee -a gpurun_out/r02b_pytest.txt
lay() { timeout 120 python bench.py --streams $1 --frames $2 --steps 40 --warmup 5 --no-cpu-baseline --no-parity --no-layouts --no-c5 --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print(f\"  {d['value']/1e6:8.1f} M spectra/s  kernel {r['kernel_ms']*1e3:7.1f} us  frac {r['frac']:.3f}  {r['kernel']}\")"; }
for l in "256 256" "512 128" "1024 64" "2048 32" "148 443" "296 221"; do set -- $l
  echo "layout $1 x $2 (auto)"; lay $1 $2
done | tee gpurun_out/r02b_layouts.txt
for w in 16 8 4; do echo "256x256 WF_TEAM_W=$w"; WF_TEAM_W=$w lay 256 256; done | tee -a gpurun_out/r02b_layouts.txt
for w in 8 4 2; do echo "512x128 WF_TEAM_W=$w"; WF_TEAM_W=$w lay 512 128; done | tee -a gpurun_out/r02b_layouts.txt
for w in 4 2 1; do echo "1024x64 WF_TEAM_W=$w"; WF_TEAM_W=$w lay 1024 64; done | tee -a gpurun_out/r02b_layouts.txt
for w in 2 1; do echo "2048x32 WF_TEAM_W=$w"; WF_TEAM_W=$w lay 2048 32; done | tee -a gpurun_out/r02b_layouts.txt
