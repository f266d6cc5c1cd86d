#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q --timeout 240 -k "team or fast2048 or golden or stall or skip" 2>&1 | tail -5 | tee gpurun_out/r02r_pytest.txt
lay() { timeout 120 python bench.py --streams $1 --frames $2 --steps 40 --warmup 5 --no-cpu-baseline --no-parity --no-layouts --no-c5 --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print(f\"  {d['value']/1e6:8.1f} M spectra/s  kernel {r['kernel_ms']*1e3:7.1f} us  frac {r['frac']:.3f}  {r['kernel']}\")"; }
for l in "256 256" "512 128" "1024 64" "148 443" "296 221"; do set -- $l
  echo "layout $1 x $2 (auto)"; lay $1 $2
done | tee gpurun_out/r02r_layouts.txt
