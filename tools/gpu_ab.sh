#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 90 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for l in "256 256" "512 128" "1024 64" "4096 16"; do set -- $l
  timeout 120 python bench.py --streams $1 --frames $2 --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print(f\"layout $1 x $2: {d['value']/1e6:8.1f} M spectra/s  frac {r['frac']:.3f}\")"
done | tee gpurun_out/layouts_default_routing.txt
