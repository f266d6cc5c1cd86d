#!/bin/bash
# round-2 first GPU call: new scale tests + bench (with parity subset, layouts, c5) + smoke + shapes baseline
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -25 | tee gpurun_out/r02a_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r02a_smoke.txt
timeout 400 python bench.py --steps 100 --warmup 5 > gpurun_out/r02a_bench.json 2>gpurun_out/r02a_bench.err
tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
WF_LAZY_HOLD=0 timeout 200 python bench.py --streams 65536 --frames 1 --steps 50 --no-cpu-baseline --no-parity --no-layouts --no-c5 --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('65536x1 WF_LAZY_HOLD=0:', d['value']/1e6, r['frac'])" | tee gpurun_out/r02a_lazy_ab.txt
timeout 300 python tools/bench_shapes.py 2>&1 | tee gpurun_out/r02a_shapes.txt
timeout 200 python tools/bench_next_rows.py 2>&1 | tail -30 | tee gpurun_out/r02a_next_rows.txt
nvidia-smi topo -m > gpurun_out/r02a_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r02a_topo.txt
