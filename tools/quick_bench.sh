#!/bin/bash
run() { timeout 90 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 "$@" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print(f\"  value {d['value']/1e6:8.1f} M/s  kernel {r['kernel_ms']*1e3:7.1f} us  frac {r['frac']:.3f}  clocks {d['clocks']['sm_mhz']}\")
except Exception as ex:
    print('  FAILED/timeout', ex)"; }
echo "S=2048xT=32"; run
echo "S=4096xT=16"; run --streams 4096 --frames 16
echo "S=4096xT=16 wpc=14"; WF_FAST_WPC=14 run --streams 4096 --frames 16
echo "S=8192xT=8"; run --streams 8192 --frames 8
echo "S=65536xT=1"; run --streams 65536 --frames 1
