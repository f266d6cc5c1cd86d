#!/bin/bash
run() { timeout 90 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 "$@" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print(f\"  value {d['value']/1e6:8.1f} M/s  kernel {r['kernel_ms']*1e3:7.1f} us  frac {r['frac']:.3f}  clocks {d['clocks']['sm_mhz']}\")
except Exception as ex:
    print('  FAILED/timeout', ex)"; }
echo "A S=4096xT=16"; WF_FAST_KERNEL=a run
echo "B S=4096xT=16"; WF_FAST_KERNEL=b run
echo "B S=4096xT=16 groups=12"; WF_FAST_KERNEL=b WF_FAST_WPC=12 run
echo "B S=2048xT=32"; WF_FAST_KERNEL=b run --streams 2048 --frames 32
echo "B S=8192xT=8"; WF_FAST_KERNEL=b run --streams 8192 --frames 8
echo "B S=65536xT=1"; WF_FAST_KERNEL=b run --streams 65536 --frames 1
