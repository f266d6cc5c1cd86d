#!/bin/bash
# quick A/B of kernel variants on the GPU box: prints spectra/s + roofline fraction per variant (each run time-boxed)
run() { timeout 90 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 1 "$@" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print(f\"  value {d['value']/1e6:8.1f} M/s  kernel {r['kernel_ms']*1e3:7.1f} us  frac {r['frac']:.3f}  clocks {d['clocks']['sm_mhz']}\")
except Exception as ex:
    print('  FAILED/timeout', ex)"; }
echo "MAXW=16 S=4096xT=16"; WF_FAST_MAXW=16 run
echo "MAXW=12 S=4096xT=16"; WF_FAST_MAXW=12 run
echo "MAXW=16 wpc=16 S=4096xT=16"; WF_FAST_MAXW=16 WF_FAST_WPC=16 run
echo "MAXW=16 S=8192xT=8"; WF_FAST_MAXW=16 run --streams 8192 --frames 8
echo "MAXW=16 S=2048xT=32"; WF_FAST_MAXW=16 run --streams 2048 --frames 32
echo "MAXW=16 S=65536xT=1"; WF_FAST_MAXW=16 run --streams 65536 --frames 1
echo "MAXW=16 S=256xT=256"; WF_FAST_MAXW=16 run --streams 256 --frames 256
