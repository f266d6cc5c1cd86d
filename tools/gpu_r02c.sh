#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -15 | tee gpurun_out/r02c_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r02c_smoke.txt
timeout 400 python bench.py --steps 100 --warmup 5 > gpurun_out/r02c_bench.json 2>gpurun_out/r02c_bench.err
tail -3 gpurun_out/r02c_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
for l in d['config']['extra']['layouts_same_batch']:
    print(l['streams'],l['frames'],round(l['value']/1e6,1),round(l['frac'],3),round(l['frac_incl_state_bytes'],3),l['kernel'])
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['parity']['ok'])
PY
