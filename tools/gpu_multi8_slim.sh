#!/bin/bash
# 8-GPU box, one pass at HEAD: the driver's N=8 line (weak scaling, e2e, strong scaling, config 5 with the NCCL MAX all-reduce) and the
# N=1 headline on the same box for the efficiency ratio
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 100 --warmup 5 \
    > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err
timeout 200 python bench.py --gpus 1 --steps 100 --warmup 5 --no-layouts --no-cpu-baseline --no-c5 > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
python - <<'PY'
import json
for n in ("1","8"):
    try:
        d=json.loads(open(f"gpurun_out/r02_scale_n{n}.json").read().strip().splitlines()[-1])
        c5=(d["config"]["extra"].get("c5") or {})
        st=(d["config"]["extra"].get("strong_scaling") or {})
        print(n, "value %.1f M"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], "e2e %.2f M"%(d["e2e"]["value"]/1e6),
              "c5 %.1f M exact=%s"%(c5.get("value",0)/1e6, c5.get("exchange_exact")), "strong %.1f M"%(st.get("value",0)/1e6), "parity", (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r02_scale_n8.err
