#!/bin/bash
# 8-GPU box: the driver's scaling sequence (N = 1, 2, 4, 8) + the NUMA-binding A/B of the end-to-end path at N = 8
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo8.txt 2>&1
run() { n=$1; shift; if [ "$n" = 1 ]; then timeout 400 python bench.py --gpus 1 --steps 100 --warmup 5 "$@"; else
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 100 --warmup 5 "$@"; fi; }
run 8 > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err
run 8 --no-numa-bind --no-c5 --no-parity > gpurun_out/r02_scale_n8_nonuma.json 2> gpurun_out/r02_scale_n8_nonuma.err
run 1 --no-layouts --no-cpu-baseline > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
run 2 > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err
run 4 > gpurun_out/r02_scale_n4.json 2> gpurun_out/r02_scale_n4.err
python - <<'PY'
import json
for n in ("1","2","4","8","8_nonuma"):
    try:
        d=json.loads(open(f"gpurun_out/r02_scale_n{n}.json").read().strip().splitlines()[-1])
        c5=(d["config"]["extra"].get("c5") or {})
        st=(d["config"]["extra"].get("strong_scaling") or {})
        print(n, "value %.1f M"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], "e2e %.2f M"%(d["e2e"]["value"]/1e6),
              "h2d", d["e2e"]["per_gpu_h2d_gbs_in_e2e"], "alone", d["e2e"]["per_gpu_h2d_gbs_copy_alone"],
              "c5 %.1f M exact=%s"%(c5.get("value",0)/1e6, c5.get("exchange_exact")), "strong %.1f M"%(st.get("value",0)/1e6))
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r02_scale_n8.err
