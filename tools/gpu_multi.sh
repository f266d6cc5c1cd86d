#!/bin/bash
# 2-GPU validation: meter tests (1 GPU), the headline bench and the config-5 exchange under torchrun
set -u
mkdir -p gpurun_out
G=${1:-2}
nvidia-smi -L
echo "== meter tests"; timeout 200 python -m pytest tests/test_meter.py -m gpu -x -q --timeout 90 2>&1 | tail -4
echo "== meter bench"; timeout 100 python tools/bench_meter.py 2>&1 | tee gpurun_out/meter.txt
echo "== bench.py x$G"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 100 --warmup 5 --e2e-steps 3 2>gpurun_out/bench_${G}gpu.err | tail -1 | tee gpurun_out/bench_${G}gpu.json | cut -c1-600
echo "== bench.py x1"
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --e2e-steps 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_1gpu.json | cut -c1-300
echo "== config 5 x$G (NCCL peak all-reduce)"
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 \
    tools/bench_c5.py --streams 128 --frames 64 --steps 10 2>gpurun_out/c5_${G}gpu.err | tail -1 | tee gpurun_out/c5_${G}gpu.json
echo "== config 5 x1"
timeout 200 python tools/bench_c5.py --streams 128 --frames 64 --steps 10 2>/dev/null | tail -1 | tee gpurun_out/c5_1gpu.json
tail -3 gpurun_out/bench_${G}gpu.err gpurun_out/c5_${G}gpu.err
