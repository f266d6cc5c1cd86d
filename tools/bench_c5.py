#!/usr/bin/env python
"""BASELINE.json configs[4]: N=16384, streams sharded over the GPUs, cross-channel peak normalisation through one
NCCL MAX all-reduce of n_frames floats per step (waveform_b200/shard.py), then the local wf_peak_normalize pass.

  torchrun --nproc-per-node G --master-addr 127.0.0.1 tools/bench_c5.py [--streams 128 --frames 64 --steps 10]

Checks on every rank that the all-reduced peak is the maximum of the gathered local peaks and that the normalised output is
the local output plus the gain; prints whole-job spectra/s (max over ranks of the device time) from rank 0."""
import argparse, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
from waveform_b200 import Engine
from waveform_b200.shard import ShardedEngine, shard_streams

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=128, help="streams per GPU")
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
N, S, T = 16384, args.streams, args.frames
first, count = shard_streams(S * world, rank, world)
assert count == S
eng = Engine({"fft_size": N, "window": "hann"}, channels=1, max_streams=S, device=local)
sh = ShardedEngine(eng)
g = torch.Generator(device=dev); g.manual_seed(0xB200 + rank)
pcm = (torch.rand((S, 1, T * N), device=dev, generator=g) - 0.5) * (0.1 + 0.2 * rank)

# ---- correctness of the exchange ----
plain = Engine({"fft_size": N, "window": "hann"}, channels=1, max_streams=S, device=local).process(pcm, T, N, want_peak=True)
out = sh.process_normalized(pcm, T, N, target_db=-3.0, max_gain=30.0)
torch.cuda.synchronize()
local_peak = plain["peak"].clone()
if world > 1:
    allp = [torch.empty_like(local_peak) for _ in range(world)]
    dist.all_gather(allp, local_peak)
    want = torch.stack(allp).max(dim=0).values
else:
    want = local_peak
assert torch.equal(out["peak"], want), "all-reduced peak != max of local peaks"
gain = torch.clamp(-3.0 - want, max=30.0)
exp = plain["db"].clone(); exp[..., 1:] += gain[None, :, None, None]
assert torch.equal(out["db"], exp), "normalised output != local output + gain"

# ---- throughput: kernel + all-reduce + normalise per step ----
def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
for _ in range(3):
    sh.process_normalized(pcm, T, N)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    sh.process_normalized(pcm, T, N)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    frames = S * T * world
    print(json.dumps({"config": "c5 N=16384 peak-normalised", "n_gpus": world, "streams_per_gpu": S, "frames": T,
                      "ms_per_step": float(ms), "spectra_per_s": frames / (float(ms) * 1e-3),
                      "exchange": "NCCL all_reduce(MAX) of n_frames floats" if world > 1 else "none (1 GPU)",
                      "checks": "peak == max(local peaks); out == local + gain (bit-exact)"}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
