#!/usr/bin/env python
"""bench.py — spectra/s of the fused window+R2C-FFT+mag+EMA+dBFS path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA engine (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N ...          # the reference's own CPU path (oracle/_ref), all host cores

One "step" = one pass of the hot path over one batch of synthetic 48 kHz PCM already resident in HBM:
65 536 mono frames of N=2048 per GPU, laid out as 4096 independent streams x 16 consecutive frames (hop = N)
so that the EMA recurrence state stays on-chip between frames.  Inputs (512 MiB)
and outputs (256 MiB) are each larger than the 126 MB L2, so every step streams from/to HBM.

The JSON line printed by rank 0 follows the driver's contract; see DESIGN.md §Measurement for the fields.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "spectra/sec fused window+R2C-FFT+mag+smooth N=2048 batch=65536; %HBM roofline"
UNIT = "spectra/s"
SETTINGS = {"fft_size": 2048, "window": "hann", "temporal_smoothing": "exp_moving_avg", "gravity": 0.65,
            "channel_mode": "mono"}
N_FFT = 2048
BINS = N_FFT // 2
BYTES_PER_FRAME = N_FFT * 4 + BINS * 4  # PCM in + bins out


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this same command
    (profiles/traffic.json, written by tools/ncu_summary.py --traffic); None if there is no capture."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        d = json.loads(p.read_text())
        return float(d["dram_bytes_per_launch"]), d.get("source")
    except Exception:
        return None, None


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref), or the oracle port if _ref is absent
# ----------------------------------------------------------------------------------------------------
def cpu_reference_throughput(frames_per_thread: int, threads: int, chunk_frames: int = 512):
    """Spectra/s of the reference's CPU path on `threads` host threads, N=2048 Hann EMA dBFS.

    Each thread owns one WAVSourceAVX2 (+ its FFTW plan) and walks `frames_per_thread` consecutive frames
    (hop = N) of its own channel through capture_audio()/tick(); the PCM chunk is reused so it stays in cache,
    which is how the plugin sees audio too.  ctypes releases the GIL, so the threads run in parallel.
    """
    import numpy as np

    from oracle import refbind

    kind = "reference"
    use_ref = refbind.available()
    if use_ref:
        mk = lambda: refbind.RefSource(dict(SETTINGS), impl=refbind.IMPL_AVX2, channels=1)
        desc = "oracle/_ref: WAVSourceAVX2::tick_spectrum + vendored FFTW 3.3.11 (AVX2 codelets)"
    else:
        from oracle import oraclebind
        kind = "port"
        mk = lambda: oraclebind.OracleSource(dict(SETTINGS), channels=1)
        desc = "oracle/liboracle.so: scalar C restatement of source_generic.cpp"
    rng = np.random.default_rng(0xB200)
    srcs = [mk() for _ in range(threads)]
    pcms = [(0.25 * rng.uniform(-1, 1, (1, (chunk_frames + 1) * N_FFT))).astype(np.float32) for _ in range(threads)]
    reps = max(1, frames_per_thread // chunk_frames)

    def work(i):
        for _ in range(reps):
            srcs[i].run_stft(pcms[i], chunk_frames, N_FFT, want_db=False)

    # warm-up (page in, FFTW plan, caches)
    for i in range(threads):
        srcs[i].run_stft(pcms[i], 8, N_FFT, want_db=False)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    frames = threads * reps * chunk_frames
    return frames / dt, {"kind": kind, "cores": threads, "desc": desc, "frames": frames, "seconds": dt}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_thread = 8192
    vals = []
    for _ in range(args.warmup):
        cpu_reference_throughput(1024, cores)
    info = None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, info = cpu_reference_throughput(per_thread, cores)
        vals.append(v)
    dt = time.perf_counter() - t0
    value = statistics.median(vals)
    frames_per_step = info["frames"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "fft_size": N_FFT, "frames_per_step": frames_per_step,
                   "note": "CPU arm: bounded sample of the same workload per step, all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                         "sample": f"{frames_per_step} frames/step ({info['desc']}), hop=N, data cache-resident"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id: str):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", gpu_id], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=3.0):
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin, t_end):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        # nvidia-smi stamps a sample up to one period after the state it shows; take the samples around the timed region
        rows = [r for (t, r) in self.rows if t_begin <= t <= t_end + 0.05] or \
               [r for (t, r) in self.rows if t_begin - 0.1 <= t <= t_end + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_name():
    return ("c3: 65536 mono frames/GPU, N=2048 Hann, EMA g=0.65, dBFS; 4096 streams x 16 consecutive frames, hop=N "
            "(BASELINE.json configs[2])")


def make_pcm_on_device(torch, S, T, device, seed):
    """SURVEY.md §8(d) synthetic signal, generated on the device: noise + two sines per stream, 1% all-zero frames."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ns = T * N_FFT
    n = torch.arange(ns, device=device, dtype=torch.float32)
    c = torch.arange(S, device=device, dtype=torch.float32)
    fc = 110.0 * torch.pow(2.0, torch.remainder(c, 60.0) / 12.0)
    pcm = torch.empty((S, 1, ns), device=device, dtype=torch.float32)
    chunk = 512
    for s0 in range(0, S, chunk):
        s1 = min(S, s0 + chunk)
        ph = 2.0 * torch.pi * fc[s0:s1, None] * n[None, :] / 48000.0
        x = 0.25 * (2.0 * torch.rand((s1 - s0, ns), device=device, generator=g) - 1.0)
        x += 0.5 * torch.sin(ph) + 0.1 * torch.sin(3.01 * ph)
        pcm[s0:s1, 0] = x
    zero = torch.rand((S, T), device=device, generator=g) < 0.01
    pcm.view(S, T, N_FFT)[zero] = 0.0
    return pcm


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from waveform_b200 import Engine

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    S, T = args.streams, args.frames
    frames_per_gpu = S * T
    eng = Engine(dict(SETTINGS), channels=1, max_streams=S, device=local_rank)
    pcm = make_pcm_on_device(torch, S, T, device, seed=0xB200 + rank)
    out = torch.empty((S, T, 1, BINS), device=device, dtype=torch.float32)
    stream = torch.cuda.Stream(device=device)  # a real (non-default) stream: timed events and launches share it
    torch.cuda.synchronize()

    def step():
        eng.process_raw(pcm.data_ptr(), S, T, N_FFT, T * N_FFT, T * N_FFT, out_db=out.data_ptr(),
                        stream=stream.cuda_stream, sync=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    props = torch.cuda.get_device_properties(device)
    gpu_id = "GPU-" + str(props.uuid) if hasattr(props, "uuid") else str(local_rank)
    sampler = ClockSampler(gpu_id) if rank == 0 else None
    if sampler:
        sampler.wait_first()
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = eng.launch_count
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_begin = time.perf_counter()
    evs[0].record(stream)
    for k in range(args.steps):
        step()
        evs[k + 1].record(stream)
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    launches = eng.launch_count - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    tmax = torch.tensor([total_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    ms_per_step = total_ms_max / args.steps
    value = frames_per_gpu * world / (ms_per_step * 1e-3)

    # ---- end-to-end: host (pinned) buffers through the C-ABI, H2D + kernel + D2H inside the timed region ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_pcm = torch.empty((S, 1, T * N_FFT), dtype=torch.float32, pin_memory=True)
    h_pcm.copy_(pcm)
    h_out = torch.empty((S, T, 1, BINS), dtype=torch.float32, pin_memory=True)

    def step_e2e():
        eng.process_raw(h_pcm.data_ptr(), S, T, N_FFT, T * N_FFT, T * N_FFT, out_db=h_out.data_ptr(),
                        stream=stream.cuda_stream, sync=False)

    step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        step_e2e()
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = frames_per_gpu * world / (float(e2e_ms.item()) * 1e-3 / e2e_steps)
    checksum = float(h_out[:: max(1, S // 64)].double().sum())  # host read of the step's result

    if rank == 0:
        peak_gbs, peak_src = load_peaks()
        traffic, traffic_src = load_traffic() if (S, T) == (4096, 16) else (None, None)
        kernel_ms = statistics.mean(per_launch)
        achieved = frames_per_gpu * BYTES_PER_FRAME / (kernel_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, info = cpu_reference_throughput(args.cpu_frames_per_thread, cores)
            cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                   "sample": f"{info['frames']} frames in {info['seconds']:.2f}s wall ({info['desc']}), hop=N, "
                             "same N=2048 Hann EMA dBFS settings, data cache-resident"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "fft_size": N_FFT, "streams_per_gpu": S, "frames_per_stream": T,
                       "hop": N_FFT, "frames_per_step_per_gpu": frames_per_gpu,
                       "l2_policy": "inputs 512 MiB + outputs 256 MiB per step exceed the 126 MB L2 (no flush needed)",
                       "parallelism": f"streams sharded over {world} GPU(s), no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_pcm.numel() * 4),
                    "d2h_bytes_per_step": int(h_out.numel() * 4), "steps": e2e_steps, "checksum": checksum},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "kernel": "stft2048_fast_kernel<16,true,true,false> (csrc/wf_fast2048.cuh)", "kernel_ms": kernel_ms,
                         "bytes_per_launch": frames_per_gpu * BYTES_PER_FRAME},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="independent streams per GPU")
    ap.add_argument("--frames", type=int, default=16, help="consecutive frames per stream")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-frames-per-thread", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
