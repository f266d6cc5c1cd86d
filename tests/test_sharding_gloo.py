"""N>1 host logic on CPU: world_size-2 gloo processes shard the streams, exchange the per-frame peak with a MAX
all-reduce and apply the gain — and must agree with the single-process result.  The spectra themselves come from
the oracle here (there is no GPU in this container); the GPU version of the same flow is in test_gpu_parity.py /
bench.py --peak-normalize."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SETTINGS = {"fft_size": 1024, "window": "hann"}
S, T, N = 7, 5, 1024


def _spectra(pcm):
    from oracle.oraclebind import OracleSource
    return np.stack([OracleSource(SETTINGS, channels=1).run_stft(pcm[s], T, N)["db"] for s in range(pcm.shape[0])])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import synth_pcm
    from waveform_b200.shard import allreduce_peak, peak_gain, shard_streams

    first, count = shard_streams(S, rank, world)
    pcm = synth_pcm(S, 1, T * N)[first:first + count]
    db = _spectra(pcm)                                       # [count, T, 1, B]
    peak = torch.from_numpy(db[..., 1:].max(axis=(0, 2, 3)).copy())
    allreduce_peak(peak)
    gain = peak_gain(peak, -3.0, 20.0).numpy()
    db[..., 1:] += gain[None, :, None, None]
    q.put((rank, first, count, db, peak.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_streams_partition():
    from waveform_b200.shard import shard_streams
    for n in (1, 7, 8, 4096, 65537):
        for w in (1, 2, 4, 8):
            blocks = [shard_streams(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_two_rank_peak_normalise_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from helpers import synth_pcm
    from waveform_b200.shard import peak_gain
    whole = _spectra(synth_pcm(S, 1, T * N))
    peak = whole[..., 1:].max(axis=(0, 2, 3))
    exp = whole.copy()
    exp[..., 1:] += peak_gain(peak, -3.0, 20.0)[None, :, None, None]
    got = np.concatenate([r[3] for r in res], axis=0)
    assert np.array_equal(res[0][4], peak) and np.array_equal(res[1][4], peak)
    assert np.allclose(got, exp, atol=1e-6)
