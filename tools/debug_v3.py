import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from helpers import synth_pcm
from waveform_b200 import Engine
settings = {"fft_size": 4096, "window": "blackman_harris", "channel_mode": "stereo"}
S, T, hop_div, R = 3, 13, 4, 2
os.environ["WF_V3"] = "1"; os.environ["WF_WIDE_R"] = "1"
e1 = Engine(settings, channels=2, max_streams=S)
os.environ["WF_WIDE_R"] = str(R)
e2 = Engine(settings, channels=2, max_streams=S)
N = 4096; hop = N // hop_div
pcm = synth_pcm(S, 2, (T - 1) * hop + N, zero_frames=[(1, 2, 6)], frame_len=N, hop=hop)
x = torch.from_numpy(pcm).cuda()
a = e1.process(x, T, hop, want_points=True, want_peak=True); n1 = e1.last_kernel_name()
b = e2.process(x, T, hop, want_points=True, want_peak=True); n2 = e2.last_kernel_name()
torch.cuda.synchronize()
print(n1, "|", n2)
for key in ("db", "points", "silent", "peak"):
    A, B = a[key].cpu().numpy(), b[key].cpu().numpy()
    if not np.array_equal(A, B):
        d = np.argwhere(A != B)
        print(key, "differs at", len(d), "places; first", d[:5].tolist(), "values", [ (float(A[tuple(i)]), float(B[tuple(i)])) for i in d[:5]])
        if key == "db":
            print(" by stream", np.unique(d[:,0], return_counts=True), "by tick", np.unique(d[:,1], return_counts=True), "by ch", np.unique(d[:,2], return_counts=True))
    else:
        print(key, "equal")
