import sys, time, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from waveform_b200 import Engine
from oracle.oraclebind import OracleSource
from helpers import synth_pcm, parity_report
print(torch.cuda.get_device_name(0))
cases = [
  ({"fft_size":2048,"window":"hann"},1),
  ({"fft_size":1024,"window":"hann","display_mode":"bars","interp_mode":"catmull_rom"},1),
  ({"fft_size":4096,"window":"blackman_harris","channel_mode":"stereo"},2),
  ({"fft_size":2048,"window":"hamming","slope":0.5,"rolloff_q":1.0,"rolloff_rate":6.0,"fast_peaks":True},2),
  ({"fft_size":8192,"window":"blackman","interp_mode":"lanczos","filter_mode":"gauss","filter_radius":2.5},2),
  ({"fft_size":128,"window":"none","temporal_smoothing":"none"},1),
  ({"fft_size":256},1), ({"fft_size":512},2), ({"fft_size":16384},1), ({"fft_size":32768},1),
]
for s, ch in cases:
    S, T = 5, 12
    eng = Engine(s, channels=ch, max_streams=S)
    N = eng.fft_size; hop = N//2
    x = synth_pcm(S, eng.capture_channels, (T-1)*hop+N, zero_frames=[(1,4,8)], frame_len=N, hop=hop)
    t=time.time(); out = eng.process(x, T, hop, want_points=True); dt=time.time()-t
    refdb = np.zeros_like(out["db"]); refpts = np.zeros_like(out["points"]); refsil=np.zeros_like(out["silent"])
    for i in range(S):
        o = OracleSource(s, channels=ch); r = o.run_stft(x[i], T, hop, want_points=True)
        refdb[i], refpts[i], refsil[i] = r["db"], r["points"], r["silent"]
    rep = parity_report(out["db"], refdb)
    print(s.get("fft_size"), ch, "ok" if rep["ok"] else "FAIL", rep, "pts", np.abs(out["points"]-refpts).max(), "silent", np.array_equal(out["silent"], refsil), f"{dt*1e3:.1f}ms")
    # device path
    xt = torch.from_numpy(x).cuda(); eng2 = Engine(s, channels=ch, max_streams=S)
    o2 = eng2.process(xt, T, hop, want_points=True); torch.cuda.synchronize()
    print("   device==host path:", np.array_equal(o2["db"].cpu().numpy(), out["db"]), "kernel ms", eng2.last_kernel_ms())
