"""The oracle against the UNMODIFIED reference compiled here (oracle/_ref/libwaveform_ref.so), run live on fresh
inputs — including the AVX2 path the plugin really executes.  Skipped where the library is absent."""
import numpy as np
import pytest

from helpers import parity_report, synth_pcm
from oracle import refbind
from oracle.oraclebind import OracleSource

pytestmark = [pytest.mark.ref, pytest.mark.skipif(not refbind.available(), reason="oracle/_ref not built")]

CASES = [
    ({"fft_size": 1024, "window": "hann", "display_mode": "bars", "interp_mode": "catmull_rom"}, 1),
    ({"fft_size": 4096, "window": "blackman_harris", "channel_mode": "stereo"}, 2),
    ({"fft_size": 2048, "window": "hann"}, 1),
    ({"fft_size": 2048, "window": "hamming", "slope": 0.5, "rolloff_q": 1.0, "rolloff_rate": 6.0, "fast_peaks": True}, 2),
    ({"fft_size": 8192, "window": "blackman", "interp_mode": "lanczos", "filter_mode": "gauss", "filter_radius": 2.5}, 2),
    ({"fft_size": 800, "window": "power_of_sine", "sine_exponent": 3, "temporal_smoothing": "tv_exp_moving_avg",
      "gravity": 0.4, "log_scale": False, "interp_mode": "point"}, 1),
    ({"fft_size": 2048, "window": "none", "temporal_smoothing": "none", "display_mode": "bars", "interp_mode": "point",
      "normalize_volume": True}, 2),
    ({"fft_size": 1024, "display_mode": "bars", "interp_mode": "lanczos", "bar_width": 4, "bar_gap": 1,
      "filter_mode": "gauss", "mirror_freq_axis": True}, 2),
    ({"fft_size": 16384, "window": "hann"}, 1),
    ({"fft_size": 1024, "channel_mode": "stereo", "gravity": 0.2, "floor": -30}, 2),
]


@pytest.mark.parametrize("settings,channels", CASES)
def test_oracle_matches_compiled_reference(settings, channels):
    _compare(settings, channels)


@pytest.mark.parametrize("seed", range(40))
def test_randomised_settings_oracle_vs_compiled_reference(seed):
    """Differential fuzz of the oracle against the unmodified reference over the settings space the GPU fuzz
    (tests/test_gpu_scale.py::test_randomised_configs_against_oracle) draws from, plus the display options: every table bit for bit,
    every tick's spectrum within the parity criterion, silent flags and points."""
    rng = np.random.default_rng(4000 + seed)
    N = int(rng.choice([128, 512, 1024, 2048, 4096, 8192, 800, 720, 1600, 1920, 1456, 352, 2000, 4160, 1088]))
    mode = str(rng.choice(["mono", "mono", "stereo"]))
    channels = 2 if mode == "stereo" or rng.uniform() < 0.25 else 1
    settings = {"fft_size": N, "channel_mode": mode,
                "window": str(rng.choice(["none", "hann", "hamming", "blackman", "blackman_harris", "power_of_sine"])),
                "temporal_smoothing": str(rng.choice(["none", "exp_moving_avg", "exp_moving_avg", "tv_exp_moving_avg"])),
                "gravity": float(rng.choice([0.2, 0.5, 0.65, 0.9])), "floor": int(rng.choice([-30, -45, -65])),
                "display_mode": str(rng.choice(["curve", "curve", "bars"])),
                "interp_mode": str(rng.choice(["point", "lanczos", "catmull_rom"])),
                "log_scale": bool(rng.uniform() < 0.8)}
    if settings["window"] == "power_of_sine":
        settings["sine_exponent"] = int(rng.choice([1, 2, 3, 4]))
    if rng.uniform() < 0.3:
        settings.update(slope=float(rng.choice([0.25, 1.0])), fast_peaks=bool(rng.uniform() < 0.5))
    if rng.uniform() < 0.3:
        settings.update(rolloff_q=1.0, rolloff_rate=float(rng.choice([3.0, 9.0])))
    if rng.uniform() < 0.3:
        settings.update(filter_mode="gauss", filter_radius=float(rng.choice([1.5, 2.5])))
    if rng.uniform() < 0.2:
        settings["normalize_volume"] = True
    if settings["display_mode"] == "bars" and rng.uniform() < 0.5:
        settings.update(bar_width=int(rng.choice([2, 4, 24])), bar_gap=int(rng.choice([0, 1, 6])))
    hop_div = int(rng.choice([1, 2, 4]))
    _compare(settings, channels, T=int(rng.choice([6, 14, 30])), hop_div=hop_div, seed=seed)


def _compare(settings, channels, T=30, hop_div=2, seed=11):
    ref = refbind.RefSource(settings, impl=refbind.IMPL_GENERIC, channels=channels)
    orc = OracleSource(settings, channels=channels)
    N = ref.fft_size
    hop = N // hop_div
    cc = ref.capture_channels
    x = synth_pcm(1, cc, (T - 1) * hop + N, seed=seed)[0]
    x[:, 6 * hop: 6 * hop + 4 * N] = 0
    if cc == 2:
        x[1, 12 * hop:] = 0  # one channel goes silent: exercises the stale-dB quirk
    rms = (0.05 + 0.3 * np.random.default_rng(1).uniform(size=T)).astype(np.float32) if settings.get("normalize_volume") else None
    a = ref.run_stft(x, T, hop, rms=rms, want_points=True)
    b = orc.run_stft(x, T, hop, rms=rms, want_points=True)
    for name in ("window", "slope", "rolloff", "interp_indices"):
        ra, ob = getattr(ref, name)(), getattr(orc, name)()
        assert (ra is None) == (ob is None), name
        if ra is not None:
            assert np.array_equal(ra, ob), name
    assert ref.window_sum == orc.window_sum
    rk, ok = ref.interp_kernel(), orc.interp_kernel()
    assert rk[0] == ok[0] and (rk[1] is None or np.array_equal(rk[1], ok[1]))
    rep = parity_report(b["db"], a["db"], db_min=ref.db_min)
    assert rep["ok"] and rep["normwise"] < 1e-6, rep
    assert np.array_equal(a["silent"], b["silent"])
    d = np.abs(a["points"].astype(np.float64) - b["points"].astype(np.float64))
    assert d.max() < 5e-3 and np.median(d) < 1e-4


def test_avx2_path_is_within_the_same_tolerance():
    """What the plugin runs on any recent x86 (WAVSourceAVX2, FMA EMA, sqrt(fma)) vs the generic parity target."""
    s = {"fft_size": 2048, "window": "hann"}
    g = refbind.RefSource(s, impl=refbind.IMPL_GENERIC, channels=1)
    a = refbind.RefSource(s, impl=refbind.IMPL_AVX2, channels=1)
    x = synth_pcm(1, 1, 20 * 2048, seed=5)[0]
    rg, ra = g.run_stft(x, 20, 2048), a.run_stft(x, 20, 2048)
    rep = parity_report(ra["db"], rg["db"], db_min=g.db_min)
    assert rep["ok"], rep


def test_av_sync_offset_selects_older_frame():
    """Positive audio/video delay: tick_spectrum takes the OLDEST N of the last delay+N samples
    (src/source_generic.cpp:50-59)."""
    s = {"fft_size": 1024, "window": "hann", "temporal_smoothing": "none", "audio_sync_offset": 0}
    N = 1024
    x = synth_pcm(1, 1, 4 * N, seed=9)[0]
    ref = refbind.RefSource(s, impl=refbind.IMPL_GENERIC, channels=1)
    ref.advance(4 * N / 48000)
    ref.push(x[0])                      # packet ends "now"
    ref.tick()
    latest = ref.decibels(0)
    o = OracleSource({k: v for k, v in s.items() if k != "audio_sync_offset"}, channels=1)
    o.tick([x[0, -N:]])
    assert parity_report(latest, o.decibels(0))["ok"]


PX_CASES = [
    ({"fft_size": 1024, "display_mode": "curve", "interp_mode": "lanczos", "height": 300}, 1),
    ({"fft_size": 2048, "display_mode": "bars", "interp_mode": "catmull_rom", "rounded_caps": True, "min_bar_height": 5,
      "bar_width": 10, "bar_gap": 2}, 1),
    ({"fft_size": 1024, "display_mode": "curve", "channel_mode": "stereo", "channel_spacing": 20, "mirror_freq_axis": True,
      "filter_mode": "gauss", "height": 400}, 2),
    ({"fft_size": 1024, "display_mode": "bars", "channel_mode": "stereo", "channel_spacing": 10, "rounded_caps": True,
      "mirror_freq_axis": True, "interp_mode": "point"}, 2),
]


@pytest.mark.parametrize("settings,channels", PX_CASES)
def test_display_stage_matches_reference_render(settings, channels):
    """dB -> pixel lerp/clamp, mirroring and (miny, minpos): the oracle against WAVSource::render() itself
    (src/source.cpp:1346-1565 run verbatim behind the fake graphics API)."""
    ref = refbind.RefSource(settings, impl=refbind.IMPL_GENERIC, channels=channels)
    orc = OracleSource(settings, channels=channels)
    N, T = ref.fft_size, 6
    x = synth_pcm(1, ref.capture_channels, T * N, seed=4)[0]
    for t in range(T):
        ref.advance(N / 48000)
        ref.push(x[0, t * N:(t + 1) * N], x[1, t * N:(t + 1) * N] if ref.capture_channels > 1 else None)
        ref.tick()
        orc.tick([x[c, t * N:(t + 1) * N] for c in range(ref.capture_channels)])
    ref.render()
    px, miny, minpos = orc.render_pixels()
    for c in range(ref.display_channels):
        assert np.abs(ref.render_buf(c) - px[c]).max() < 5e-4
    # fed with the reference's own m_decibels only the interpolation's summation order differs: on an AVX machine
    # render() takes apply_interp_filter_fma3 (src/source.cpp:1383-1387), the oracle restates the generic template
    px2, _ = orc.pixels_of(np.stack([ref.decibels(c) for c in range(ref.display_channels)]))
    for c in range(ref.display_channels):
        assert np.abs(ref.render_buf(c) - px2[c]).max() < 1e-4
