"""ctypes binding for oracle/liboracle.so — the plain-C restatement of the reference's spectrum path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this; the product path (waveform_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "liboracle.so"

WINDOWS = {"none": 0, "hann": 1, "hamming": 2, "blackman": 3, "blackman_harris": 4, "power_of_sine": 5}
INTERPS = {"point": 0, "lanczos": 1, "catmull_rom": 2}
FILTERS = {"none": 0, "gauss": 1}
TSMOOTH = {"none": 0, "exp_moving_avg": 1, "tv_exp_moving_avg": 2}
DISPLAYS = {"curve": 0, "bars": 1, "stepped_bars": 1}


class WfoConfig(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_uint32), ("capture_channels", C.c_int32), ("fft_size", C.c_int32),
        ("window", C.c_int32), ("sine_exponent", C.c_int32), ("tsmoothing", C.c_int32),
        ("gravity", C.c_float), ("fast_peaks", C.c_int32), ("slope", C.c_float),
        ("rolloff_q", C.c_float), ("rolloff_rate", C.c_float),
        ("cutoff_low", C.c_int32), ("cutoff_high", C.c_int32),
        ("floor_db", C.c_int32), ("ceiling_db", C.c_int32), ("stereo", C.c_int32),
        ("normalize_volume", C.c_int32), ("volume_target", C.c_float), ("max_gain", C.c_float),
        ("silence_gate", C.c_int32), ("display_mode", C.c_int32),
        ("width", C.c_int32), ("bar_width", C.c_int32), ("bar_gap", C.c_int32),
        ("log_scale", C.c_int32), ("mirror_freq_axis", C.c_int32),
        ("interp_mode", C.c_int32), ("filter_mode", C.c_int32), ("filter_radius", C.c_float),
        ("height", C.c_int32), ("channel_spacing", C.c_int32), ("rounded_caps", C.c_int32), ("min_bar_height", C.c_int32),
    ]


class WfoWaveConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("capture_channels", C.c_int32), ("stereo", C.c_int32), ("width", C.c_int32),
                ("meter_ms", C.c_int32), ("normalize_volume", C.c_int32), ("volume_target", C.c_float),
                ("max_gain", C.c_float)]


class WfoMeterConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("capture_channels", C.c_int32), ("meter_ms", C.c_int32),
                ("rms_mode", C.c_int32), ("tsmoothing", C.c_int32), ("gravity", C.c_float),
                ("fast_peaks", C.c_int32), ("floor_db", C.c_int32)]


_lib = None


def build():
    subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        build()
    L = C.CDLL(str(LIB_PATH))
    vp, f32p = C.c_void_p, C.POINTER(C.c_float)
    L.wfo_config_defaults.argtypes = [C.POINTER(WfoConfig)]
    L.wfo_create.restype = vp
    L.wfo_create.argtypes = [C.POINTER(WfoConfig)]
    L.wfo_destroy.argtypes = [vp]
    L.wfo_reset.argtypes = [vp]
    for n in ("bins", "display_channels", "num_points", "last_silent"):
        getattr(L, "wfo_" + n).restype = C.c_int
        getattr(L, "wfo_" + n).argtypes = [vp]
    L.wfo_window_sum.restype = C.c_float
    L.wfo_window_sum.argtypes = [vp]
    L.wfo_db_min.restype = C.c_float
    L.wfo_gravity.restype = C.c_float
    L.wfo_gravity.argtypes = [vp, C.c_float]
    for n in ("window", "slope", "rolloff", "interp_indices"):
        getattr(L, "wfo_get_" + n).argtypes = [vp, f32p]
    L.wfo_get_band_widths.argtypes = [vp, C.POINTER(C.c_int32)]
    L.wfo_get_interp_weights.argtypes = [vp, f32p, C.POINTER(C.c_int)]
    L.wfo_get_gauss_kernel.argtypes = [vp, f32p, C.POINTER(C.c_int), f32p]
    L.wfo_tick.argtypes = [vp, C.POINTER(f32p), C.c_float, C.c_float]
    L.wfo_decibels.restype = f32p
    L.wfo_decibels.argtypes = [vp, C.c_int]
    L.wfo_tsmooth.restype = f32p
    L.wfo_tsmooth.argtypes = [vp, C.c_int]
    L.wfo_set_state.argtypes = [vp, C.c_int, f32p, f32p]
    L.wfo_interp.argtypes = [vp, C.c_int, f32p]
    L.wfo_run_stft.argtypes = [vp, f32p, f32p, C.c_int, C.c_int, C.c_float, f32p, f32p, f32p, C.POINTER(C.c_ubyte)]
    L.wfo_run_stft_px.argtypes = [vp, f32p, f32p, C.c_int, C.c_int, C.c_float, f32p, f32p, f32p, C.POINTER(C.c_ubyte),
                                  f32p, f32p]
    L.wfo_render_pixels.argtypes = [vp, f32p, f32p, f32p]
    L.wfo_r2c.argtypes = [f32p, C.c_int, f32p]
    L.wfo_wave_create.restype = vp
    L.wfo_wave_create.argtypes = [C.POINTER(WfoWaveConfig)]
    L.wfo_wave_destroy.argtypes = [vp]
    L.wfo_wave_run.argtypes = [vp, f32p, f32p, C.c_int, C.c_int, f32p, f32p, C.POINTER(C.c_ubyte)]
    L.wfo_meter_window.argtypes = [C.POINTER(WfoMeterConfig)]
    L.wfo_meter_create.restype = vp
    L.wfo_meter_create.argtypes = [C.POINTER(WfoMeterConfig)]
    L.wfo_meter_destroy.argtypes = [vp]
    L.wfo_meter_reset.argtypes = [vp]
    L.wfo_meter_run.argtypes = [vp, f32p, f32p, C.c_int, C.c_int, C.c_float, f32p, f32p, C.POINTER(C.c_ubyte), f32p]
    _lib = L
    return L


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def config_from_settings(settings: dict | None = None, sample_rate=48000, channels=2) -> WfoConfig:
    """Translate reference setting keys (src/settings.hpp) into the oracle's POD config."""
    L = lib()
    c = WfoConfig()
    L.wfo_config_defaults(C.byref(c))
    c.sample_rate = sample_rate
    s = dict(settings or {})
    mode = s.pop("channel_mode", "mono")
    c.stereo = int(mode == "stereo")
    c.capture_channels = min(channels, 2) if mode != "single" else min(channels, 1)
    simple = {
        "fft_size": "fft_size", "sine_exponent": "sine_exponent", "gravity": "gravity", "fast_peaks": "fast_peaks",
        "slope": "slope", "rolloff_q": "rolloff_q", "rolloff_rate": "rolloff_rate", "cutoff_low": "cutoff_low",
        "cutoff_high": "cutoff_high", "floor": "floor_db", "ceiling": "ceiling_db",
        "normalize_volume": "normalize_volume", "volume_target": "volume_target", "max_gain": "max_gain",
        "width": "width", "bar_width": "bar_width", "bar_gap": "bar_gap", "log_scale": "log_scale",
        "mirror_freq_axis": "mirror_freq_axis", "filter_radius": "filter_radius", "silence_gate": "silence_gate",
        "height": "height", "channel_spacing": "channel_spacing", "rounded_caps": "rounded_caps",
        "min_bar_height": "min_bar_height",
    }
    enums = {"window": ("window", WINDOWS), "interp_mode": ("interp_mode", INTERPS),
             "filter_mode": ("filter_mode", FILTERS), "temporal_smoothing": ("tsmoothing", TSMOOTH),
             "display_mode": ("display_mode", DISPLAYS)}
    for k, v in s.items():
        if k in simple:
            field = simple[k]
            cur = getattr(c, field)
            setattr(c, field, type(cur)(v) if not isinstance(cur, float) else float(v))
        elif k in enums:
            field, table = enums[k]
            setattr(c, field, table.get(v, 0))
        elif k in ("auto_fft_size", "audio_sync_offset"):
            pass
        else:
            raise KeyError(f"unsupported setting for the oracle: {k}")
    return c


class OracleSource:
    def __init__(self, settings: dict | None = None, sample_rate=48000, channels=2, config: WfoConfig | None = None):
        self.L = lib()
        self.cfg = config if config is not None else config_from_settings(settings, sample_rate, channels)
        self.h = self.L.wfo_create(C.byref(self.cfg))

    def close(self):
        if self.h:
            self.L.wfo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def bins(self):
        return self.L.wfo_bins(self.h)

    @property
    def display_channels(self):
        return self.L.wfo_display_channels(self.h)

    @property
    def num_points(self):
        return self.L.wfo_num_points(self.h)

    @property
    def last_silent(self):
        return bool(self.L.wfo_last_silent(self.h))

    @property
    def window_sum(self):
        return float(self.L.wfo_window_sum(self.h))

    @property
    def db_min(self):
        return float(self.L.wfo_db_min())

    def gravity(self, seconds):
        return float(self.L.wfo_gravity(self.h, seconds))

    def reset(self):
        self.L.wfo_reset(self.h)

    def _vec(self, getter, dtype=np.float32):
        n = getter(self.h, None)
        if n == 0:
            return None
        out = np.zeros(n, dtype=dtype)
        getter(self.h, out.ctypes.data_as(C.POINTER(C.c_float if dtype == np.float32 else C.c_int32)))
        return out

    def window(self):
        return self._vec(self.L.wfo_get_window)

    def slope(self):
        return self._vec(self.L.wfo_get_slope)

    def rolloff(self):
        return self._vec(self.L.wfo_get_rolloff)

    def interp_indices(self):
        return self._vec(self.L.wfo_get_interp_indices)

    def band_widths(self):
        return self._vec(self.L.wfo_get_band_widths, np.int32)

    def interp_kernel(self):
        taps = C.c_int(0)
        n = self.L.wfo_get_interp_weights(self.h, None, C.byref(taps))
        if n == 0:
            return 0, None
        out = np.zeros(n, dtype=np.float32)
        self.L.wfo_get_interp_weights(self.h, _fp(out), C.byref(taps))
        return taps.value, out.reshape(-1, taps.value)

    def gauss_kernel(self):
        radius, ksum = C.c_int(0), C.c_float(0)
        n = self.L.wfo_get_gauss_kernel(self.h, None, C.byref(radius), C.byref(ksum))
        out = np.zeros(max(n, 1), dtype=np.float32)
        if n:
            self.L.wfo_get_gauss_kernel(self.h, _fp(out), C.byref(radius), C.byref(ksum))
        return out[:n], radius.value, ksum.value

    def tick(self, frames, seconds=1.0 / 60.0, input_rms=0.0):
        """frames: list of per-channel arrays (or None for 'not enough audio')."""
        arrs = [None if f is None else np.ascontiguousarray(f, dtype=np.float32) for f in frames]
        ptrs = (C.POINTER(C.c_float) * 2)()
        for i in range(2):
            ptrs[i] = _fp(arrs[i]) if i < len(arrs) and arrs[i] is not None else None
        self.L.wfo_tick(self.h, ptrs, seconds, input_rms)

    def decibels(self, ch=0):
        p = self.L.wfo_decibels(self.h, ch)
        return np.ctypeslib.as_array(p, shape=(self.bins,)).copy()

    def tsmooth(self, ch=0):
        p = self.L.wfo_tsmooth(self.h, ch)
        return np.ctypeslib.as_array(p, shape=(self.bins,)).copy() if p else None

    def interp(self, ch=0):
        out = np.zeros(self.num_points, dtype=np.float32)
        self.L.wfo_interp(self.h, ch, _fp(out))
        return out

    def interp_of(self, db):
        """Display points the reference's interpolation (+ Gaussian) yields for GIVEN dB spectra db[..., dch, B]
        (loads them into m_decibels, then src/filter.hpp:133-211 as restated in wfo_interp)."""
        db = np.ascontiguousarray(db, dtype=np.float32)
        lead = db.shape[:-2]
        flat = db.reshape(-1, db.shape[-2], db.shape[-1])
        out = np.zeros((flat.shape[0], flat.shape[1], self.num_points), dtype=np.float32)
        for i in range(flat.shape[0]):
            for ch in range(flat.shape[1]):
                row = np.ascontiguousarray(flat[i, ch])
                self.L.wfo_set_state(self.h, ch, None, _fp(row))
                self.L.wfo_interp(self.h, ch, _fp(out[i, ch]))
        return out.reshape(*lead, flat.shape[1], self.num_points)

    def render_pixels(self):
        """Pixel heights render_curve / render_bars would leave in m_interp_bufs for the current m_decibels, + (miny, minpos)."""
        out = np.zeros((self.display_channels, self.num_points), dtype=np.float32)
        miny, minpos = C.c_float(0), C.c_float(0)
        self.L.wfo_render_pixels(self.h, _fp(out), C.byref(miny), C.byref(minpos))
        return out, miny.value, minpos.value

    def pixels_of(self, db):
        """(pixels[..., dch, P], min[..., 2]) for GIVEN dB spectra db[..., dch, B]."""
        db = np.ascontiguousarray(db, dtype=np.float32)
        lead = db.shape[:-2]
        flat = db.reshape(-1, db.shape[-2], db.shape[-1])
        px = np.zeros((flat.shape[0], flat.shape[1], self.num_points), dtype=np.float32)
        mn = np.zeros((flat.shape[0], 2), dtype=np.float32)
        for i in range(flat.shape[0]):
            for ch in range(flat.shape[1]):
                self.L.wfo_set_state(self.h, ch, None, _fp(np.ascontiguousarray(flat[i, ch])))
            p, a, b = self.render_pixels()
            px[i], mn[i] = p, (a, b)
        return px.reshape(*lead, flat.shape[1], self.num_points), mn.reshape(*lead, 2)

    def run_stft(self, pcm, n_frames, hop, seconds=1.0 / 60.0, rms=None, want_db=True, want_points=False):
        pcm = np.ascontiguousarray(np.atleast_2d(pcm), dtype=np.float32)
        ch0 = pcm[0]
        ch1 = pcm[1] if pcm.shape[0] > 1 else None
        T, dch, B = n_frames, self.display_channels, self.bins
        need = (T - 1) * hop + 2 * B
        assert pcm.shape[1] >= need, (pcm.shape, need)
        db = np.zeros((T, dch, B), dtype=np.float32) if want_db else None
        pts = np.zeros((T, dch, self.num_points), dtype=np.float32) if want_points else None
        silent = np.zeros(T, dtype=np.uint8)
        if rms is not None:
            rms = np.ascontiguousarray(rms, dtype=np.float32)
        self.L.wfo_run_stft(self.h, _fp(ch0), _fp(ch1), T, hop, seconds, _fp(rms), _fp(db), _fp(pts),
                            silent.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return {"frames": T, "db": db, "points": pts, "silent": silent}


def r2c(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = len(x)
    out = np.zeros(2 * (n // 2 + 1), dtype=np.float32)
    lib().wfo_r2c(_fp(x), n, _fp(out))
    return out.view(np.complex64)


def meter_config_from_settings(settings: dict | None = None, sample_rate=48000, channels=2) -> WfoMeterConfig:
    """Reference setting keys (src/settings.hpp) -> the level meter's POD config (defaults: src/source.cpp:119-174)."""
    s = dict(settings or {})
    c = WfoMeterConfig()
    c.sample_rate = sample_rate
    c.capture_channels = min(channels, 2)
    c.meter_ms = int(s.get("meter_buf", 150))
    c.rms_mode = int(bool(s.get("rms_mode", True)))
    c.tsmoothing = TSMOOTH.get(s.get("temporal_smoothing", "exp_moving_avg"), 0)
    c.gravity = float(s.get("gravity", 0.65))
    c.fast_peaks = int(bool(s.get("fast_peaks", False)))
    c.floor_db = int(s.get("floor", -65))
    return c


class OracleMeter:
    """Level meter (tick_meter) and RMS feed (update_input_rms) restated; see wf_oracle_meter.c."""

    def __init__(self, settings: dict | None = None, sample_rate=48000, channels=2):
        self.L = lib()
        self.cfg = meter_config_from_settings(settings, sample_rate, channels)
        self.h = self.L.wfo_meter_create(C.byref(self.cfg))
        self.window = int(self.L.wfo_meter_window(C.byref(self.cfg)))

    def __del__(self):
        try:
            if self.h:
                self.L.wfo_meter_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset(self):
        self.L.wfo_meter_reset(self.h)

    def run(self, pcm: np.ndarray, n_ticks: int, hop: int, seconds: float = 1.0 / 60.0, meter=True, rms=False):
        pcm = np.ascontiguousarray(np.atleast_2d(pcm), dtype=np.float32)
        cc = self.cfg.capture_channels
        assert pcm.shape[0] >= cc and pcm.shape[1] >= n_ticks * hop
        db = np.zeros((n_ticks, cc), np.float32) if meter else None
        lin = np.zeros((n_ticks, cc), np.float32) if meter else None
        sil = np.zeros(n_ticks, np.uint8) if meter else None
        r = np.zeros(n_ticks, np.float32) if rms else None
        self.L.wfo_meter_run(self.h, _fp(pcm[0]), _fp(pcm[1]) if cc > 1 else None, n_ticks, hop, seconds, _fp(db), _fp(lin),
                             None if sil is None else sil.ctypes.data_as(C.POINTER(C.c_ubyte)), _fp(r))
        return {"db": db, "lin": lin, "silent": sil, "rms": r}


def wave_config_from_settings(settings: dict | None = None, sample_rate=48000, channels=2) -> WfoWaveConfig:
    """Reference setting keys -> the waveform mode's POD config (defaults src/source.cpp:119-174)."""
    s = dict(settings or {})
    c = WfoWaveConfig()
    c.sample_rate = sample_rate
    mode = s.get("channel_mode", "mono")
    c.stereo = int(mode == "stereo")
    c.capture_channels = min(channels, 2) if mode != "single" else 1
    c.width = int(s.get("width", 800))
    c.meter_ms = int(s.get("meter_buf", 150))
    c.normalize_volume = int(bool(s.get("normalize_volume", False)))
    c.volume_target = float(s.get("volume_target", -8.0))
    c.max_gain = float(s.get("max_gain", 30.0))
    return c


class OracleWave:
    """tick_waveform restated (wf_oracle_meter.c)."""

    def __init__(self, settings: dict | None = None, sample_rate=48000, channels=2):
        self.L = lib()
        self.cfg = wave_config_from_settings(settings, sample_rate, channels)
        self.h = self.L.wfo_wave_create(C.byref(self.cfg))

    def __del__(self):
        try:
            if self.h:
                self.L.wfo_wave_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def run(self, pcm: np.ndarray, n_ticks: int, hop: int, rms: np.ndarray | None = None):
        pcm = np.ascontiguousarray(np.atleast_2d(pcm), dtype=np.float32)
        cc, dch = self.cfg.capture_channels, (2 if self.cfg.stereo else 1)
        out = np.zeros((n_ticks, dch, self.cfg.width), np.float32)
        sil = np.zeros(n_ticks, np.uint8)
        if rms is not None:
            rms = np.ascontiguousarray(rms, dtype=np.float32)
        self.L.wfo_wave_run(self.h, _fp(pcm[0]), _fp(pcm[1]) if cc > 1 else None, n_ticks, hop, _fp(rms), _fp(out),
                            sil.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return {"out": out, "silent": sil}
