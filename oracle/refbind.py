"""ctypes binding for oracle/_ref/libwaveform_ref.so — the UNMODIFIED reference compiled here.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's reference /
cpu_baseline legs may import this.  The product path (waveform_b200/) never does.

The library is built by `make -C oracle/ref_build` from the sources where they lie under
/root/reference (see that Makefile).  Settings are passed with the reference's own setting keys
(/root/reference/src/settings.hpp:29-135), e.g. {"fft_size": 2048, "window": "hann"}.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "_ref" / "libwaveform_ref.so"
# The same unmodified reference objects + the product's plugin-side binding (waveform_b200/host/source_cuda.hpp:
# WAVSourceCUDA, impl 3).  A separate library on purpose: the CPU baseline / oracle (LIB_PATH) never loads the product.
CUDA_LIB_PATH = _HERE / "_ref" / "libwaveform_ref_cuda.so"

IMPL_GENERIC, IMPL_AVX, IMPL_AVX2, IMPL_CUDA = 0, 1, 2, 3

_lib = None
_lib_cuda = None


def available() -> bool:
    return LIB_PATH.exists()


def cuda_seam_available() -> bool:
    return CUDA_LIB_PATH.exists()


def lib(cuda: bool = False):
    global _lib, _lib_cuda
    if cuda and _lib_cuda is not None:
        return _lib_cuda
    if not cuda and _lib is not None:
        return _lib
    path = CUDA_LIB_PATH if cuda else LIB_PATH
    if not path.exists():
        raise FileNotFoundError(
            f"{path} missing: run `make -C oracle/ref_build -j8` where /root/reference exists")
    L = C.CDLL(str(path))
    vp, f32p, i32p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.wfref_create.restype = vp
    L.wfref_create.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32]
    L.wfref_destroy.argtypes = [vp]
    L.wfref_set_int.argtypes = [vp, C.c_char_p, C.c_longlong]
    L.wfref_set_double.argtypes = [vp, C.c_char_p, C.c_double]
    L.wfref_set_bool.argtypes = [vp, C.c_char_p, C.c_int]
    L.wfref_set_string.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.wfref_update.argtypes = [vp]
    L.wfref_set_showing.argtypes = [vp, C.c_int]
    L.wfref_advance_clock_ns.argtypes = [vp, C.c_uint64]
    L.wfref_clock_ns.restype = C.c_uint64
    L.wfref_clock_ns.argtypes = [vp]
    L.wfref_push_audio.argtypes = [vp, f32p, f32p, C.c_uint32, C.c_int64]
    L.wfref_tick.argtypes = [vp, C.c_float]
    L.wfref_render.argtypes = [vp]
    L.wfref_fft_size.restype = C.c_uint64
    L.wfref_fft_size.argtypes = [vp]
    for name in ("capture_channels", "output_channels", "width"):
        getattr(L, "wfref_" + name).restype = C.c_uint32
        getattr(L, "wfref_" + name).argtypes = [vp]
    for name in ("stereo", "last_silent", "num_bars"):
        getattr(L, "wfref_" + name).restype = C.c_int
        getattr(L, "wfref_" + name).argtypes = [vp]
    for name in ("window_sum", "db_min", "input_rms"):
        getattr(L, "wfref_" + name).restype = C.c_float
        getattr(L, "wfref_" + name).argtypes = [vp]
    L.wfref_gravity.restype = C.c_float
    L.wfref_gravity.argtypes = [vp, C.c_float]
    L.wfref_force_input_rms.argtypes = [vp, C.c_int, C.c_float]
    L.wfref_get_decibels.argtypes = [vp, C.c_int, f32p]
    L.wfref_get_tsmooth.argtypes = [vp, C.c_int, f32p]
    for name in ("window", "slope", "rolloff", "interp_indices"):
        getattr(L, "wfref_get_" + name).argtypes = [vp, f32p]
    L.wfref_get_band_widths.argtypes = [vp, i32p]
    L.wfref_get_interp_kernel.argtypes = [vp, f32p, C.c_int]
    L.wfref_get_gauss_kernel.argtypes = [vp, f32p, i32p, f32p]
    L.wfref_interp.argtypes = [vp, C.c_int, C.c_int, f32p]
    L.wfref_get_render_buf.argtypes = [vp, C.c_int, f32p]
    L.wfref_run_stft.argtypes = [vp, f32p, f32p, C.c_int64, C.c_int, C.c_int, C.c_float, f32p, f32p, f32p,
                                 C.c_int, C.POINTER(C.c_ubyte)]
    L.wfref_run_wave.argtypes = [vp, f32p, f32p, C.c_int, C.c_int, C.c_float, f32p, f32p, C.POINTER(C.c_ubyte)]
    L.wfref_run_meter.argtypes = [vp, f32p, f32p, C.c_int, C.c_int, C.c_float, f32p, f32p, C.POINTER(C.c_ubyte), f32p]
    if cuda:
        _lib_cuda = L
    else:
        _lib = L
    return L


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


class RefSource:
    """One reference WAVSource{Generic,AVX,AVX2} instance behind the fake libobs."""

    def __init__(self, settings: dict | None = None, impl: int = IMPL_GENERIC, sample_rate: int = 48000,
                 channels: int = 2, fps: tuple[int, int] = (60, 1)):
        self.L = lib(cuda=(impl == IMPL_CUDA))
        self.h = self.L.wfref_create(impl, sample_rate, channels, fps[0], fps[1])
        if not self.h:
            raise RuntimeError(f"wfref_create(impl={impl}) failed")
        self.sample_rate = sample_rate
        self.channels = channels
        if settings:
            self.set(**settings)
        self.update()

    def close(self):
        if self.h:
            self.L.wfref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, **kv):
        for k, v in kv.items():
            kb = k.encode()
            if isinstance(v, bool):
                self.L.wfref_set_bool(self.h, kb, int(v))
            elif isinstance(v, int):
                self.L.wfref_set_int(self.h, kb, v)
            elif isinstance(v, float):
                self.L.wfref_set_double(self.h, kb, v)
            elif isinstance(v, str):
                self.L.wfref_set_string(self.h, kb, v.encode())
            else:
                raise TypeError(f"{k}: {type(v)}")

    def update(self):
        self.L.wfref_update(self.h)

    # ---- properties ----
    @property
    def fft_size(self) -> int:
        return int(self.L.wfref_fft_size(self.h))

    @property
    def bins(self) -> int:
        return self.fft_size // 2

    @property
    def capture_channels(self) -> int:
        return int(self.L.wfref_capture_channels(self.h))

    @property
    def output_channels(self) -> int:
        return int(self.L.wfref_output_channels(self.h))

    @property
    def stereo(self) -> bool:
        return bool(self.L.wfref_stereo(self.h))

    @property
    def display_channels(self) -> int:
        return 2 if self.stereo else 1

    @property
    def last_silent(self) -> bool:
        return bool(self.L.wfref_last_silent(self.h))

    @property
    def window_sum(self) -> float:
        return float(self.L.wfref_window_sum(self.h))

    @property
    def db_min(self) -> float:
        return float(self.L.wfref_db_min(self.h))

    @property
    def num_bars(self) -> int:
        return int(self.L.wfref_num_bars(self.h))

    @property
    def width(self) -> int:
        return int(self.L.wfref_width(self.h))

    def gravity(self, seconds: float) -> float:
        return float(self.L.wfref_gravity(self.h, seconds))

    def _vec(self, getter, n, *pre):
        out = np.zeros(n, dtype=np.float32)
        got = getter(self.h, *pre, _fp(out))
        return out if got else None

    def decibels(self, ch=0):
        return self._vec(self.L.wfref_get_decibels, self.bins, ch)

    def tsmooth(self, ch=0):
        return self._vec(self.L.wfref_get_tsmooth, self.bins, ch)

    def window(self):
        return self._vec(self.L.wfref_get_window, self.fft_size)

    def slope(self):
        return self._vec(self.L.wfref_get_slope, self.bins)

    def rolloff(self):
        return self._vec(self.L.wfref_get_rolloff, self.bins)

    def interp_indices(self):
        n = self.L.wfref_get_interp_indices(self.h, None)
        out = np.zeros(n, dtype=np.float32)
        self.L.wfref_get_interp_indices(self.h, _fp(out))
        return out

    def band_widths(self):
        n = self.L.wfref_get_band_widths(self.h, None)
        out = np.zeros(n, dtype=np.int32)
        if n:
            self.L.wfref_get_band_widths(self.h, out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def interp_kernel(self):
        """(taps, weights[points, taps]) of the Lanczos / Catmull-Rom table, or (0, None)."""
        taps = self.L.wfref_get_interp_kernel(self.h, None, 0)
        if taps == 0:
            return 0, None
        n = len(self.interp_indices()) * taps
        out = np.zeros(n, dtype=np.float32)
        self.L.wfref_get_interp_kernel(self.h, _fp(out), n)
        return taps, out.reshape(-1, taps)

    def gauss_kernel(self):
        radius, ksum = C.c_int(0), C.c_float(0)
        size = self.L.wfref_get_gauss_kernel(self.h, None, C.byref(radius), C.byref(ksum))
        out = np.zeros(max(size, 1), dtype=np.float32)
        if size:
            self.L.wfref_get_gauss_kernel(self.h, _fp(out), C.byref(radius), C.byref(ksum))
        return out[:size], radius.value, ksum.value

    # ---- driving ----
    def set_showing(self, showing: bool):
        self.L.wfref_set_showing(self.h, int(showing))

    def advance(self, seconds: float):
        self.L.wfref_advance_clock_ns(self.h, int(round(seconds * 1e9)))

    def push(self, ch0: np.ndarray, ch1: np.ndarray | None = None, ts_adjust_ns: int = 0):
        ch0 = np.ascontiguousarray(ch0, dtype=np.float32)
        if ch1 is not None:
            ch1 = np.ascontiguousarray(ch1, dtype=np.float32)
        self.L.wfref_push_audio(self.h, _fp(ch0), _fp(ch1), len(ch0), ts_adjust_ns)

    def tick(self, seconds: float = 1.0 / 60.0):
        self.L.wfref_tick(self.h, seconds)

    def force_input_rms(self, value: float | None):
        self.L.wfref_force_input_rms(self.h, int(value is not None), float(value or 0.0))

    def interp(self, ch=0, fma3=False):
        n = self.L.wfref_interp(self.h, ch, int(fma3), None)
        out = np.zeros(n, dtype=np.float32)
        self.L.wfref_interp(self.h, ch, int(fma3), _fp(out))
        return out

    def render(self):
        self.L.wfref_render(self.h)

    def render_buf(self, ch=0):
        n = self.L.wfref_get_render_buf(self.h, ch, None)
        out = np.zeros(n, dtype=np.float32)
        self.L.wfref_get_render_buf(self.h, ch, _fp(out))
        return out

    def run_stft(self, pcm: np.ndarray, n_frames: int, hop: int, seconds: float = 1.0 / 60.0,
                 rms: np.ndarray | None = None, want_db=True, want_points=False, fma3=False):
        """pcm: [capture_channels, samples] float32.  Returns dict(db=[T,dch,B], points=[T,dch,P], silent=[T])."""
        pcm = np.ascontiguousarray(np.atleast_2d(pcm), dtype=np.float32)
        ch0 = pcm[0]
        ch1 = pcm[1] if pcm.shape[0] > 1 else None
        T, dch, B = n_frames, self.display_channels, self.bins
        db = np.zeros((T, dch, B), dtype=np.float32) if want_db else None
        pts = None
        if want_points:
            npts = self.L.wfref_interp(self.h, 0, int(fma3), None)
            pts = np.zeros((T, dch, npts), dtype=np.float32)
        silent = np.zeros(T, dtype=np.uint8)
        if rms is not None:
            rms = np.ascontiguousarray(rms, dtype=np.float32)
        done = self.L.wfref_run_stft(self.h, _fp(ch0), _fp(ch1), pcm.shape[1], T, hop, seconds, _fp(rms), _fp(db),
                                     _fp(pts), int(fma3), silent.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return {"frames": done, "db": db, "points": pts, "silent": silent}

    def run_meter(self, pcm: np.ndarray, n_ticks: int, hop: int, seconds: float = 1.0 / 60.0):
        """Level-meter ticks (display_mode level_meter/stepped_meter) or, with normalize_volume on in a spectrum
        mode, the per-tick m_input_rms.  pcm: [capture_channels, >= n_ticks*hop].
        Returns dict(db=[T,cc] m_meter_val, lin=[T,cc] m_meter_buf, silent=[T], rms=[T] m_input_rms)."""
        pcm = np.ascontiguousarray(np.atleast_2d(pcm), dtype=np.float32)
        assert pcm.shape[1] >= n_ticks * hop
        ch0 = pcm[0]
        ch1 = pcm[1] if pcm.shape[0] > 1 else None
        cc = self.capture_channels
        db = np.zeros((n_ticks, cc), dtype=np.float32)
        lin = np.zeros((n_ticks, cc), dtype=np.float32)
        silent = np.zeros(n_ticks, dtype=np.uint8)
        rms = np.zeros(n_ticks, dtype=np.float32)
        self.L.wfref_run_meter(self.h, _fp(ch0), _fp(ch1), n_ticks, hop, seconds, _fp(db), _fp(lin),
                               silent.ctypes.data_as(C.POINTER(C.c_ubyte)), _fp(rms))
        return {"db": db, "lin": lin, "silent": silent, "rms": rms}

    def run_wave(self, pcm: np.ndarray, n_ticks: int, hop: int, seconds: float = 1.0 / 60.0, rms: np.ndarray | None = None):
        """Waveform (oscilloscope) mode (display_mode "waveform"): out=[T, display_channels, width] m_decibels after each tick."""
        pcm = np.ascontiguousarray(np.atleast_2d(pcm), dtype=np.float32)
        assert pcm.shape[1] >= n_ticks * hop
        ch1 = pcm[1] if pcm.shape[0] > 1 else None
        out = np.zeros((n_ticks, self.display_channels, self.fft_size), dtype=np.float32)
        silent = np.zeros(n_ticks, dtype=np.uint8)
        if rms is not None:
            rms = np.ascontiguousarray(rms, dtype=np.float32)
        self.L.wfref_run_wave(self.h, _fp(pcm[0]), _fp(ch1), n_ticks, hop, seconds, _fp(rms), _fp(out),
                              silent.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return {"out": out, "silent": silent}
