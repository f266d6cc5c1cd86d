"""waveform_b200 — B200 (sm_100a) batched STFT engine behind phandasm/waveform's tick_spectrum seam.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C-ABI of include/wfstft.h),
host/ (C++ host-side mirror of the plugin's capture/tick plumbing) and this thin ctypes binding.
"""
from .engine import Engine, MeterEngine, WfError, load_library, make_config, make_meter_config  # noqa: F401
