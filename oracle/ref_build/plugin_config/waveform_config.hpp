// Hand-written equivalent of what CMake generates from src/waveform_config.hpp.in
// (/root/reference/src/waveform_config.hpp.in:1-22) for an x86-64 SIMD build.
// Parity-oracle test infrastructure only. ENABLE_X86_SIMD comes from the command line.
#pragma once
#define WAVEFORM_VERSION "1.9.1"
#define WAVEFORM_ARCH "x64";
#define WAV_FORCE_INLINE __attribute__((always_inline)) inline
