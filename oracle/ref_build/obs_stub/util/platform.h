// Fake libobs util/platform.h (parity-oracle test infrastructure; see obs-module.h).
#pragma once
#include <cstdint>
#ifdef __cplusplus
extern "C" {
#endif
uint64_t os_gettime_ns(void);   // controllable fake clock, set by ref_harness.cpp
#ifdef __cplusplus
}
#endif
