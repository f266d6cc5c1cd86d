// Hooks that let ref_harness.cpp steer the fake libobs (TEST INFRASTRUCTURE ONLY).
#pragma once
#include <obs-module.h>
#include <cstdint>

extern "C" {
void wfstub_set_clock_ns(uint64_t ns);
uint64_t wfstub_get_clock_ns(void);
void wfstub_set_audio(uint32_t sample_rate, int channels);
void wfstub_set_fps(uint32_t num, uint32_t den);
void wfstub_set_showing(bool showing);
void wfstub_set_log_level(int lvl);
const obs_source_info *wfstub_registered_info(void);

obs_data_t *wfstub_data_create(void);
void wfstub_data_destroy(obs_data_t *d);
void wfstub_data_set_int(obs_data_t *d, const char *k, long long v);
void wfstub_data_set_double(obs_data_t *d, const char *k, double v);
void wfstub_data_set_bool(obs_data_t *d, const char *k, bool v);
void wfstub_data_set_string(obs_data_t *d, const char *k, const char *v);
}
