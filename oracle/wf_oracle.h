/* wf_oracle.h — CPU restatement of phandasm/waveform's spectrum path (the parity ORACLE).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may compile, link or call this.  The product (waveform_b200/,
 * include/wfstft.h, libwfstft.so) never does and fails loudly without its CUDA library.
 *
 * Pinning status: the reference has no tests/golden vectors of its own for this path
 * (SURVEY.md §4, §8c), so this restatement is pinned against the UNMODIFIED reference compiled
 * here (oracle/_ref/libwaveform_ref.so, see oracle/ref_build/Makefile) and against golden fixtures
 * generated from it (tests/golden/, tests/golden/make_golden.py).
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 */
#ifndef WF_ORACLE_H
#define WF_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enums: src/source.hpp:32-93 (same order) */
enum { WFO_WINDOW_NONE, WFO_WINDOW_HANN, WFO_WINDOW_HAMMING, WFO_WINDOW_BLACKMAN, WFO_WINDOW_BLACKMAN_HARRIS,
       WFO_WINDOW_POWER_OF_SINE };
enum { WFO_INTERP_POINT, WFO_INTERP_LANCZOS, WFO_INTERP_CATROM };
enum { WFO_FILTER_NONE, WFO_FILTER_GAUSS };
enum { WFO_TSMOOTH_NONE, WFO_TSMOOTH_EXPONENTIAL, WFO_TSMOOTH_TVEXPONENTIAL };
enum { WFO_DISPLAY_CURVE, WFO_DISPLAY_BAR };

/* DSP-relevant subset of the plugin settings (src/source.cpp:501-674 after clamping). */
typedef struct wfo_config {
    uint32_t sample_rate;     /* obs_audio_info.samples_per_sec */
    int32_t capture_channels; /* m_capture_channels: 1 or 2 */
    int32_t fft_size;         /* m_fft_size (>=128, multiple of 16) */
    int32_t window;
    int32_t sine_exponent;
    int32_t tsmoothing;
    float gravity;
    int32_t fast_peaks;
    float slope;
    float rolloff_q, rolloff_rate;
    int32_t cutoff_low, cutoff_high;
    int32_t floor_db, ceiling_db;
    int32_t stereo; /* m_stereo */
    int32_t normalize_volume;
    float volume_target, max_gain;
    int32_t silence_gate; /* 1 = reference behaviour (src/source_generic.cpp:63-95) */
    int32_t display_mode;
    int32_t width, bar_width, bar_gap;
    int32_t log_scale, mirror_freq_axis;
    int32_t interp_mode;
    int32_t filter_mode;
    float filter_radius;
    int32_t height, channel_spacing, rounded_caps, min_bar_height; /* display stage */
} wfo_config;

typedef struct wfo_source wfo_source; /* one WAVSource worth of state */

void wfo_config_defaults(wfo_config *cfg); /* src/source.cpp:119-174 */

wfo_source *wfo_create(const wfo_config *cfg); /* ≙ WAVSource::update, src/source.cpp:1077-1322 */
void wfo_destroy(wfo_source *s);
void wfo_reset(wfo_source *s); /* ≙ timeout/hidden branch, src/source_generic.cpp:36-48 */

int wfo_bins(const wfo_source *s);             /* N/2 */
int wfo_display_channels(const wfo_source *s); /* m_stereo ? 2 : 1 */
int wfo_num_points(const wfo_source *s);       /* width (curve) or num_bars (bars) */
int wfo_last_silent(const wfo_source *s);
float wfo_window_sum(const wfo_source *s);
float wfo_db_min(void);
float wfo_gravity(const wfo_source *s, float seconds); /* src/source.hpp:301-312 */

/* table access for tests (returns element count; copies if out != NULL) */
int wfo_get_window(const wfo_source *s, float *out);
int wfo_get_slope(const wfo_source *s, float *out);
int wfo_get_rolloff(const wfo_source *s, float *out);
int wfo_get_interp_indices(const wfo_source *s, float *out);
int wfo_get_band_widths(const wfo_source *s, int32_t *out);
int wfo_get_interp_weights(const wfo_source *s, float *out, int *taps);
int wfo_get_gauss_kernel(const wfo_source *s, float *out, int *radius, float *sum);

/* One tick_spectrum on explicit frames: frames[c] points at N samples for capture channel c
 * (NULL = "not enough audio", src/source_generic.cpp:55-61).  input_rms feeds volume normalisation.
 * Afterwards wfo_decibels(s, ch) is m_decibels[ch].                      src/source_generic.cpp:26-180 */
void wfo_tick(wfo_source *s, const float *const frames[2], float seconds, float input_rms);
const float *wfo_decibels(const wfo_source *s, int ch);
const float *wfo_tsmooth(const wfo_source *s, int ch);
void wfo_set_state(wfo_source *s, int ch, const float *tsmooth, const float *decibels);

/* Interpolation (+ Gaussian) of the current m_decibels[ch] to display points, before dB->pixel.
 * src/source.cpp:1381-1406,1510-1546; src/filter.hpp:133-211.  Returns points written. */
int wfo_interp(wfo_source *s, int ch, float *out);

/* The rest of render_curve / render_bars up to (excluding) vertex generation: interpolation + Gaussian for every display
 * channel, dB -> pixel lerp/clamp with the running (miny, minpos), then frequency-axis mirroring.
 * out: [display_channels][points]; src/source.cpp:1376-1425 (curve), :1481-1565 (bars). */
void wfo_render_pixels(wfo_source *s, float *out, float *miny, float *minpos);

/* Sliding STFT of one source: frame t = pcm[c][t*hop .. t*hop+N).  Layouts as the engine's:
 * out_db [n_frames][display_channels][bins], out_points [n_frames][display_channels][points]. */
int wfo_run_stft(wfo_source *s, const float *pcm0, const float *pcm1, int n_frames, int hop, float seconds,
                 const float *input_rms, float *out_db, float *out_points, unsigned char *out_silent);
/* same, additionally out_pixels [n_frames][display_channels][points] and out_min [n_frames][2] */
int wfo_run_stft_px(wfo_source *s, const float *pcm0, const float *pcm1, int n_frames, int hop, float seconds,
                    const float *input_rms, float *out_db, float *out_points, unsigned char *out_silent,
                    float *out_pixels, float *out_min);

/* The bare transform restated on its own (forward unnormalised r2c, N/2+1 outputs interleaved),
 * deps/fftw-3.3.11/doc/reference.texi:1926-1936, api/plan-dft-r2c-1d.c:23-26. Any N >= 2. */
void wfo_r2c(const float *in, int n, float *out_interleaved);

/* ---- level meter + RMS feed (wf_oracle_meter.c): src/source_generic.cpp:182-270, :392-403, src/source.cpp:810-836,
 *      :1105-1128, :1842-1871 ------------------------------------------------------------------------------------ */
typedef struct wfo_meter_config {
    uint32_t sample_rate;
    int32_t capture_channels; /* 1 or 2 */
    int32_t meter_ms;         /* m_meter_ms: ring length = (sample_rate * ms / 1000) & -16 */
    int32_t rms_mode;         /* m_meter_rms: 1 = RMS, 0 = peak */
    int32_t tsmoothing;
    float gravity;
    int32_t fast_peaks;
    int32_t floor_db;
} wfo_meter_config;
typedef struct wfo_meter wfo_meter;
int wfo_meter_window(const wfo_meter_config *cfg);
wfo_meter *wfo_meter_create(const wfo_meter_config *cfg);
void wfo_meter_destroy(wfo_meter *m);
void wfo_meter_reset(wfo_meter *m); /* capture-timeout branch, src/source_generic.cpp:184-199 */
void wfo_meter_tick(wfo_meter *m, const float *const x[2], size_t n, float seconds);
float wfo_meter_feed_rms(wfo_meter *m, const float *const x[2], size_t n); /* -> m_input_rms */
float wfo_meter_val(const wfo_meter *m, int c);
float wfo_meter_buf(const wfo_meter *m, int c);
int wfo_meter_last_silent(const wfo_meter *m);
void wfo_meter_run(wfo_meter *m, const float *pcm0, const float *pcm1, int n_ticks, int hop, float seconds,
                   float *out_db, float *out_lin, unsigned char *out_silent, float *out_rms);

/* ---- waveform (oscilloscope) mode: src/source_generic.cpp:272-390, capture side src/source.cpp:1817-1888 ---- */
typedef struct wfo_wave_config {
    uint32_t sample_rate;
    int32_t capture_channels; /* 1 or 2 */
    int32_t stereo;           /* m_stereo */
    int32_t width;            /* m_width: points in the scrolling buffer (m_fft_size in this mode) */
    int32_t meter_ms;         /* m_meter_ms: time span of the buffer */
    int32_t normalize_volume;
    float volume_target, max_gain;
} wfo_wave_config;
typedef struct wfo_wave wfo_wave;
wfo_wave *wfo_wave_create(const wfo_wave_config *cfg);
void wfo_wave_destroy(wfo_wave *w);
void wfo_wave_tick(wfo_wave *w, const float *const x[2], size_t n, float input_rms);
int wfo_wave_last_silent(const wfo_wave *w);
const float *wfo_wave_buffer(const wfo_wave *w, int ch);
void wfo_wave_run(wfo_wave *w, const float *pcm0, const float *pcm1, int n_ticks, int hop, const float *input_rms,
                  float *out, unsigned char *out_silent);

#ifdef __cplusplus
}
#endif
#endif
