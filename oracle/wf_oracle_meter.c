/* wf_oracle_meter.c — CPU restatement of the plugin's level-meter tick and of the RMS feed of the volume
 * normalisation.  TEST INFRASTRUCTURE ONLY (see wf_oracle.h).
 *
 * Restated reference code (paths relative to /root/reference):
 *   WAVSourceGeneric::tick_meter          src/source_generic.cpp:182-270
 *   meter setup in WAVSource::update      src/source.cpp:1105-1128, :1181 (ring := 0), :1236 (m_last_silent := false)
 *   WAVSource::capture_audio (RMS feed)   src/source.cpp:1842-1871
 *   WAVSource::sync_rms_buffer            src/source.cpp:810-836
 *   WAVSourceGeneric::update_input_rms    src/source_generic.cpp:392-403
 *   dbfs / get_gravity                    src/source.hpp:293-312
 * The arithmetic order is the reference's (sequential fp32 sums in RING order), so this oracle is bit-exact against
 * the compiled reference (tests/test_oracle_vs_reference.py); the CUDA path sums in blocks and is compared with a
 * relative tolerance for RMS values and bit-exactly for peak values.
 */
#include "wf_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct wfo_meter {
    wfo_meter_config cfg;
    int W;            /* ring length: m_fft_size repurposed, src/source.cpp:1121 */
    float *ring[2];   /* m_decibels[c] repurposed as the sample ring, src/source_generic.cpp:205-222 */
    size_t pos[2];    /* m_meter_pos */
    float buf[2];     /* m_meter_buf */
    float val[2];     /* m_meter_val */
    int last_silent;
    /* RMS feed */
    int RW;           /* m_input_rms_size = sample_rate & -16, src/source.cpp:1148 */
    float *rms_ring;  /* m_input_rms_buf */
    size_t rms_pos;
    float input_rms;
};

static float dbfs_m(float mag) { return (mag > 0.0f) ? 20.0f * log10f(mag) : wfo_db_min(); }

/* get_gravity, src/source.hpp:301-312 */
static float gravity_m(const wfo_meter_config *c, float seconds)
{
    const float denom = 0.03868924705242879469662125316986f;
    const float hi = denom * 5.0f;
    if((c->tsmoothing == WFO_TSMOOTH_NONE) || (c->gravity <= 0.0f))
        return 0.0f;
    if(c->tsmoothing == WFO_TSMOOTH_TVEXPONENTIAL)
    {
        /* std::lerp(0, hi, g): a == 0 -> t*b + (1-t)*a */
        const float l = c->gravity * hi + (1.0f - c->gravity) * 0.0f;
        return expf(-seconds / l);
    }
    return c->gravity;
}

int wfo_meter_window(const wfo_meter_config *c)
{
    /* src/source.cpp:1121: size_t(samples_per_sec * (meter_ms / 1000.0)) & -16 */
    return (int)(((size_t)((double)c->sample_rate * ((double)c->meter_ms / 1000.0))) & ~(size_t)15);
}

wfo_meter *wfo_meter_create(const wfo_meter_config *cfg)
{
    wfo_meter *m = (wfo_meter *)calloc(1, sizeof(*m));
    m->cfg = *cfg;
    m->W = wfo_meter_window(cfg);
    m->RW = (int)(cfg->sample_rate & ~15u);
    for(int c = 0; c < 2; ++c)
    {
        m->ring[c] = (float *)calloc((size_t)(m->W > 0 ? m->W : 1), sizeof(float)); /* src/source.cpp:1181 */
        m->buf[c] = wfo_db_min();                                                     /* :1124-1127 (sic: DB_MIN) */
        m->val[c] = wfo_db_min();
    }
    m->rms_ring = (float *)calloc((size_t)m->RW, sizeof(float)); /* :1152 */
    m->last_silent = 0;                                          /* :1236 */
    return m;
}

void wfo_meter_destroy(wfo_meter *m)
{
    if(!m)
        return;
    free(m->ring[0]);
    free(m->ring[1]);
    free(m->rms_ring);
    free(m);
}

/* capture timeout branch, src/source_generic.cpp:184-199 */
void wfo_meter_reset(wfo_meter *m)
{
    if(m->last_silent)
        return;
    for(int c = 0; c < m->cfg.capture_channels; ++c)
        memset(m->ring[c], 0, (size_t)m->W * sizeof(float));
    for(int c = 0; c < 2; ++c)
    {
        m->buf[c] = 0.0f;
        m->val[c] = wfo_db_min();
    }
    m->last_silent = 1;
}

static void ring_push(float *ring, int W, size_t *pos, const float *x, size_t n)
{
    /* the pop_front loop of src/source_generic.cpp:205-222 (and src/source.cpp:818-832) with `n` samples pending */
    while(n > 0)
    {
        size_t max = (size_t)W - *pos;
        if(n >= max)
        {
            memcpy(ring + *pos, x, max * sizeof(float));
            *pos = 0;
            x += max;
            n -= max;
        }
        else
        {
            memcpy(ring + *pos, x, n * sizeof(float));
            *pos += n;
            n = 0;
        }
    }
}

/* One tick_meter with `n` new samples per channel (get_audio_sync() == 0: everything captured is consumed). */
void wfo_meter_tick(wfo_meter *m, const float *const x[2], size_t n, float seconds)
{
    const int cc = m->cfg.capture_channels;
    const int W = m->W;
    /* capture_audio keeps at most m_fft_size pending samples per channel (src/source.cpp:1881-1884) */
    const size_t skip = (n > (size_t)W) ? n - (size_t)W : 0;
    for(int c = 0; c < cc; ++c)
        ring_push(m->ring[c], W, &m->pos[c], x[c] + skip, n - skip);
    for(int c = 0; c < cc; ++c)
    {
        float out = 0.0f;
        if(m->cfg.rms_mode)
        {
            for(int i = 0; i < W; ++i)
            {
                const float v = m->ring[c][i];
                out += v * v;
            }
            out = sqrtf(out / (float)W);
        }
        else
        {
            for(int i = 0; i < W; ++i)
                out = fmaxf(out, fabsf(m->ring[c][i]));
        }
        if(m->cfg.tsmoothing != WFO_TSMOOTH_NONE)
        {
            const float g = gravity_m(&m->cfg, seconds);
            const float g2 = 1.0f - g;
            if(!m->cfg.fast_peaks || (out <= m->buf[c]))
                out = (g * m->buf[c]) + (g2 * out);
        }
        m->buf[c] = out;
        m->val[c] = dbfs_m(out);
    }
    int silent_channels = 0;
    for(int c = 0; c < cc; ++c)
        if(m->val[c] < (float)(m->cfg.floor_db - 10))
            ++silent_channels;
    m->last_silent = (silent_channels >= cc);
}

/* capture_audio's RMS feed followed by update_input_rms with `n` new samples per channel. */
float wfo_meter_feed_rms(wfo_meter *m, const float *const x[2], size_t n)
{
    const int cc = m->cfg.capture_channels;
    if(n == 0)
        return m->input_rms; /* sync_rms_buffer() returned false */
    float *tmp = (float *)malloc(n * sizeof(float));
    for(size_t i = 0; i < n; ++i)
    {
        float val = 0.0f;
        for(int c = 0; c < cc; ++c)
            val = fmaxf(fabsf(x[c][i]), val);
        tmp[i] = val * val;
    }
    /* m_rms_sync_buf keeps at most m_input_rms_size pending values (src/source.cpp:1867-1870) */
    const float *src = tmp;
    size_t cnt = n;
    if(cnt > (size_t)m->RW)
    {
        src += cnt - (size_t)m->RW;
        cnt = (size_t)m->RW;
    }
    ring_push(m->rms_ring, m->RW, &m->rms_pos, src, cnt);
    free(tmp);
    float sum = 0.0f;
    for(int i = 0; i < m->RW; ++i)
        sum += m->rms_ring[i];
    m->input_rms = sqrtf(sum / (float)m->RW);
    return m->input_rms;
}

float wfo_meter_val(const wfo_meter *m, int c) { return m->val[c]; }
float wfo_meter_buf(const wfo_meter *m, int c) { return m->buf[c]; }
int wfo_meter_last_silent(const wfo_meter *m) { return m->last_silent; }

/* Batch form: tick t consumes samples [t*hop, (t+1)*hop) of each channel.
 * out_db [n_ticks][cc] = m_meter_val, out_lin [n_ticks][cc] = m_meter_buf, out_silent [n_ticks], out_rms [n_ticks]. */
void wfo_meter_run(wfo_meter *m, const float *pcm0, const float *pcm1, int n_ticks, int hop, float seconds,
                   float *out_db, float *out_lin, unsigned char *out_silent, float *out_rms)
{
    const int cc = m->cfg.capture_channels;
    for(int t = 0; t < n_ticks; ++t)
    {
        const float *x[2] = {pcm0 + (size_t)t * hop, pcm1 ? pcm1 + (size_t)t * hop : NULL};
        if(out_rms)
            out_rms[t] = wfo_meter_feed_rms(m, x, (size_t)hop);
        if(out_db || out_lin || out_silent)
        {
            wfo_meter_tick(m, x, (size_t)hop, seconds);
            for(int c = 0; c < cc; ++c)
            {
                if(out_db)
                    out_db[t * cc + c] = m->val[c];
                if(out_lin)
                    out_lin[t * cc + c] = m->buf[c];
            }
            if(out_silent)
                out_silent[t] = (unsigned char)m->last_silent;
        }
    }
}

/* ============================================================================================================= */
/* Waveform (oscilloscope) mode: WAVSourceGeneric::tick_waveform, src/source_generic.cpp:272-390, with the capture  */
/* side of src/source.cpp:1817-1888 for packets whose end is stamped "now" (get_audio_sync() == 0, reserve == 0).   */
/* Setup: src/source.cpp:1129-1143 (m_fft_size := m_width, m_waveform_samples, m_waveform_ts := 0), :1181 (DB_MIN).  */
/* ============================================================================================================= */
struct wfo_wave {
    wfo_wave_config cfg;
    size_t outsz;        /* m_fft_size = m_width */
    size_t ws;           /* m_waveform_samples */
    uint64_t clock, audio_ts, waveform_ts;
    int output_channels; /* src/source.cpp:1171 */
    float *dec[2];       /* m_decibels */
    int last_silent;
    size_t prefill;      /* zeros still pending in m_capturebufs: update() pushes m_fft_size of them, src/source.cpp:1243-1248 */
};

static uint64_t frames_to_ns(uint64_t sr, uint64_t frames) { return (uint64_t)(((__uint128_t)frames * 1000000000ull) / sr); }
static uint64_t ns_to_frames(uint64_t sr, uint64_t ns) { return (uint64_t)(((__uint128_t)ns * sr) / 1000000000ull); }

wfo_wave *wfo_wave_create(const wfo_wave_config *cfg)
{
    wfo_wave *w = (wfo_wave *)calloc(1, sizeof(*w));
    w->cfg = *cfg;
    w->outsz = (size_t)cfg->width;
    w->ws = (size_t)((double)cfg->sample_rate * ((double)cfg->meter_ms / 1000.0));
    w->clock = 10ull * 1000000000ull;
    w->output_channels = ((cfg->capture_channels > 1) || cfg->stereo) ? 2 : 1;
    w->prefill = w->outsz;
    for(int c = 0; c < 2; ++c)
    {
        w->dec[c] = (float *)malloc(w->outsz * sizeof(float));
        for(size_t i = 0; i < w->outsz; ++i)
            w->dec[c][i] = wfo_db_min();
    }
    return w;
}

void wfo_wave_destroy(wfo_wave *w)
{
    if(!w)
        return;
    free(w->dec[0]);
    free(w->dec[1]);
    free(w);
}

int wfo_wave_last_silent(const wfo_wave *w) { return w->last_silent; }
const float *wfo_wave_buffer(const wfo_wave *w, int ch) { return w->dec[ch]; }

static void rotate_left(float *a, size_t k, size_t n)
{
    if(k == 0 || k >= n)
        return;
    float *tmp = (float *)malloc(k * sizeof(float));
    memcpy(tmp, a, k * sizeof(float));
    memmove(a, a + k, (n - k) * sizeof(float));
    memcpy(a + (n - k), tmp, k * sizeof(float));
    free(tmp);
}

/* one capture packet of n samples per channel ending "now", then one tick_waveform */
void wfo_wave_tick(wfo_wave *w, const float *const x[2], size_t n, float input_rms)
{
    const uint32_t sr = w->cfg.sample_rate;
    const int cc = w->cfg.capture_channels;
    const size_t outsz = w->outsz;
    w->clock += frames_to_ns(sr, n);
    w->audio_ts = w->clock; /* timestamp + audio_len, src/source.cpp:1836 */
    /* pending samples = the start-up zeros (first tick only) ++ the new packet; the ring keeps at most m_waveform_samples
     * of them (src/source.cpp:1881-1884) and every tick consumes all (:327) */
    const size_t avail = w->prefill + n;
    const size_t total = (avail > w->ws) ? w->ws : avail;
    const size_t skip = avail - total; /* oldest samples dropped */
    float *pend[2] = {NULL, NULL};
    for(int ch = 0; ch < cc; ++ch)
    {
        pend[ch] = (float *)calloc(avail ? avail : 1, sizeof(float));
        memcpy(pend[ch] + w->prefill, x[ch], n * sizeof(float));
    }
    w->prefill = 0;
    if(total == 0)
    {
        free(pend[0]);
        free(pend[1]);
        return; /* :296-298 */
    }
    size_t counts[2] = {0, 0};
    unsigned silent_channels = 0;
    const uint64_t step_ns = ((uint64_t)w->cfg.meter_ms * 1000000ull) / (uint64_t)outsz; /* :303 */
    for(int ch = 0; ch < cc; ++ch)
    {
        const float *buf = pend[ch] + skip;
        const uint64_t start_ts = w->audio_ts - frames_to_ns(sr, total);
        const uint64_t stop_ts = w->audio_ts;
        if((start_ts >= w->audio_ts) || (stop_ts > w->audio_ts))
        {
            free(pend[0]);
            free(pend[1]);
            return; /* :321-322 */
        }
        if(w->waveform_ts < start_ts)
            w->waveform_ts = start_ts; /* :323-324 */
        if((w->waveform_ts > stop_ts) && ((w->waveform_ts - stop_ts) > step_ns))
            w->waveform_ts = start_ts; /* :325-326 */
        for(size_t i = 0; i < outsz; ++i)
        {
            const uint64_t ts = w->waveform_ts + (i * step_ns);
            if(ts >= stop_ts)
                break;
            if(ts < w->waveform_ts)
                break;
            uint64_t index = ns_to_frames(sr, w->audio_ts - ts);
            if(index < 1u)
                index = 1u;
            if(index > total)
                index = total; /* std::clamp(.., reserve_samples + 1, total_samples), :336 */
            w->dec[ch][counts[ch]++] = buf[total - index];
        }
        rotate_left(w->dec[ch], counts[ch], outsz); /* :339 */
        int silent = 1;
        for(size_t i = 0; i < outsz; ++i)
            if(w->dec[ch][i] != 0.0f)
            {
                silent = 0;
                w->last_silent = 0;
                break;
            }
        if(silent)
            if(++silent_channels >= (unsigned)cc)
                w->last_silent = 1;
    }
    free(pend[0]);
    free(pend[1]);
    w->waveform_ts += (counts[0] * step_ns); /* :358 */
    const int dch = w->cfg.stereo ? 2 : 1;
    if(w->last_silent)
    {
        for(int ch = 0; ch < dch; ++ch)
            for(size_t i = 0; i < outsz; ++i)
                w->dec[ch][i] = wfo_db_min();
        return;
    }
    if(w->output_channels > cc)
        memcpy(w->dec[1], w->dec[0], outsz * sizeof(float));
    if(w->cfg.stereo)
    {
        for(int ch = 0; ch < 2; ++ch)
            for(size_t i = outsz - counts[ch]; i < outsz; ++i)
                w->dec[ch][i] = dbfs_m(fabsf(w->dec[ch][i]));
    }
    else if(cc > 1)
    {
        for(size_t i = outsz - counts[0]; i < outsz; ++i)
            w->dec[0][i] = dbfs_m((fabsf(w->dec[0][i]) + fabsf(w->dec[1][i])) * 0.5f);
    }
    else
    {
        for(size_t i = outsz - counts[0]; i < outsz; ++i)
            w->dec[0][i] = dbfs_m(fabsf(w->dec[0][i]));
    }
    if(w->cfg.normalize_volume)
    {
        const float vc = fminf(w->cfg.volume_target - dbfs_m(input_rms), w->cfg.max_gain);
        for(int ch = 0; ch < dch; ++ch)
            for(size_t i = outsz - counts[ch]; i < outsz; ++i)
                w->dec[ch][i] += vc;
    }
}

/* out [n_ticks][display_channels][width]; tick t consumes samples [t*hop, (t+1)*hop) */
void wfo_wave_run(wfo_wave *w, const float *pcm0, const float *pcm1, int n_ticks, int hop, const float *input_rms,
                  float *out, unsigned char *out_silent)
{
    const int dch = w->cfg.stereo ? 2 : 1;
    for(int t = 0; t < n_ticks; ++t)
    {
        const float *x[2] = {pcm0 + (size_t)t * hop, pcm1 ? pcm1 + (size_t)t * hop : NULL};
        wfo_wave_tick(w, x, (size_t)hop, input_rms ? input_rms[t] : 0.0f);
        for(int c = 0; c < dch; ++c)
            memcpy(out + ((size_t)t * dch + c) * w->outsz, w->dec[c], w->outsz * sizeof(float));
        if(out_silent)
            out_silent[t] = (unsigned char)w->last_silent;
    }
}
