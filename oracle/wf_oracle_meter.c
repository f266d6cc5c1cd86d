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
