/* wf_oracle.c — plain-C restatement of the reference's spectrum path (see wf_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY (the checker, never the thing measured or shipped).
 * Compile with -O2 -ffp-contract=off (no FMA contraction: the reference's generic path is built
 * without -mfma, so every float expression rounds after each operation).
 *
 * All citations are relative to /root/reference.
 */
#include "wf_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI_F 3.14159265358979323846f /* std::numbers::pi_v<float> */

typedef struct { float re, im; } cf;

struct wfo_source {
    wfo_config cfg;
    int N, B;               /* fft size, bins = N/2 (src/source_avx2.cpp:29) */
    int output_channels;    /* m_output_channels, src/source.cpp:1170 */
    int num_bars;           /* src/source.cpp:1267-1271 */
    int last_silent;        /* m_last_silent */
    float window_sum;       /* m_window_sum */
    float *window;          /* m_window_coefficients[N] or NULL */
    float *slope;           /* m_slope_modifiers[B] or NULL */
    float *rolloff;         /* m_rolloff_modifiers[B] or NULL */
    float *tsmooth[2];      /* m_tsmooth_buf */
    float *decibels[2];     /* m_decibels */
    float *fft_in;          /* m_fft_input[N] */
    cf *fft_out;            /* m_fft_output (N/2+1 used) */
    /* interpolation */
    float *interp_idx;      /* m_interp_indices */
    int n_idx;
    int *band_widths;       /* m_band_widths[num_bars] */
    float *interp_w;        /* m_interp_kernel.weights */
    int interp_radius, interp_taps;
    /* gaussian */
    float *gauss_w;
    int gauss_radius, gauss_size;
    float gauss_sum;
    /* fft plan */
    cf *tw;                 /* W_M^k, M = N/2 (or N if N odd) */
    cf *tw_post;            /* W_N^k, k < N/2 */
    cf *work;
    float *tmp_a, *tmp_b;
};

/* ------------------------------------------------------------------------------------------- */
/* math helpers: src/math_funcs.hpp                                                             */
/* ------------------------------------------------------------------------------------------- */

/* log_interp, src/math_funcs.hpp:25-29 */
static float log_interp_f(float a, float b, float t) { return a * powf(b / a, t); }

/* lerp -> std::lerp(float,float,float), src/math_funcs.hpp:31-35.  libstdc++'s documented
 * algorithm (P0811R3): exact at the endpoints, monotonic. */
static float lerp_f(float a, float b, float t)
{
    if((a <= 0 && b >= 0) || (a >= 0 && b <= 0))
        return t * b + (1 - t) * a;
    if(t == 1)
        return b;
    const float x = a + t * (b - a);
    return ((t > 1) == (b > a)) ? (b < x ? x : b) : (b > x ? x : b);
}

static float clamp_f(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); } /* std::clamp */

/* sinc / lanczos, src/math_funcs.hpp:37-52 */
static float sinc_f(float x)
{
    if(x == 0.0)
        return 1.0f;
    const float tmp = PI_F * x;
    return sinf(tmp) / tmp;
}
static float lanczos_f(float x, float w)
{
    if(fabsf(x) < w)
        return sinc_f(x) * sinc_f(x / w);
    return 0.0f;
}

/* DB_MIN, src/source.cpp:43 */
float wfo_db_min(void) { return 20.0f * log10f(FLT_MIN); }

/* dbfs, src/source.hpp:293-299 */
static float dbfs(float mag) { return (mag > 0.0f) ? 20.0f * log10f(mag) : wfo_db_min(); }

/* get_gravity, src/source.hpp:301-312 */
float wfo_gravity(const wfo_source *s, float seconds)
{
    const float denom = 0.03868924705242879469662125316986f;
    const float hi = denom * 5.0f;
    const float lo = 0.0f;
    if((s->cfg.tsmoothing == WFO_TSMOOTH_NONE) || (s->cfg.gravity <= 0.0f))
        return 0.0f;
    return (s->cfg.tsmoothing == WFO_TSMOOTH_TVEXPONENTIAL) ? expf(-seconds / lerp_f(lo, hi, s->cfg.gravity))
                                                            : s->cfg.gravity;
}

/* ------------------------------------------------------------------------------------------- */
/* FFT: forward unnormalised DFT Y[k] = sum_j x[j] e^{-2 pi i jk/N}                             */
/* (deps/fftw-3.3.11/doc/reference.texi:1926-1936).  FFTW 3.3.11 is vendored in the reference   */
/* but its planner-chosen codelet sequence cannot be restated bit-for-bit; what is restated is  */
/* its structure for even N — complex DFT of size N/2 on (even + i*odd) followed by the hc2c    */
/* twiddle pass (rdft/ct-hc2c.c:59-70) — in float arithmetic with twiddles rounded from double  */
/* (kernel/trig.c:57-80).  Agreement with the real FFTW is checked in tests to ~1e-6 normwise.  */
/* ------------------------------------------------------------------------------------------- */

static int smallest_factor(int n)
{
    if((n & 1) == 0)
        return 2;
    for(int p = 3; p * p <= n; p += 2)
        if(n % p == 0)
            return p;
    return n;
}

static cf cmul(cf a, cf b)
{
    cf r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}

/* recursive mixed-radix decimation-in-time; tw = W_root^k table, tws = root/n */
static void fft_rec(const cf *in, cf *out, int n, int stride, const cf *tw, int tws)
{
    if(n == 1)
    {
        out[0] = in[0];
        return;
    }
    const int p = smallest_factor(n);
    const int m = n / p;
    for(int q = 0; q < p; ++q)
        fft_rec(in + (size_t)q * stride, out + (size_t)q * m, m, stride * p, tw, tws * p);
    cf t[64];
    cf *tp = (p <= 64) ? t : (cf *)malloc(sizeof(cf) * (size_t)p);
    for(int k = 0; k < m; ++k)
    {
        for(int q = 0; q < p; ++q)
            tp[q] = (q == 0) ? out[k] : cmul(out[(size_t)q * m + k], tw[(size_t)q * k * tws]);
        for(int r = 0; r < p; ++r)
        {
            cf acc = tp[0];
            for(int q = 1; q < p; ++q)
            {
                /* W_p^{q r} = W_n^{q r m} */
                const int e = (int)(((long long)q * r % p) * m);
                cf w = tw[(size_t)e * tws];
                cf v = cmul(tp[q], w);
                acc.re += v.re;
                acc.im += v.im;
            }
            out[(size_t)r * m + k] = acc;
        }
    }
    if(tp != t)
        free(tp);
}

static cf *make_twiddles(int n, int count)
{
    cf *tw = (cf *)malloc(sizeof(cf) * (size_t)(count > 0 ? count : 1));
    for(int k = 0; k < count; ++k)
    {
        const double a = -2.0 * M_PI * (double)k / (double)n;
        tw[k].re = (float)cos(a);
        tw[k].im = (float)sin(a);
    }
    return tw;
}

/* r2c with caller-provided plan pieces (tw: W_M^k k<M, tw_post: W_N^k k<=N/2... k<M) */
static void r2c_planned(const float *in, int n, cf *out, const cf *tw, const cf *tw_post, cf *work)
{
    if(n & 1)
    {
        cf *z = work;
        for(int i = 0; i < n; ++i)
        {
            z[i].re = in[i];
            z[i].im = 0.0f;
        }
        cf *full = work + n;
        fft_rec(z, full, n, 1, tw, 1);
        for(int k = 0; k <= n / 2; ++k)
            out[k] = full[k];
        return;
    }
    const int m = n / 2;
    /* z[j] = x[2j] + i x[2j+1]: the input viewed as m complex numbers */
    const cf *z = (const cf *)in;
    cf *Z = work;
    fft_rec(z, Z, m, 1, tw, 1);
    /* split (hc2c) pass: X[k] = E[k] + W_N^k O[k],
       E[k] = (Z[k] + conj Z[m-k])/2, O[k] = (Z[k] - conj Z[m-k])/(2i) */
    for(int k = 0; k <= m; ++k)
    {
        const cf a = Z[k % m];
        cf b = Z[(m - k) % m];
        b.im = -b.im;
        cf e, o;
        e.re = 0.5f * (a.re + b.re);
        e.im = 0.5f * (a.im + b.im);
        /* (a-b)/(2i) = -i (a-b)/2 */
        o.re = 0.5f * (a.im - b.im);
        o.im = -0.5f * (a.re - b.re);
        cf w;
        if(k < m)
            w = tw_post[k];
        else
        {
            w.re = -1.0f;
            w.im = 0.0f;
        }
        const cf wo = cmul(w, o);
        out[k].re = e.re + wo.re;
        out[k].im = e.im + wo.im;
    }
}

void wfo_r2c(const float *in, int n, float *out_interleaved)
{
    const int m = (n & 1) ? n : n / 2;
    cf *tw = make_twiddles(m, m);
    cf *twp = make_twiddles(n, n / 2 + 1);
    cf *work = (cf *)malloc(sizeof(cf) * (size_t)(2 * n + 2));
    r2c_planned(in, n, (cf *)out_interleaved, tw, twp, work);
    free(tw);
    free(twp);
    free(work);
}

/* ------------------------------------------------------------------------------------------- */
/* settings + tables (≙ WAVSource::update)                                                      */
/* ------------------------------------------------------------------------------------------- */

void wfo_config_defaults(wfo_config *c)
{
    /* src/source.cpp:119-174 */
    memset(c, 0, sizeof(*c));
    c->sample_rate = 48000;
    c->capture_channels = 2;
    c->fft_size = 4096;
    c->window = WFO_WINDOW_HANN;
    c->sine_exponent = 2;
    c->tsmoothing = WFO_TSMOOTH_EXPONENTIAL;
    c->gravity = 0.65f;
    c->fast_peaks = 0;
    c->slope = 0.0f;
    c->rolloff_q = 0.0f;
    c->rolloff_rate = 0.0f;
    c->cutoff_low = 30;
    c->cutoff_high = 17500;
    c->floor_db = -65;
    c->ceiling_db = 0;
    c->stereo = 0;
    c->normalize_volume = 0;
    c->volume_target = -8.0f;
    c->max_gain = 30.0f;
    c->silence_gate = 1;
    c->display_mode = WFO_DISPLAY_CURVE;
    c->width = 800;
    c->bar_width = 24;
    c->bar_gap = 6;
    c->log_scale = 1;
    c->mirror_freq_axis = 0;
    c->interp_mode = WFO_INTERP_CATROM;
    c->filter_mode = WFO_FILTER_NONE;
    c->filter_radius = 1.5f;
    c->height = 225;
    c->channel_spacing = 0;
    c->rounded_caps = 0;
    c->min_bar_height = 0;
}

/* make_gauss_kernel, src/filter.hpp:40-65 */
static void make_gauss(wfo_source *s, float sigma)
{
    sigma = fmaxf(fabsf(sigma), 0.01f);
    const int w = (int)ceilf(3.0f * sigma);
    const int size = (2 * w) - 1;
    s->gauss_w = (float *)malloc(sizeof(float) * (size_t)size);
    s->gauss_radius = w;
    s->gauss_size = size;
    s->gauss_sum = 0.0f;
    const float pi2 = PI_F * 2.0f;
    const float sigsqr = sigma * sigma;
    const float expdenom = 2.0f * sigsqr;
    const float coeff = (1.0f / (sqrtf(pi2) * sigma));
    int j = 0;
    for(int i = -w + 1; i < w; ++i)
    {
        const float exponent = -((float)(i * i) / expdenom);
        const float weight = coeff * expf(exponent);
        s->gauss_w[j++] = weight;
        s->gauss_sum += weight;
    }
}

/* make_catrom_kernel, src/filter.hpp:67-103 */
static void make_catrom(wfo_source *s, float t)
{
    const float matrix[4][4] = {{0, -t, 2 * t, -t}, {1, 0, t - 3, 2 - t}, {0, t, 3 - (2 * t), t - 2}, {0, 0, -t, t}};
    const int size = s->n_idx;
    s->interp_radius = 2;
    s->interp_taps = 4;
    s->interp_w = (float *)malloc(sizeof(float) * (size_t)(size > 0 ? size * 4 : 1));
    for(int i = 0; i < size; ++i)
    {
        const float u = s->interp_idx[i] - floorf(s->interp_idx[i]);
        const float row[4] = {1, u, u * u, u * u * u};
        for(int j = 0; j < 4; ++j)
        {
            float sum = 0;
            for(int k = 0; k < 4; ++k)
                sum += row[k] * matrix[j][k];
            s->interp_w[(i * 4) + j] = sum;
        }
    }
}

/* make_lanczos_kernel, src/filter.hpp:106-131 */
static void make_lanczos(wfo_source *s, int radius)
{
    const int size = s->n_idx;
    s->interp_radius = radius;
    s->interp_taps = radius * 2;
    s->interp_w = (float *)malloc(sizeof(float) * (size_t)(size > 0 ? size * radius * 2 : 1));
    const float fradius = (float)radius;
    for(int i = 0; i < size; ++i)
    {
        const float x = s->interp_idx[i];
        const long ix = (long)x;
        const long start = ix - radius + 1;
        const long stop = ix + radius;
        const long base = (long)i * radius * 2;
        for(long j = start; j <= stop; ++j)
            s->interp_w[base + (j - start)] = lanczos_f(x - (float)j, fradius);
    }
}

/* init_interp, src/source.cpp:837-896 */
static void init_interp(wfo_source *s, unsigned sz)
{
    const wfo_config *c = &s->cfg;
    const size_t fft_size = (size_t)s->N;
    const size_t maxbin = (fft_size / 2) - 1;
    const float sr = (float)c->sample_rate;
    const float lowbin = clamp_f((float)c->cutoff_low * fft_size / sr, 1.0f, (float)maxbin);
    const float highbin = clamp_f((float)c->cutoff_high * fft_size / sr, 1.0f, (float)maxbin);

    s->interp_idx = (float *)malloc(sizeof(float) * (size_t)(sz > 0 ? sz : 1));
    s->n_idx = (int)sz;
    for(unsigned i = 0; i < sz; ++i)
    {
        const float t = (c->mirror_freq_axis ? i * 2.0f : (float)i) / (float)(sz - 1);
        const float v = c->log_scale ? log_interp_f(lowbin, highbin, t) : lerp_f(lowbin, highbin, t);
        s->interp_idx[i] = clamp_f(v, lowbin, highbin);
    }

    const int bars = (c->display_mode == WFO_DISPLAY_BAR);
    if(bars)
    {
        s->band_widths = (int *)malloc(sizeof(int) * (size_t)(s->num_bars > 0 ? s->num_bars : 1));
        for(int i = 0; i < s->num_bars; ++i)
        {
            const int w = (int)(s->interp_idx[i + 1] - s->interp_idx[i]);
            s->band_widths[i] = (w > 1) ? w : 1;
        }
    }

    if(c->interp_mode != WFO_INTERP_POINT)
    {
        if(bars)
        {
            /* expand band starts to every sample point of every band, :876-889 */
            size_t total = 0;
            for(int i = 0; i < s->num_bars; ++i)
                total += (size_t)s->band_widths[i];
            float *samples = (float *)malloc(sizeof(float) * (total > 0 ? total : 1));
            size_t k = 0;
            for(int i = 0; i < s->num_bars; ++i)
                for(int j = 0; j < s->band_widths[i]; ++j)
                    samples[k++] = s->interp_idx[i] + j;
            free(s->interp_idx);
            s->interp_idx = samples;
            s->n_idx = (int)total;
        }
        if(c->interp_mode == WFO_INTERP_LANCZOS)
            make_lanczos(s, 4);
        else
            make_catrom(s, 0.5f);
    }
}

/* init_rolloff, src/source.cpp:898-918 */
static void init_rolloff(wfo_source *s)
{
    const wfo_config *c = &s->cfg;
    const size_t sz = (size_t)s->B;
    const float sr = (float)c->sample_rate;
    const float coeff = sr / (float)(size_t)s->N;
    const float ratio = exp2f(c->rolloff_q);
    const float freq_low = (float)c->cutoff_low * ratio;
    const float freq_high = (float)c->cutoff_high / ratio;
    s->rolloff = (float *)malloc(sizeof(float) * sz);
    s->rolloff[0] = 0.0f;
    for(size_t i = 1u; i < sz; ++i)
    {
        const float freq = i * coeff;
        const float ratio_low = freq_low / freq;
        const float ratio_high = freq / freq_high;
        const float low_att = (ratio_low > 1.0f) ? (c->rolloff_rate * log2f(ratio_low)) : 0.0f;
        const float high_att = (ratio_high > 1.0f) ? (c->rolloff_rate * log2f(ratio_high)) : 0.0f;
        s->rolloff[i] = low_att + high_att;
    }
}

wfo_source *wfo_create(const wfo_config *cfg_in)
{
    wfo_source *s = (wfo_source *)calloc(1, sizeof(*s));
    s->cfg = *cfg_in;
    wfo_config *c = &s->cfg;

    /* clamps of get_settings, src/source.cpp:562-577 */
    if(c->fft_size < 128)
        c->fft_size = 128;
    else if(c->fft_size & 15)
        c->fft_size &= -16;
    if((c->cutoff_high - c->cutoff_low) < 0)
    {
        c->cutoff_high = 17500;
        c->cutoff_low = 120;
    }
    if((c->ceiling_db - c->floor_db) < 1)
    {
        c->ceiling_db = 0;
        c->floor_db = -120;
    }
    if(c->height < 1)
        c->height = 225;
    if(!c->stereo || ((int)c->height - c->channel_spacing) < 1) /* src/source.cpp:579-580 */
        c->channel_spacing = 0;
    if(c->display_mode != WFO_DISPLAY_BAR) /* src/source.cpp:655-656 */
        c->rounded_caps = 0;
    if(c->capture_channels < 1)
        c->capture_channels = 1;
    if(c->capture_channels > 2)
        c->capture_channels = 2; /* src/source.cpp:1089 */

    s->N = c->fft_size;
    s->B = s->N / 2;
    const size_t N = (size_t)s->N;

    /* buffers, src/source.cpp:1169-1188 */
    s->output_channels = ((c->capture_channels > 1) || c->stereo) ? 2 : 1;
    for(int i = 0; i < s->output_channels; ++i)
    {
        s->decibels[i] = (float *)malloc(sizeof(float) * (size_t)s->B);
        if(c->tsmoothing != WFO_TSMOOTH_NONE)
            s->tsmooth[i] = (float *)calloc((size_t)s->B, sizeof(float));
        for(int k = 0; k < s->B; ++k)
            s->decibels[i][k] = wfo_db_min();
    }
    s->fft_in = (float *)calloc(N, sizeof(float));
    s->fft_out = (cf *)calloc(N, sizeof(cf));
    {
        const int m = (s->N & 1) ? s->N : s->N / 2;
        s->tw = make_twiddles(m, m);
        s->tw_post = make_twiddles(s->N, s->N / 2 + 1);
        s->work = (cf *)malloc(sizeof(cf) * (2 * N + 2));
    }

    /* window, src/source.cpp:1190-1234 */
    if(c->window != WFO_WINDOW_NONE)
    {
        s->window = (float *)malloc(sizeof(float) * N);
        const size_t Nm1 = N - 1;
        const float pi = PI_F;
        const float pi2 = 2 * pi, pi4 = 4 * pi, pi6 = 6 * pi;
        float *w = s->window;
        switch(c->window)
        {
        case WFO_WINDOW_HAMMING:
            for(size_t i = 0; i < N; ++i)
                w[i] = 0.53836f - (0.46164f * cosf((pi2 * i) / Nm1));
            break;
        case WFO_WINDOW_BLACKMAN:
            for(size_t i = 0; i < N; ++i)
                w[i] = 0.42f - (0.5f * cosf((pi2 * i) / Nm1)) + (0.08f * cosf((pi4 * i) / Nm1));
            break;
        case WFO_WINDOW_BLACKMAN_HARRIS:
            for(size_t i = 0; i < N; ++i)
                w[i] = 0.35875f - (0.48829f * cosf((pi2 * i) / Nm1)) + (0.14128f * cosf((pi4 * i) / Nm1)) -
                       (0.01168f * cosf((pi6 * i) / Nm1));
            break;
        case WFO_WINDOW_POWER_OF_SINE:
            for(size_t i = 0; i < N; ++i)
                w[i] = powf(sinf((pi * i) / Nm1), (float)c->sine_exponent);
            break;
        case WFO_WINDOW_HANN:
        default:
            for(size_t i = 0; i < N; ++i)
                w[i] = 0.5f * (1 - cosf((pi2 * i) / Nm1));
            break;
        }
        float sum = 0.0f;
        for(size_t i = 0; i < N; ++i)
            sum += w[i];
        s->window_sum = sum;
    }
    else
        s->window_sum = (float)N;

    s->last_silent = 0;

    /* interpolation tables, src/source.cpp:1250-1276 */
    if(c->display_mode == WFO_DISPLAY_CURVE)
        init_interp(s, (unsigned)c->width);
    else
    {
        const int bar_stride = c->bar_width + c->bar_gap;
        s->num_bars = (int)((unsigned)c->width / (unsigned)bar_stride);
        if(((int)c->width - (s->num_bars * bar_stride)) >= c->bar_width)
            ++s->num_bars;
        init_interp(s, (unsigned)(s->num_bars + 1));
    }

    /* gaussian, :1278-1280 */
    if(c->filter_mode == WFO_FILTER_GAUSS)
        make_gauss(s, c->filter_radius);

    /* slope, :1282-1290 */
    if(c->slope > 0.0f)
    {
        const size_t num_mods = (size_t)s->B;
        const float maxmod = (float)(num_mods - 1);
        s->slope = (float *)malloc(sizeof(float) * num_mods);
        for(size_t i = 0; i < num_mods; ++i)
            s->slope[i] = log10f(log_interp_f(10.0f, 10000.0f, ((float)i * c->slope) / maxmod));
    }

    /* roll-off, :1315-1317 */
    if((c->rolloff_q > 0.0f) && (c->rolloff_rate > 0.0f))
        init_rolloff(s);

    const int npts = wfo_num_points(s);
    s->tmp_a = (float *)malloc(sizeof(float) * (size_t)(npts > 0 ? npts : 1));
    s->tmp_b = (float *)malloc(sizeof(float) * (size_t)(npts > 0 ? npts : 1));
    return s;
}

void wfo_destroy(wfo_source *s)
{
    if(!s)
        return;
    free(s->window);
    free(s->slope);
    free(s->rolloff);
    for(int i = 0; i < 2; ++i)
    {
        free(s->tsmooth[i]);
        free(s->decibels[i]);
    }
    free(s->fft_in);
    free(s->fft_out);
    free(s->interp_idx);
    free(s->band_widths);
    free(s->interp_w);
    free(s->gauss_w);
    free(s->tw);
    free(s->tw_post);
    free(s->work);
    free(s->tmp_a);
    free(s->tmp_b);
    free(s);
}

int wfo_bins(const wfo_source *s) { return s->B; }
int wfo_display_channels(const wfo_source *s) { return s->cfg.stereo ? 2 : 1; }
int wfo_num_points(const wfo_source *s) { return (s->cfg.display_mode == WFO_DISPLAY_CURVE) ? s->cfg.width : s->num_bars; }
int wfo_last_silent(const wfo_source *s) { return s->last_silent; }
float wfo_window_sum(const wfo_source *s) { return s->window_sum; }
const float *wfo_decibels(const wfo_source *s, int ch) { return s->decibels[ch]; }
const float *wfo_tsmooth(const wfo_source *s, int ch) { return s->tsmooth[ch]; }

void wfo_set_state(wfo_source *s, int ch, const float *tsmooth, const float *decibels)
{
    if(tsmooth && s->tsmooth[ch])
        memcpy(s->tsmooth[ch], tsmooth, sizeof(float) * (size_t)s->B);
    if(decibels && s->decibels[ch])
        memcpy(s->decibels[ch], decibels, sizeof(float) * (size_t)s->B);
}

static int copy_f(const float *src, int n, float *out)
{
    if(!src)
        return 0;
    if(out)
        memcpy(out, src, sizeof(float) * (size_t)n);
    return n;
}
int wfo_get_window(const wfo_source *s, float *out) { return copy_f(s->window, s->N, out); }
int wfo_get_slope(const wfo_source *s, float *out) { return copy_f(s->slope, s->B, out); }
int wfo_get_rolloff(const wfo_source *s, float *out) { return copy_f(s->rolloff, s->B, out); }
int wfo_get_interp_indices(const wfo_source *s, float *out) { return copy_f(s->interp_idx, s->n_idx, out); }
int wfo_get_band_widths(const wfo_source *s, int32_t *out)
{
    if(!s->band_widths)
        return 0;
    if(out)
        memcpy(out, s->band_widths, sizeof(int) * (size_t)s->num_bars);
    return s->num_bars;
}
int wfo_get_interp_weights(const wfo_source *s, float *out, int *taps)
{
    if(taps)
        *taps = s->interp_taps;
    return copy_f(s->interp_w, s->n_idx * s->interp_taps, out);
}
int wfo_get_gauss_kernel(const wfo_source *s, float *out, int *radius, float *sum)
{
    if(radius)
        *radius = s->gauss_radius;
    if(sum)
        *sum = s->gauss_sum;
    return copy_f(s->gauss_w, s->gauss_size, out);
}

/* timeout / hidden branch, src/source_generic.cpp:36-48 */
void wfo_reset(wfo_source *s)
{
    if(s->last_silent)
        return;
    for(int ch = 0; ch < s->cfg.capture_channels; ++ch)
        if(s->tsmooth[ch])
            memset(s->tsmooth[ch], 0, sizeof(float) * (size_t)s->B);
    for(int ch = 0; ch < (s->cfg.stereo ? 2 : 1); ++ch)
        for(int i = 0; i < s->B; ++i)
            s->decibels[ch][i] = wfo_db_min();
    s->last_silent = 1;
}

/* ------------------------------------------------------------------------------------------- */
/* tick_spectrum, src/source_generic.cpp:26-180 (from the frame fetch onwards)                  */
/* ------------------------------------------------------------------------------------------- */
void wfo_tick(wfo_source *s, const float *const frames[2], float seconds, float input_rms)
{
    const wfo_config *c = &s->cfg;
    const size_t N = (size_t)s->N;
    const size_t outsz = (size_t)s->B;
    const float DB_MIN = wfo_db_min();
    unsigned silent_channels = 0u;

    for(unsigned channel = 0u; channel < (unsigned)c->capture_channels; ++channel)
    {
        /* :55-61 frame fetch ("not enough audio" -> continue) */
        if(frames[channel] != NULL)
            memcpy(s->fft_in, frames[channel], N * sizeof(float));
        else
            continue;

        /* :63-72 silence scan */
        int silent = 1;
        for(size_t i = 0; i < N; ++i)
        {
            if(s->fft_in[i] != 0.0f)
            {
                silent = 0;
                s->last_silent = 0;
                break;
            }
        }

        /* :74-95 "wait for gravity" */
        if(silent && c->silence_gate)
        {
            if(s->last_silent)
                continue;
            int outsilent = 1;
            const float floor = (float)(c->floor_db - 10);
            for(size_t i = 0; i < outsz; ++i)
            {
                const unsigned ch = (c->stereo) ? channel : 0u;
                if(s->decibels[ch][i] > floor)
                {
                    outsilent = 0;
                    break;
                }
            }
            if(outsilent)
            {
                if(++silent_channels >= (unsigned)c->capture_channels)
                    s->last_silent = 1;
                continue;
            }
        }

        /* :97-103 window */
        if(c->window != WFO_WINDOW_NONE)
            for(size_t i = 0; i < N; ++i)
                s->fft_in[i] *= s->window[i];

        /* :105-108 FFT */
        r2c_planned(s->fft_in, s->N, s->fft_out, s->tw, s->tw_post, s->work);

        /* :110-135 magnitude, slope, EMA */
        const float mag_coefficient = 2.0f / s->window_sum;
        const float g = wfo_gravity(s, seconds);
        const float g2 = 1.0f - g;
        const int slope = c->slope > 0.0f;
        for(size_t i = 0; i < outsz; ++i)
        {
            const float real = s->fft_out[i].re;
            const float imag = s->fft_out[i].im;
            float mag = hypotf(real, imag) * mag_coefficient;
            if(slope)
                mag *= s->slope[i];
            if(c->tsmoothing != WFO_TSMOOTH_NONE)
            {
                float oldval = s->tsmooth[channel][i];
                if(c->fast_peaks)
                    oldval = (mag > oldval) ? mag : oldval; /* std::max(mag, oldval) */
                mag = (g * oldval) + (g2 * mag);
                s->tsmooth[channel][i] = mag;
            }
            s->decibels[channel][i] = mag;
        }
    }

    /* :138-139 */
    if(s->last_silent)
        return;

    /* :141-142 */
    if(s->output_channels > c->capture_channels)
        memcpy(s->decibels[1], s->decibels[0], outsz * sizeof(float));

    /* :144-159 dBFS */
    if(c->stereo)
    {
        for(int channel = 0; channel < 2; ++channel)
            for(size_t i = 0; i < outsz; ++i)
                s->decibels[channel][i] = dbfs(s->decibels[channel][i]);
    }
    else if(c->capture_channels > 1)
    {
        for(size_t i = 0; i < outsz; ++i)
            s->decibels[0][i] = dbfs((s->decibels[0][i] + s->decibels[1][i]) * 0.5f);
    }
    else
    {
        for(size_t i = 0; i < outsz; ++i)
            s->decibels[0][i] = dbfs(s->decibels[0][i]);
    }

    /* :161-167 volume normalisation (bins >= 1) */
    if(c->normalize_volume)
    {
        const float comp = c->volume_target - dbfs(input_rms);
        const float volume_compensation = (comp < c->max_gain) ? comp : c->max_gain; /* std::min */
        for(int channel = 0; channel < (c->stereo ? 2 : 1); ++channel)
            for(size_t i = 1; i < outsz; ++i)
                s->decibels[channel][i] += volume_compensation;
    }

    /* :169-179 roll-off (bins >= 1) */
    if((c->rolloff_q > 0.0f) && (c->rolloff_rate > 0.0f))
    {
        for(int channel = 0; channel < (c->stereo ? 2 : 1); ++channel)
            for(size_t i = 1; i < outsz; ++i)
            {
                const float val = s->decibels[channel][i] - s->rolloff[i];
                s->decibels[channel][i] = (val > DB_MIN) ? val : DB_MIN; /* std::max(val, DB_MIN) */
            }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* interpolation + gaussian, src/filter.hpp:133-211                                             */
/* ------------------------------------------------------------------------------------------- */

/* kernel_convolve, src/filter.hpp:160-169 */
static float kernel_convolve(const float *samples, size_t sz, const float *weights, int radius, long index,
                             long kernel_base)
{
    const long start = (index - radius) + 1;
    long stop = index + radius + 1;
    if(stop > (long)sz)
        stop = (long)sz;
    float sum = 0.0f;
    for(long i = (start > 0 ? start : 0); i < stop; ++i)
        sum += samples[i] * weights[kernel_base + (i - start)];
    return sum;
}

/* weighted_avg, src/filter.hpp:133-158 */
static float weighted_avg(const float *samples, long n, const wfo_source *s, long index)
{
    const long start = (index - s->gauss_radius) + 1;
    const long stop = index + s->gauss_radius;
    float sum = 0.0f;
    if((start < 0) || (stop > n))
    {
        const long loopstart = (start > 0) ? start : 0;
        const long loopstop = (stop < n) ? stop : n;
        float wsum = 0.0f;
        for(long i = loopstart; i < loopstop; ++i)
        {
            const float weight = s->gauss_w[i - start];
            wsum += weight;
            sum += samples[i] * weight;
        }
        return sum / wsum;
    }
    for(long i = start; i < stop; ++i)
        sum += samples[i] * s->gauss_w[i - start];
    return sum / s->gauss_sum;
}

int wfo_interp(wfo_source *s, int ch, float *out)
{
    const wfo_config *c = &s->cfg;
    const float *db = s->decibels[ch];
    const size_t sz = (size_t)s->B;
    const int npts = wfo_num_points(s);
    float *a = s->tmp_a;
    const long d = (long)s->interp_radius * 2;

    if(c->display_mode == WFO_DISPLAY_CURVE)
    {
        if(c->interp_mode != WFO_INTERP_POINT)
        {
            /* apply_interp_filter (curve), src/filter.hpp:182-192; call site src/source.cpp:1381-1390 */
            long j = 0;
            for(long i = 0; i < npts; ++i, j += d)
                a[i] = kernel_convolve(db, sz, s->interp_w, s->interp_radius, (long)s->interp_idx[i], j);
        }
        else /* src/source.cpp:1392-1394 */
            for(int i = 0; i < npts; ++i)
                a[i] = db[(int)s->interp_idx[i]];
    }
    else
    {
        if(c->interp_mode != WFO_INTERP_POINT)
        {
            /* apply_interp_filter (bars), src/filter.hpp:195-211; call site src/source.cpp:1513-1521 */
            long k = 0, l = 0;
            for(long i = 0; i < npts; ++i)
            {
                float sum = 0.0f;
                const long count = (long)s->band_widths[i];
                for(long j = 0; j < count; ++j, ++k, l += d)
                    sum += kernel_convolve(db, sz, s->interp_w, s->interp_radius, (long)s->interp_idx[k], l);
                a[i] = sum / (float)count;
            }
        }
        else
        {
            /* src/source.cpp:1523-1532 */
            for(int i = 0; i < npts; ++i)
            {
                float sum = 0.0f;
                const size_t count = (size_t)s->band_widths[i];
                for(size_t j = 0; j < count; ++j)
                    sum += db[(size_t)s->interp_idx[i] + j];
                a[i] = sum / (float)count;
            }
        }
    }

    if(c->filter_mode != WFO_FILTER_NONE)
    {
        /* apply_filter, src/filter.hpp:171-180; call sites src/source.cpp:1396-1406,1535-1545 */
        float *b = s->tmp_b;
        for(int i = 0; i < npts; ++i)
            b[i] = weighted_avg(a, npts, s, i);
        a = b;
    }
    if(out)
        memcpy(out, a, sizeof(float) * (size_t)npts);
    return npts;
}

void wfo_render_pixels(wfo_source *s, float *out, float *miny_out, float *minpos_out)
{
    const wfo_config *c = &s->cfg;
    const int npts = wfo_num_points(s);
    const int dch = wfo_display_channels(s);
    const float center = (float)c->height / 2;
    const float bottom = (float)c->height;
    const int dbrange = c->ceiling_db - c->floor_db;
    const float cpos = c->stereo ? center : bottom;
    const float channel_offset = c->channel_spacing * 0.5f;
    float lo, hi;
    if(c->display_mode == WFO_DISPLAY_CURVE)
    {
        lo = 0.0f; /* src/source.cpp:1410 */
        hi = cpos - channel_offset;
    }
    else
    {
        /* src/source.cpp:1481-1493 */
        const float cap_radius = (float)c->bar_width / 2.0f; /* :1297 */
        float border_top = c->rounded_caps ? cap_radius : 0.0f;
        float border_bottom = (c->rounded_caps && (!c->stereo || (c->channel_spacing > 0))) ? cpos - cap_radius : cpos;
        if(c->channel_spacing > 0)
            border_bottom -= channel_offset;
        if(c->min_bar_height > 0)
            border_bottom -= c->min_bar_height;
        border_bottom = clamp_f(border_bottom, border_top, cpos);
        lo = border_top;
        hi = border_bottom;
    }
    float miny = cpos;
    unsigned minpos = 0u;
    for(int channel = 0; channel < dch; ++channel)
    {
        float *buf = out + (size_t)channel * npts;
        wfo_interp(s, channel, buf);
        for(int i = 0; i < npts; ++i)
        {
            /* src/source.cpp:1408-1417, :1548-1557 */
            const float val = lerp_f(lo, hi, clamp_f(c->ceiling_db - buf[i], 0.0f, (float)dbrange) / dbrange);
            if(val < miny)
            {
                miny = val;
                minpos = (unsigned)i;
            }
            buf[i] = val;
        }
        if(c->mirror_freq_axis)
        {
            /* src/source.cpp:1419-1424, :1559-1564 */
            const unsigned half = (unsigned)npts / 2u;
            for(unsigned i = half + 1; i < (unsigned)npts; ++i)
                buf[i] = buf[half - (i - half)];
        }
    }
    if(miny_out)
        *miny_out = miny;
    if(minpos_out)
        *minpos_out = (float)minpos;
}

int wfo_run_stft_px(wfo_source *s, const float *pcm0, const float *pcm1, int n_frames, int hop, float seconds,
                    const float *input_rms, float *out_db, float *out_points, unsigned char *out_silent,
                    float *out_pixels, float *out_min)
{
    const int dch = wfo_display_channels(s);
    const size_t B = (size_t)s->B;
    const int npts = wfo_num_points(s);
    for(int t = 0; t < n_frames; ++t)
    {
        const float *frames[2];
        frames[0] = pcm0 + (size_t)t * (size_t)hop;
        frames[1] = pcm1 ? pcm1 + (size_t)t * (size_t)hop : NULL;
        wfo_tick(s, frames, seconds, input_rms ? input_rms[t] : 0.0f);
        if(out_db)
            for(int ch = 0; ch < dch; ++ch)
                memcpy(out_db + ((size_t)t * dch + ch) * B, s->decibels[ch], B * sizeof(float));
        if(out_points)
            for(int ch = 0; ch < dch; ++ch)
                wfo_interp(s, ch, out_points + ((size_t)t * dch + ch) * (size_t)npts);
        if(out_pixels)
            wfo_render_pixels(s, out_pixels + (size_t)t * dch * (size_t)npts, out_min ? out_min + 2 * t : NULL,
                              out_min ? out_min + 2 * t + 1 : NULL);
        if(out_silent)
            out_silent[t] = (unsigned char)s->last_silent;
    }
    return n_frames;
}

int wfo_run_stft(wfo_source *s, const float *pcm0, const float *pcm1, int n_frames, int hop, float seconds,
                 const float *input_rms, float *out_db, float *out_points, unsigned char *out_silent)
{
    const int dch = wfo_display_channels(s);
    const size_t B = (size_t)s->B;
    const int npts = wfo_num_points(s);
    for(int t = 0; t < n_frames; ++t)
    {
        const float *frames[2];
        frames[0] = pcm0 + (size_t)t * (size_t)hop;
        frames[1] = pcm1 ? pcm1 + (size_t)t * (size_t)hop : NULL;
        wfo_tick(s, frames, seconds, input_rms ? input_rms[t] : 0.0f);
        if(out_db)
            for(int ch = 0; ch < dch; ++ch)
                memcpy(out_db + ((size_t)t * dch + ch) * B, s->decibels[ch], B * sizeof(float));
        if(out_points)
            for(int ch = 0; ch < dch; ++ch)
                wfo_interp(s, ch, out_points + ((size_t)t * dch + ch) * (size_t)npts);
        if(out_silent)
            out_silent[t] = (unsigned char)s->last_silent;
    }
    return n_frames;
}
