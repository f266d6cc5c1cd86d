// Harness that drives the UNMODIFIED reference plugin classes through their own public surface
// (update / capture_output_bus / tick / render) and exposes their protected state to the tests.
//
// TEST INFRASTRUCTURE ONLY — this is how the parity oracle is pinned to the real reference and how
// bench.py's `--impl reference` arm times the reference's own CPU path.  Nothing in the product
// path (waveform_b200/, include/) may link or call this.
//
// Reference surface used (all /root/reference/src):
//   WAVSource::update(obs_data_t*)                 source.cpp:1077-1322   (tables, buffers, FFTW plan)
//   WAVSource::capture_output_bus(...)             source.cpp:1890-1893 -> capture_audio :1817-1888
//   WAVSource::tick(float)                         source.cpp:1324-1344 -> tick_spectrum (virtual)
//   WAVSource::render(gs_effect_t*)                source.cpp:1346-1358 -> render_curve/bars
//   apply_interp_filter[_fma3], apply_filter[_fma3] filter.hpp:171-211, filter_fma3.cpp:58-219
#include "waveform_config.hpp"
#include "source.hpp"
#include "settings.hpp"
#include "obs_stub_hooks.h"
#ifdef WFREF_WITH_CUDA
// libwaveform_ref_cuda.so only: the product's plugin-side binding compiled against the same unmodified reference sources
// (waveform_b200/host/source_cuda.hpp), so that the seam of INTEGRATION.md §1 is a compiled, tested artefact.
#include "source_cuda.hpp"
#endif

#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

namespace {

std::once_flag g_register_once;
std::mutex g_update_mtx; // FFTW's planner is not thread-safe (only fftwf_execute is)

struct ProbeAPI {
    virtual ~ProbeAPI() = default;
    virtual WAVSource *src() = 0;
    virtual size_t fft_size() = 0;
    virtual uint32_t capture_channels() = 0;
    virtual uint32_t output_channels() = 0;
    virtual bool stereo() = 0;
    virtual bool last_silent() = 0;
    virtual float window_sum() = 0;
    virtual const float *decibels(int ch) = 0;
    virtual const float *tsmooth(int ch) = 0;
    virtual const float *window() = 0;
    virtual const float *slope() = 0;
    virtual const float *rolloff() = 0;
    virtual const std::vector<float> &interp_indices() = 0;
    virtual const std::vector<int> &band_widths() = 0;
    virtual const Kernel<float> &interp_kernel() = 0;
    virtual const Kernel<float> &gauss_kernel() = 0;
    virtual std::vector<float> *interp_bufs() = 0;
    virtual int num_bars() = 0;
    virtual unsigned width_px() = 0;
    virtual int interp_mode() = 0;
    virtual int filter_mode() = 0;
    virtual int display_mode() = 0;
    virtual float input_rms() = 0;
    virtual void force_input_rms(bool enable, float v) = 0;
    virtual float gravity_for(float seconds) = 0;
    virtual float db_min() = 0;
    virtual float meter_val(int ch) = 0;
    virtual float meter_buf(int ch) = 0;
};

template<class Base>
class Probe final : public Base, public ProbeAPI {
    bool m_force_rms = false;
    float m_forced_rms = 0.0f;

protected:
    void update_input_rms() override
    {
        if(m_force_rms)
            this->m_input_rms = m_forced_rms;
        else
            Base::update_input_rms();
    }

public:
    using Base::Base;
    WAVSource *src() override { return this; }
    size_t fft_size() override { return this->m_fft_size; }
    uint32_t capture_channels() override { return this->m_capture_channels; }
    uint32_t output_channels() override { return this->m_output_channels; }
    bool stereo() override { return this->m_stereo; }
    bool last_silent() override { return this->m_last_silent; }
    float window_sum() override { return this->m_window_sum; }
    const float *decibels(int ch) override { return this->m_decibels[ch].get(); }
    const float *tsmooth(int ch) override { return this->m_tsmooth_buf[ch].get(); }
    const float *window() override { return this->m_window_coefficients.get(); }
    const float *slope() override { return this->m_slope_modifiers.get(); }
    const float *rolloff() override { return this->m_rolloff_modifiers.get(); }
    const std::vector<float> &interp_indices() override { return this->m_interp_indices; }
    const std::vector<int> &band_widths() override { return this->m_band_widths; }
    const Kernel<float> &interp_kernel() override { return this->m_interp_kernel; }
    const Kernel<float> &gauss_kernel() override { return this->m_kernel; }
    std::vector<float> *interp_bufs() override { return this->m_interp_bufs; }
    int num_bars() override { return this->m_num_bars; }
    unsigned width_px() override { return this->m_width; }
    int interp_mode() override { return (int)this->m_interp_mode; }
    int filter_mode() override { return (int)this->m_filter_mode; }
    int display_mode() override { return (int)this->m_display_mode; }
    float input_rms() override { return this->m_input_rms; }
    void force_input_rms(bool enable, float v) override
    {
        m_force_rms = enable;
        m_forced_rms = v;
    }
    float gravity_for(float seconds) override { return this->get_gravity(seconds); }
    float db_min() override { return WAVSource::DB_MIN; }
    float meter_val(int ch) override { return this->m_meter_val[ch]; }
    float meter_buf(int ch) override { return this->m_meter_buf[ch]; }
};

struct Ref {
    std::unique_ptr<ProbeAPI> probe;
    obs_data_t *settings = nullptr;
    uint64_t clock_ns = 10ull * 1000000000ull;
    uint32_t sample_rate = 48000;
    int channels = 2;
    uint32_t fps_num = 60, fps_den = 1;
    bool showing = true;

    void bind() // make the fake libobs reflect this instance on the calling thread
    {
        wfstub_set_clock_ns(clock_ns);
        wfstub_set_audio(sample_rate, channels);
        wfstub_set_fps(fps_num, fps_den);
        wfstub_set_showing(showing);
    }
};

} // namespace

extern "C" {

// impl: 0 = WAVSourceGeneric (parity target), 1 = WAVSourceAVX, 2 = WAVSourceAVX2,
//       3 = WAVSourceCUDA (libwaveform_ref_cuda.so only; NULL elsewhere)
void *wfref_create(int impl, uint32_t sample_rate, int channels, uint32_t fps_num, uint32_t fps_den)
{
    std::call_once(g_register_once, [] { WAVSource::register_source(); });
    auto r = new Ref();
    r->sample_rate = sample_rate;
    r->channels = channels;
    r->fps_num = fps_num;
    r->fps_den = fps_den ? fps_den : 1;
    r->bind();
    r->settings = wfstub_data_create();
    auto info = wfstub_registered_info();
    if(info && info->get_defaults)
        info->get_defaults(r->settings); // reference defaults, src/source.cpp:119-174
    // capture from the "output bus" so no obs_source is needed (src/source.cpp:685-703)
    wfstub_data_set_string(r->settings, P_AUDIO_SRC, P_OUTPUT_BUS);
    {
        std::lock_guard<std::mutex> lk(g_update_mtx);
        switch(impl)
        {
#ifdef WFREF_WITH_CUDA
        case 3: r->probe = std::make_unique<Probe<WAVSourceCUDA>>(nullptr); break;
#else
        case 3: wfstub_data_destroy(r->settings); delete r; return nullptr;
#endif
        case 2: r->probe = std::make_unique<Probe<WAVSourceAVX2>>(nullptr); break;
        case 1: r->probe = std::make_unique<Probe<WAVSourceAVX>>(nullptr); break;
        default: r->probe = std::make_unique<Probe<WAVSourceGeneric>>(nullptr); break;
        }
    }
    return r;
}

void wfref_destroy(void *h)
{
    auto r = static_cast<Ref *>(h);
    if(!r)
        return;
    r->bind();
    {
        std::lock_guard<std::mutex> lk(g_update_mtx);
        r->probe.reset();
    }
    wfstub_data_destroy(r->settings);
    delete r;
}

void wfref_set_int(void *h, const char *k, long long v) { wfstub_data_set_int(static_cast<Ref *>(h)->settings, k, v); }
void wfref_set_double(void *h, const char *k, double v) { wfstub_data_set_double(static_cast<Ref *>(h)->settings, k, v); }
void wfref_set_bool(void *h, const char *k, int v) { wfstub_data_set_bool(static_cast<Ref *>(h)->settings, k, v != 0); }
void wfref_set_string(void *h, const char *k, const char *v) { wfstub_data_set_string(static_cast<Ref *>(h)->settings, k, v); }

void wfref_update(void *h)
{
    auto r = static_cast<Ref *>(h);
    r->bind();
    std::lock_guard<std::mutex> lk(g_update_mtx);
    r->probe->src()->update(r->settings);
}

void wfref_set_showing(void *h, int showing)
{
    auto r = static_cast<Ref *>(h);
    r->showing = showing != 0;
    r->bind();
    if(showing)
        r->probe->src()->show();
    else
        r->probe->src()->hide();
}

void wfref_advance_clock_ns(void *h, uint64_t ns) { static_cast<Ref *>(h)->clock_ns += ns; }
uint64_t wfref_clock_ns(void *h) { return static_cast<Ref *>(h)->clock_ns; }

// Feed `frames` planar float samples.  The packet is stamped so that the end of the packet
// coincides with the fake "now" (=> get_audio_sync()==0 at a tick issued at the same instant,
// src/source.hpp:279-285), unless ts_adjust_ns shifts it (A/V-sync tests).
void wfref_push_audio(void *h, const float *ch0, const float *ch1, uint32_t frames, int64_t ts_adjust_ns)
{
    auto r = static_cast<Ref *>(h);
    r->bind();
    audio_data ad{};
    ad.data[0] = (uint8_t *)ch0;
    ad.data[1] = (uint8_t *)ch1;
    ad.frames = frames;
    ad.timestamp = (uint64_t)((int64_t)(r->clock_ns - audio_frames_to_ns(r->sample_rate, frames)) + ts_adjust_ns);
    r->probe->src()->capture_output_bus(0, &ad);
}

void wfref_tick(void *h, float seconds)
{
    auto r = static_cast<Ref *>(h);
    r->bind();
    r->probe->src()->tick(seconds);
}

void wfref_render(void *h)
{
    auto r = static_cast<Ref *>(h);
    r->bind();
    r->probe->src()->render(nullptr);
}

uint64_t wfref_fft_size(void *h) { return static_cast<Ref *>(h)->probe->fft_size(); }
uint32_t wfref_capture_channels(void *h) { return static_cast<Ref *>(h)->probe->capture_channels(); }
uint32_t wfref_output_channels(void *h) { return static_cast<Ref *>(h)->probe->output_channels(); }
int wfref_stereo(void *h) { return static_cast<Ref *>(h)->probe->stereo() ? 1 : 0; }
int wfref_last_silent(void *h) { return static_cast<Ref *>(h)->probe->last_silent() ? 1 : 0; }
float wfref_window_sum(void *h) { return static_cast<Ref *>(h)->probe->window_sum(); }
float wfref_db_min(void *h) { return static_cast<Ref *>(h)->probe->db_min(); }
float wfref_gravity(void *h, float seconds) { return static_cast<Ref *>(h)->probe->gravity_for(seconds); }
float wfref_input_rms(void *h) { return static_cast<Ref *>(h)->probe->input_rms(); }
void wfref_force_input_rms(void *h, int enable, float v) { static_cast<Ref *>(h)->probe->force_input_rms(enable != 0, v); }
int wfref_num_bars(void *h) { return static_cast<Ref *>(h)->probe->num_bars(); }
uint32_t wfref_width(void *h) { return static_cast<Ref *>(h)->probe->width_px(); }

static int copy_out(const float *src, size_t n, float *dst)
{
    if(!src)
        return 0;
    if(dst)
        memcpy(dst, src, n * sizeof(float));
    return (int)n;
}

int wfref_get_decibels(void *h, int ch, float *out)
{
    auto p = static_cast<Ref *>(h)->probe.get();
    return copy_out(p->decibels(ch), p->fft_size() / 2, out);
}
int wfref_get_tsmooth(void *h, int ch, float *out)
{
    auto p = static_cast<Ref *>(h)->probe.get();
    return copy_out(p->tsmooth(ch), p->fft_size() / 2, out);
}
int wfref_get_window(void *h, float *out)
{
    auto p = static_cast<Ref *>(h)->probe.get();
    return copy_out(p->window(), p->fft_size(), out);
}
int wfref_get_slope(void *h, float *out)
{
    auto p = static_cast<Ref *>(h)->probe.get();
    return copy_out(p->slope(), p->fft_size() / 2, out);
}
int wfref_get_rolloff(void *h, float *out)
{
    auto p = static_cast<Ref *>(h)->probe.get();
    return copy_out(p->rolloff(), p->fft_size() / 2, out);
}
int wfref_get_interp_indices(void *h, float *out)
{
    auto &v = static_cast<Ref *>(h)->probe->interp_indices();
    if(out && !v.empty())
        memcpy(out, v.data(), v.size() * sizeof(float));
    return (int)v.size();
}
int wfref_get_band_widths(void *h, int *out)
{
    auto &v = static_cast<Ref *>(h)->probe->band_widths();
    if(out && !v.empty())
        memcpy(out, v.data(), v.size() * sizeof(int));
    return (int)v.size();
}
// returns taps per point (kernel.size); weights has indices*size entries
int wfref_get_interp_kernel(void *h, float *out, int max_floats)
{
    auto p = static_cast<Ref *>(h)->probe.get();
    auto &k = p->interp_kernel();
    auto n = (int)p->interp_indices().size() * k.size;
    if(out && k.weights.get() && n <= max_floats)
        memcpy(out, k.weights.get(), (size_t)n * sizeof(float));
    return k.size;
}
// returns kernel.size; *radius, *sum filled
int wfref_get_gauss_kernel(void *h, float *out, int *radius, float *sum)
{
    auto &k = static_cast<Ref *>(h)->probe->gauss_kernel();
    if(out && k.weights.get())
        memcpy(out, k.weights.get(), (size_t)k.size * sizeof(float));
    if(radius)
        *radius = k.radius;
    if(sum)
        *sum = k.sum;
    return k.size;
}

// Interpolation (+ optional Gaussian smoothing) exactly as render_curve/render_bars perform it on the
// current m_decibels (src/source.cpp:1381-1406, 1510-1546), stopping BEFORE the dB->pixel mapping.
// use_fma3 selects the reference's SIMD variants (what the plugin runs on an AVX machine).
// Returns the number of display points written per channel.
int wfref_interp(void *h, int channel, int use_fma3, float *out)
{
    auto p = static_cast<Ref *>(h)->probe.get();
    const auto sz = p->fft_size() / 2;
    const float *db = p->decibels(channel);
    auto &idx = p->interp_indices();
    auto &bw = p->band_widths();
    auto &kern = p->interp_kernel();
    const int disp = p->display_mode();
    const bool curve = (disp == (int)DisplayMode::CURVE);
    std::vector<float> a, b;
    size_t npts;
    if(curve)
    {
        npts = p->width_px();
        a.resize(npts);
        if(p->interp_mode() != (int)InterpMode::POINT)
        {
            if(use_fma3)
                apply_interp_filter_fma3(db, sz, idx, kern, a);
            else
                apply_interp_filter(db, sz, idx, kern, a);
        }
        else
            for(size_t i = 0; i < npts; ++i)
                a[i] = db[(int)idx[i]];
    }
    else
    {
        npts = (size_t)p->num_bars();
        a.resize(npts);
        if(p->interp_mode() != (int)InterpMode::POINT)
        {
            if(use_fma3)
                apply_interp_filter_fma3(db, sz, bw, idx, kern, a);
            else
                apply_interp_filter(db, sz, bw, idx, kern, a);
        }
        else
        {
            for(size_t i = 0; i < npts; ++i)
            {
                float sum = 0.0f;
                auto count = (size_t)bw[i];
                for(size_t j = 0; j < count; ++j)
                    sum += db[(size_t)idx[i] + j];
                a[i] = sum / (float)count;
            }
        }
    }
    if(p->filter_mode() != (int)FilterMode::NONE)
    {
        b.resize(npts);
        if(use_fma3)
            apply_filter_fma3(a, p->gauss_kernel(), b);
        else
            apply_filter(a, p->gauss_kernel(), b);
        a.swap(b);
    }
    if(out)
        memcpy(out, a.data(), npts * sizeof(float));
    return (int)npts;
}

// After wfref_render(): the pixel-space values render_* left in m_interp_bufs[channel].
int wfref_get_render_buf(void *h, int channel, float *out)
{
    auto p = static_cast<Ref *>(h)->probe.get();
    auto &v = p->interp_bufs()[channel];
    if(out && !v.empty())
        memcpy(out, v.data(), v.size() * sizeof(float));
    return (int)v.size();
}

// Sliding STFT through the reference's own capture/tick loop:
//   frame t = pcm[t*hop : t*hop + N]  (first push is N samples, then `hop` per tick; each tick takes
//   the latest N samples because the packet end is stamped "now", src/source_generic.cpp:50-59).
// out_db  : [n_frames][display_channels][N/2] or NULL (timing only);  display_channels = stereo?2:1
// out_pts : [n_frames][display_channels][points] interpolated display points or NULL
// rms     : optional per-frame forced m_input_rms (volume normalisation input), or NULL
// Returns frames processed.
int wfref_run_stft(void *h, const float *pcm0, const float *pcm1, int64_t n_samples, int n_frames, int hop,
                   float seconds, const float *rms, float *out_db, float *out_pts, int use_fma3,
                   unsigned char *out_silent)
{
    auto r = static_cast<Ref *>(h);
    auto p = r->probe.get();
    const auto N = (int64_t)p->fft_size();
    const auto B = (size_t)(N / 2);
    const int dch = p->stereo() ? 2 : 1;
    int64_t pos = 0;
    int done = 0;
    for(int t = 0; t < n_frames; ++t)
    {
        const int64_t want = (t == 0) ? N : hop;
        if(pos + want > n_samples)
            break;
        r->clock_ns += audio_frames_to_ns(r->sample_rate, (uint64_t)want);
        wfref_push_audio(h, pcm0 + pos, pcm1 ? pcm1 + pos : nullptr, (uint32_t)want, 0);
        pos += want;
        if(rms)
            p->force_input_rms(true, rms[t]);
        wfref_tick(h, seconds);
        if(out_db)
            for(int c = 0; c < dch; ++c)
                memcpy(out_db + ((size_t)t * dch + c) * B, p->decibels(c), B * sizeof(float));
        if(out_pts)
        {
            int npts = wfref_interp(h, 0, use_fma3, nullptr);
            for(int c = 0; c < dch; ++c)
                wfref_interp(h, c, use_fma3, out_pts + ((size_t)t * dch + c) * (size_t)npts);
        }
        if(out_silent)
            out_silent[t] = p->last_silent() ? 1 : 0;
        ++done;
    }
    return done;
}

// Level meter / RMS feed through the reference's own capture/tick loop: tick t is preceded by a push of samples
// [t*hop, (t+1)*hop) stamped "now" (get_audio_sync()==0, so tick_meter consumes everything captured).
//   meter mode (display_mode = level_meter / stepped_meter): out_db [n_ticks][capture_channels] = m_meter_val,
//   out_lin = m_meter_buf, out_silent = m_last_silent            (src/source_generic.cpp:182-270)
//   spectrum mode with normalize_volume: out_rms [n_ticks] = m_input_rms after the tick's update_input_rms()
//                                                                 (src/source.cpp:1330-1331, 1842-1871)
int wfref_run_meter(void *h, const float *pcm0, const float *pcm1, int n_ticks, int hop, float seconds, float *out_db,
                    float *out_lin, unsigned char *out_silent, float *out_rms)
{
    auto r = static_cast<Ref *>(h);
    auto p = r->probe.get();
    const int cc = (int)p->capture_channels();
    for(int t = 0; t < n_ticks; ++t)
    {
        r->clock_ns += audio_frames_to_ns(r->sample_rate, (uint64_t)hop);
        wfref_push_audio(h, pcm0 + (size_t)t * hop, pcm1 ? pcm1 + (size_t)t * hop : nullptr, (uint32_t)hop, 0);
        wfref_tick(h, seconds);
        for(int c = 0; c < cc; ++c)
        {
            if(out_db)
                out_db[t * cc + c] = p->meter_val(c);
            if(out_lin)
                out_lin[t * cc + c] = p->meter_buf(c);
        }
        if(out_silent)
            out_silent[t] = p->last_silent() ? 1 : 0;
        if(out_rms)
            out_rms[t] = p->input_rms();
    }
    return n_ticks;
}

// Waveform (oscilloscope) mode through the reference's own capture/tick loop (display_mode = "waveform"): tick t is preceded
// by a push of samples [t*hop, (t+1)*hop) stamped "now".  out [n_ticks][display_channels][m_width] = m_decibels after the
// tick (src/source_generic.cpp:272-390); rms: optional forced m_input_rms per tick (volume normalisation).
int wfref_run_wave(void *h, const float *pcm0, const float *pcm1, int n_ticks, int hop, float seconds, const float *rms,
                   float *out, unsigned char *out_silent)
{
    auto r = static_cast<Ref *>(h);
    auto p = r->probe.get();
    const size_t outsz = p->fft_size();
    const int dch = p->stereo() ? 2 : 1;
    for(int t = 0; t < n_ticks; ++t)
    {
        r->clock_ns += audio_frames_to_ns(r->sample_rate, (uint64_t)hop);
        wfref_push_audio(h, pcm0 + (size_t)t * hop, pcm1 ? pcm1 + (size_t)t * hop : nullptr, (uint32_t)hop, 0);
        if(rms)
            p->force_input_rms(true, rms[t]);
        wfref_tick(h, seconds);
        for(int c = 0; c < dch; ++c)
            memcpy(out + ((size_t)t * dch + c) * outsz, p->decibels(c), outsz * sizeof(float));
        if(out_silent)
            out_silent[t] = p->last_silent() ? 1 : 0;
    }
    return n_ticks;
}

} // extern "C"
