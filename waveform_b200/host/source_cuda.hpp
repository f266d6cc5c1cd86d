// source_cuda.hpp — the binding a maintainer of phandasm/waveform adds to put libwfstft.so behind the plugin's spectrum
// backend: one more WAVSource subclass next to WAVSourceGeneric / WAVSourceAVX / WAVSourceAVX2 (src/source.hpp:349-386),
// selected in callbacks::create (src/source.cpp:87-102).  Compiled in the plugin's own tree (it includes the plugin's
// source.hpp); this repository compiles it against the UNMODIFIED reference sources + the fake libobs of
// oracle/ref_build to prove the seam (oracle/ref_build/Makefile target libwaveform_ref_cuda.so, tests/test_gpu_seam.py).
//
// Everything above tick_spectrum() — OBS properties, the audio callback, CircularBuffer, A/V sync, update_input_rms(),
// render() — is the reference's own code, untouched.  Only the per-frame pipeline of src/source_generic.cpp:63-180 runs on
// the GPU, through the C-ABI of include/wfstft.h.
#pragma once
#include <algorithm>
#include <cstring>
#include <mutex>

#include "log.hpp"    // the plugin's (LogError)
#include "source.hpp" // the plugin's
#include "wfstft.h"

class WAVSourceCUDA : public WAVSourceGeneric // meter / waveform modes and update_input_rms stay on the generic path
{
    wf_engine *m_engine = nullptr;
    // page-locked, device-mapped staging (wf_host_alloc): the kernel reads the frame and writes the spectrum in place,
    // so a tick is one launch + one synchronisation, no copies
    float *m_frame = nullptr;   // [capture_channels][N]
    float *m_out = nullptr;     // [display_channels][N/2]
    float *m_rms = nullptr;     // [1] m_input_rms of the tick
    uint8_t *m_flags = nullptr; // [0] skip ("not enough audio"), [1] m_last_silent after the tick

    void release()
    {
        if(m_engine)
            wf_destroy(m_engine);
        m_engine = nullptr;
        wf_host_free(m_frame);
        wf_host_free(m_out);
        wf_host_free(m_rms);
        wf_host_free(m_flags);
        m_frame = m_out = m_rms = nullptr;
        m_flags = nullptr;
    }

public:
    using WAVSourceGeneric::WAVSourceGeneric;
    ~WAVSourceCUDA() override { release(); }

    void update(obs_data_t *settings) override
    {
        WAVSourceGeneric::update(settings); // get_settings(), ring buffers, tables, m_* members (src/source.cpp:1077-1322)
        std::lock_guard lock(m_mtx);
        release();
        if(m_meter_mode || m_display_mode == DisplayMode::WAVEFORM || m_capture_channels == 0)
            return;
        wf_config c;
        wf_config_init(&c);
        c.max_streams = 1; // one WAVSource = one stream
        c.sample_rate = m_audio_info.samples_per_sec;
        c.capture_channels = (int32_t)m_capture_channels;
        c.fft_size = (int32_t)m_fft_size;
        c.window = (int32_t)m_window_func; // FFTWindow / TSmoothingMode share the order of wf_window / wf_tsmooth
        c.sine_exponent = m_sine_exponent;
        c.tsmoothing = (int32_t)m_tsmoothing;
        c.gravity = m_gravity;
        c.fast_peaks = m_fast_peaks;
        c.slope = m_slope;
        c.rolloff_q = m_rolloff_q;
        c.rolloff_rate = m_rolloff_rate;
        c.cutoff_low = m_cutoff_low;
        c.cutoff_high = m_cutoff_high;
        c.floor_db = m_floor;
        c.ceiling_db = m_ceiling;
        c.stereo = m_stereo;
        c.normalize_volume = m_normalize_volume;
        c.volume_target = m_volume_target;
        c.max_gain = m_max_gain;
        // display settings keep their defaults: interpolation stays in render() here (INTEGRATION.md shows how to move it
        // into the kernel epilogue with out_points)
        if(wf_create(&c, &m_engine) != WF_OK)
        {
            LogError << "wfstft: " << wf_last_error(nullptr);
            m_engine = nullptr;
            return;
        }
        const size_t outsz = m_fft_size / 2;
        m_frame = (float *)wf_host_alloc(m_capture_channels * m_fft_size * sizeof(float));
        m_out = (float *)wf_host_alloc((m_stereo ? 2 : 1) * outsz * sizeof(float));
        m_rms = (float *)wf_host_alloc(sizeof(float));
        m_flags = (uint8_t *)wf_host_alloc(16);
        if(!m_frame || !m_out || !m_rms || !m_flags)
        {
            LogError << "wfstft: pinned staging allocation failed";
            release();
        }
    }

protected:
    void tick_spectrum(float seconds) override // replaces src/source_generic.cpp:26-180
    {
        if(!m_engine)
            return;
        const auto bufsz = m_fft_size * sizeof(float);
        const auto outsz = m_fft_size / 2;
        const auto dch = m_stereo ? 2 : 1;
        if(!m_show || ((m_tick_ts - m_capture_ts) > CAPTURE_TIMEOUT)) // :36-48
        {
            if(m_last_silent)
                return;
            wf_reset_state(m_engine, 0, 1);
            for(auto ch = 0; ch < dch; ++ch)
                std::fill_n(m_decibels[ch].get(), outsz, DB_MIN);
            m_last_silent = true;
            return;
        }
        const int64_t dtaudio = get_audio_sync(m_tick_ts); // :50-51
        const size_t dtsize =
            ((dtaudio > 0) ? size_t(ns_to_audio_frames(m_audio_info.samples_per_sec, (uint64_t)dtaudio)) * sizeof(float) : 0) + bufsz;
        m_flags[0] = 0;
        for(auto ch = 0u; ch < m_capture_channels; ++ch) // :55-61, the frame fetch stays on the host
        {
            if(m_capturebufs[ch].size() >= dtsize)
            {
                m_capturebufs[ch].pop_front(nullptr, m_capturebufs[ch].size() - dtsize);
                m_capturebufs[ch].peek_front(m_frame + ch * m_fft_size, bufsz);
            }
            else
                m_flags[0] = 1; // not enough audio: both rings fill together
        }
        *m_rms = m_input_rms; // WAVSource::tick already ran update_input_rms() (src/source.cpp:1330-1331)
        wf_batch b{};
        b.struct_size = sizeof(b);
        b.n_streams = 1;
        b.n_frames = 1;
        b.hop = (int32_t)m_fft_size;
        b.seconds = seconds;
        b.pcm = m_frame;
        b.stream_stride = (int64_t)(m_capture_channels * m_fft_size);
        b.channel_stride = (int64_t)m_fft_size;
        b.input_rms = m_normalize_volume ? m_rms : nullptr;
        b.skip_mask = m_flags;
        b.out_db = m_out;
        b.out_silent = m_flags + 1;
        if(wf_process(m_engine, &b) != WF_OK)
        {
            LogError << "wfstft: " << wf_last_error(m_engine);
            return;
        }
        m_last_silent = m_flags[1] != 0;
        for(auto ch = 0; ch < dch; ++ch)
            memcpy(m_decibels[ch].get(), m_out + ch * outsz, outsz * sizeof(float));
    }
};
