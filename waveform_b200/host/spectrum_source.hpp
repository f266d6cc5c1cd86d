// spectrum_source.hpp — host-side C++ mirror of the plugin's capture/tick plumbing for the CUDA backend.
//
// This is what a `WAVSourceCUDA : WAVSource` subclass boils down to once OBS is factored out: the audio callback,
// the per-channel ring buffers, the A/V-sync frame selection and the timeout handling stay on the host exactly as in
// the reference, and only the per-frame spectrum pipeline goes through the C-ABI (include/wfstft.h).
//
// Reference counterparts (paths relative to the reference tree):
//   RingBuffer                     ≙ CircularBuffer               src/circular_buffer.hpp:29-139
//   SpectrumSourceCUDA::update     ≙ WAVSource::update            src/source.cpp:1077-1322 (buffers, pre-fill :1243-1248)
//   SpectrumSourceCUDA::capture_audio ≙ WAVSource::capture_audio  src/source.cpp:1817-1888
//   SpectrumSourceCUDA::tick       ≙ WAVSource::tick + tick_spectrum src/source.cpp:1324-1344, src/source_generic.cpp:26-61
//   audio_sync                     ≙ WAVSource::get_audio_sync    src/source.hpp:279-285
//   update_input_rms               ≙ sync_rms_buffer + WAVSourceGeneric::update_input_rms  src/source.cpp:810-836,
//                                    src/source_generic.cpp:392-403 (the RMS pre-accumulate of capture_audio: src/source.cpp:1842-1871)
// Threading is the caller's, as in the plugin (m_mtx around tick/render/update; capture_audio try-locks).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "wfstft.h"

namespace wfhost {

// Byte-granular FIFO of float samples (grows in 1 KiB steps like the reference's).
class RingBuffer {
public:
    size_t size() const { return m_size; } // bytes
    void reset();
    void push_back(const void *data, size_t bytes);
    void push_back_zero(size_t bytes);
    void pop_front(void *dst, size_t bytes); // dst may be null
    void peek_front(void *dst, size_t bytes) const;

private:
    void reserve(size_t bytes);
    std::vector<uint8_t> m_buf;
    size_t m_head = 0, m_size = 0;
};

class SpectrumSourceCUDA {
public:
    static constexpr uint64_t CAPTURE_TIMEOUT = 1000000ull * 500u;  // src/source.hpp:290
    static constexpr uint64_t MAX_TS_DELTA = 1000000000ull * 16u;   // src/source.hpp:291

    SpectrumSourceCUDA() = default;
    ~SpectrumSourceCUDA();
    SpectrumSourceCUDA(const SpectrumSourceCUDA &) = delete;
    SpectrumSourceCUDA &operator=(const SpectrumSourceCUDA &) = delete;

    // (Re)build everything for new settings; returns a wf_status.  ts_offset_ns ≙ m_ts_offset (audio sync offset).
    int update(const wf_config &cfg, int64_t ts_offset_ns, uint64_t now_ns);
    // Audio thread: planar float channels (data[c] may be null = silence), packet timestamp as OBS provides it.
    void capture_audio(const float *const *data, uint32_t frames, uint64_t timestamp_ns, uint64_t now_ns, bool muted);
    // Graphics thread, once per video frame.  Returns a wf_status (WF_OK also when the tick was a no-op).
    int tick(float seconds, uint64_t now_ns);

    void show() { m_show = true; }
    void hide() { m_show = false; }
    const float *decibels(int display_channel) const { return m_decibels[display_channel].data(); }
    bool last_silent() const { return m_last_silent; }
    float input_rms() const { return m_input_rms; } // m_input_rms: RMS over the last second, volume normalisation's input
    int bins() const { return m_info.bins; }
    int display_channels() const { return m_info.display_channels; }
    const wf_info &info() const { return m_info; }
    const char *last_error() const;

private:
    int64_t audio_sync(uint64_t ts) const;
    void update_input_rms();
    wf_engine *m_engine = nullptr;
    wf_config m_cfg{};
    wf_info m_info{};
    RingBuffer m_capturebufs[2];
    std::vector<float> m_frame;          // [capture_channels][N] staging handed to wf_process
    std::vector<float> m_out;            // [display_channels][bins]
    std::vector<float> m_decibels[2];
    bool m_show = true, m_last_silent = false;
    uint64_t m_capture_ts = 0, m_audio_ts = 0, m_tick_ts = 0;
    int64_t m_ts_offset = 0;
    // volume normalisation feed (≙ m_rms_sync_buf / m_input_rms_buf / m_input_rms, src/source.hpp:238-244): per captured
    // sample (max over channels |x|)^2, A/V-synchronised, RMS over the last sample_rate & -16 values
    RingBuffer m_rms_sync;
    std::vector<float> m_rms_ring, m_rms_tmp;
    size_t m_rms_pos = 0;
    float m_input_rms = 0.0f;
};

} // namespace wfhost
