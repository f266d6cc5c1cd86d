// spectrum_source.cpp — see spectrum_source.hpp.  Host plumbing only; every spectrum value comes from libwfstft.
#include "spectrum_source.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace wfhost {

// ---- RingBuffer (≙ CircularBuffer, src/circular_buffer.hpp) ----------------------------------------------------
void RingBuffer::reset()
{
    m_head = 0;
    m_size = 0;
}
void RingBuffer::reserve(size_t bytes)
{
    if(bytes <= m_buf.size())
        return;
    const size_t cap = (bytes + 1023) & ~size_t(1023); // 1 KiB steps, src/circular_buffer.hpp:29-41
    std::vector<uint8_t> n(cap);
    if(m_size)
    {
        const size_t first = std::min(m_size, m_buf.size() - m_head);
        memcpy(n.data(), m_buf.data() + m_head, first);
        memcpy(n.data() + first, m_buf.data(), m_size - first);
    }
    m_buf.swap(n);
    m_head = 0;
}
void RingBuffer::push_back(const void *data, size_t bytes)
{
    if(bytes == 0) // nothing to do — and no modulo by an empty buffer's size (the reference guards size == 0 too,
        return;    // src/circular_buffer.hpp:46,70)
    reserve(m_size + bytes);
    const size_t tail = (m_head + m_size) % m_buf.size();
    const size_t first = std::min(bytes, m_buf.size() - tail);
    memcpy(m_buf.data() + tail, data, first);
    memcpy(m_buf.data(), (const uint8_t *)data + first, bytes - first);
    m_size += bytes;
}
void RingBuffer::push_back_zero(size_t bytes)
{
    if(bytes == 0)
        return;
    reserve(m_size + bytes);
    const size_t tail = (m_head + m_size) % m_buf.size();
    const size_t first = std::min(bytes, m_buf.size() - tail);
    memset(m_buf.data() + tail, 0, first);
    memset(m_buf.data(), 0, bytes - first);
    m_size += bytes;
}
void RingBuffer::peek_front(void *dst, size_t bytes) const
{
    if(bytes == 0 || m_buf.empty())
        return;
    const size_t first = std::min(bytes, m_buf.size() - m_head);
    memcpy(dst, m_buf.data() + m_head, first);
    memcpy((uint8_t *)dst + first, m_buf.data(), bytes - first);
}
void RingBuffer::pop_front(void *dst, size_t bytes)
{
    if(dst)
        peek_front(dst, bytes);
    if(bytes)
    {
        m_head = (m_head + bytes) % m_buf.size();
        m_size -= bytes;
    }
}

// ---- SpectrumSourceCUDA ----------------------------------------------------------------------------------------
SpectrumSourceCUDA::~SpectrumSourceCUDA()
{
    if(m_engine)
        wf_destroy(m_engine);
}

const char *SpectrumSourceCUDA::last_error() const { return wf_last_error(m_engine); }

// libobs' integer ns <-> frames conversion
static inline uint64_t ns_to_frames(uint64_t sample_rate, uint64_t ns)
{
    return (uint64_t)(((__uint128_t)ns * sample_rate) / 1000000000ull);
}
static inline uint64_t frames_to_ns(uint64_t sample_rate, uint64_t frames)
{
    return (uint64_t)(((__uint128_t)frames * 1000000000ull) / sample_rate);
}

int SpectrumSourceCUDA::update(const wf_config &cfg, int64_t ts_offset_ns, uint64_t now_ns)
{
    // release + free_bufs, src/source.cpp:1082-1083
    if(m_engine)
    {
        wf_destroy(m_engine);
        m_engine = nullptr;
    }
    for(auto &b : m_capturebufs)
        b.reset();
    m_capture_ts = m_audio_ts = 0;

    m_cfg = cfg;
    m_cfg.struct_size = sizeof(wf_config);
    m_cfg.max_streams = 1; // one WAVSource = one stream
    int rc = wf_create(&m_cfg, &m_engine);
    if(rc != WF_OK)
        return rc;
    wf_get_info(m_engine, &m_info);
    m_ts_offset = ts_offset_ns;
    m_frame.assign((size_t)m_info.capture_channels * m_info.fft_size, 0.0f);
    m_out.assign((size_t)m_info.display_channels * m_info.bins, m_info.db_min);
    for(int c = 0; c < 2; ++c)
        m_decibels[c].assign((size_t)m_info.bins, m_info.db_min); // src/source.cpp:1181
    m_last_silent = false;
    // volume normalisation feed, src/source.cpp:1145-1152
    m_rms_sync.reset();
    m_rms_ring.clear();
    m_rms_pos = 0;
    m_input_rms = 0.0f;
    if(m_cfg.normalize_volume)
        m_rms_ring.assign((size_t)(m_cfg.sample_rate & ~15u), 0.0f);
    m_capture_ts = now_ns; // src/source.cpp:1242
    // pre-fill with silence to avoid start-up lag, src/source.cpp:1243-1248
    for(int c = 0; c < m_info.capture_channels; ++c)
        m_capturebufs[c].push_back_zero((size_t)m_info.fft_size * sizeof(float));
    return WF_OK;
}

// get_audio_sync, src/source.hpp:279-285
int64_t SpectrumSourceCUDA::audio_sync(uint64_t ts) const
{
    const uint64_t audio_ts = m_audio_ts + (uint64_t)m_ts_offset;
    uint64_t delta = std::max(audio_ts, ts) - std::min(audio_ts, ts);
    delta = std::min(delta, MAX_TS_DELTA);
    return (audio_ts < ts) ? -(int64_t)delta : (int64_t)delta;
}

// capture_audio, src/source.cpp:1817-1888
void SpectrumSourceCUDA::capture_audio(const float *const *data, uint32_t frames, uint64_t timestamp_ns, uint64_t now_ns,
                                       bool muted)
{
    if(!m_engine || m_info.capture_channels == 0)
        return;
    const uint32_t sr = m_cfg.sample_rate;
    m_capture_ts = now_ns;
    const uint64_t audio_len = frames_to_ns(sr, frames);
    const uint64_t delta = std::max(timestamp_ns, m_capture_ts) - std::min(timestamp_ns, m_capture_ts);
    m_audio_ts = (delta > MAX_TS_DELTA) ? m_capture_ts : timestamp_ns + audio_len; // bogus-timestamp clamp :1833-1837
    const size_t bufsz = (size_t)m_info.fft_size * sizeof(float);
    const int64_t dtaudio = audio_sync(m_capture_ts);
    const size_t dtsamples = (dtaudio > 0) ? (size_t)ns_to_frames(sr, (uint64_t)dtaudio) : 0;
    // RMS pre-accumulate: one value per time point, the square of the LARGEST channel (src/source.cpp:1842-1871)
    if(m_cfg.normalize_volume)
    {
        m_rms_tmp.resize(frames);
        for(uint32_t i = 0; i < frames; ++i)
        {
            float val = 0.0f;
            for(int j = 0; j < m_info.capture_channels; ++j)
                if(data[j] != nullptr)
                    val = std::max(std::abs(data[j][i]), val);
            m_rms_tmp[i] = val * val;
        }
        m_rms_sync.push_back(m_rms_tmp.data(), (size_t)frames * sizeof(float));
        const size_t max_rms = (dtsamples + m_rms_ring.size()) * sizeof(float);
        if(m_rms_sync.size() > max_rms)
            m_rms_sync.pop_front(nullptr, m_rms_sync.size() - max_rms);
    }
    const size_t sz = (size_t)frames * sizeof(float);
    for(int j = 0; j < m_info.capture_channels; ++j)
    {
        if(muted || data[j] == nullptr)
            m_capturebufs[j].push_back_zero(sz);
        else
            m_capturebufs[j].push_back(data[j], sz);
        const size_t max_size = (dtsamples * sizeof(float)) + bufsz; // keep delay + N samples, :1883-1886
        const size_t total = m_capturebufs[j].size();
        if(total > max_size)
            m_capturebufs[j].pop_front(nullptr, total - max_size);
    }
}

// sync_rms_buffer (src/source.cpp:810-836) + update_input_rms (src/source_generic.cpp:392-403): move the values that are
// due (everything but the A/V-sync delay) into the one-second ring, then sqrt(mean) with ONE float accumulator in ring-index
// order (the reference's rounding; the order is not time order once the ring has wrapped).
void SpectrumSourceCUDA::update_input_rms()
{
    const int64_t dtaudio = audio_sync(m_tick_ts);
    const size_t dtsize = (dtaudio > 0) ? (size_t)ns_to_frames(m_cfg.sample_rate, (uint64_t)dtaudio) * sizeof(float) : 0;
    if(m_rms_sync.size() <= dtsize)
        return;
    const size_t n = m_rms_ring.size();
    while(m_rms_sync.size() > dtsize)
    {
        const size_t consume = m_rms_sync.size() - dtsize;
        const size_t room = (n - m_rms_pos) * sizeof(float);
        if(consume >= room)
        {
            m_rms_sync.pop_front(&m_rms_ring[m_rms_pos], room);
            m_rms_pos = 0;
        }
        else
        {
            m_rms_sync.pop_front(&m_rms_ring[m_rms_pos], consume);
            m_rms_pos += consume / sizeof(float);
        }
    }
    float sum = 0.0f;
    for(size_t i = 0; i < n; ++i)
        sum += m_rms_ring[i];
    m_input_rms = std::sqrt(sum / n);
}

// tick (src/source.cpp:1324-1344) + the host half of tick_spectrum (src/source_generic.cpp:26-61)
int SpectrumSourceCUDA::tick(float seconds, uint64_t now_ns)
{
    if(!m_engine || m_info.capture_channels == 0)
        return WF_OK;
    m_tick_ts = now_ns;
    if(m_cfg.normalize_volume)
        update_input_rms(); // src/source.cpp:1330-1331
    const size_t N = (size_t)m_info.fft_size, B = (size_t)m_info.bins;
    const size_t bufsz = N * sizeof(float);

    // timeout / hidden: reset state once and blank the graph, src/source_generic.cpp:36-48
    const uint64_t dtcapture = m_tick_ts - m_capture_ts;
    if(!m_show || (dtcapture > CAPTURE_TIMEOUT))
    {
        if(m_last_silent)
            return WF_OK;
        int rc = wf_reset_state(m_engine, 0, 1);
        if(rc != WF_OK)
            return rc;
        for(int c = 0; c < m_info.display_channels; ++c)
            std::fill(m_decibels[c].begin(), m_decibels[c].end(), m_info.db_min);
        m_last_silent = true;
        return WF_OK;
    }

    // A/V sync: take the OLDEST N of the last (delay + N) samples, src/source_generic.cpp:50-59
    const int64_t dtaudio = audio_sync(m_tick_ts);
    const size_t dtsize =
        ((dtaudio > 0) ? (size_t)ns_to_frames(m_cfg.sample_rate, (uint64_t)dtaudio) * sizeof(float) : 0) + bufsz;
    uint8_t skip = 0;
    for(int c = 0; c < m_info.capture_channels; ++c)
    {
        if(m_capturebufs[c].size() >= dtsize)
        {
            m_capturebufs[c].pop_front(nullptr, m_capturebufs[c].size() - dtsize);
            m_capturebufs[c].peek_front(m_frame.data() + (size_t)c * N, bufsz);
        }
        else
            skip = 1; // "not enough audio": both rings fill together, so this is per tick
    }

    wf_batch b{};
    b.struct_size = sizeof(wf_batch);
    b.n_streams = 1;
    b.n_frames = 1;
    b.hop = (int32_t)N;
    b.seconds = seconds;
    b.pcm = m_frame.data();
    b.stream_stride = (int64_t)(m_info.capture_channels * N);
    b.channel_stride = (int64_t)N;
    b.skip_mask = &skip;
    b.input_rms = m_cfg.normalize_volume ? &m_input_rms : nullptr;
    b.out_db = m_out.data();
    uint8_t silent = 0;
    b.out_silent = &silent;
    int rc = wf_process(m_engine, &b);
    if(rc != WF_OK)
        return rc;
    m_last_silent = silent != 0;
    for(int c = 0; c < m_info.display_channels; ++c)
        memcpy(m_decibels[c].data(), m_out.data() + (size_t)c * B, B * sizeof(float));
    return WF_OK;
}

} // namespace wfhost
