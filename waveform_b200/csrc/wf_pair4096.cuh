// wf_pair4096.cuh — N = 4096 (the plugin's DEFAULT fft size, src/source.cpp:119-174), one capture channel, spectrum output:
// the warp-synchronous machinery of wf_fast2048.cuh with TWO warps per stream.
//
// The packed 2048-point complex FFT of a frame is split by one radix-2 decimation-in-frequency stage across the pair:
//     warp 0:  e[n] =  z[n] + z[n+1024]                 -> X[2k']   = FFT1024(e)[k']
//     warp 1:  o[n] = (z[n] - z[n+1024]) W_2048^n       -> X[2k'+1] = FFT1024(o)[k']
// Each warp reads BOTH halves of the TMA-staged frame (window multiply folded in), forms its own 1024-point sequence in
// registers and then runs exactly the N=2048 kernel's FFT (two radix-32 register passes, one padded shared-memory
// transpose, no block barrier).  The real-FFT split pairs bin k with 2048-k: even bins pair with even bins and odd with
// odd, so every pair lives inside ONE warp — k' <-> 1024-k' for warp 0 (the N=2048 kernel's pairing and twiddles
// W_2048^k'), k' <-> 1023-k' for warp 1 (twiddles W_4096^(2k'+1), partner lane 31-lane, no special cases).  EMA state stays
// in shared memory per warp; the only cross-warp traffic is one 64-thread named barrier per frame (both halves read) and
// two flag bytes for the silence gate's all-bins test.
//
// Replaces stft_v3_kernel<4096,1,1,*> (one 128-thread CTA per stream, six block barriers per frame, 49 % of the HBM
// roofline) where there are enough streams to fill the SMs' eight pairs.  Semantics: src/source_generic.cpp:26-180.
#pragma once
#include "wf_fast2048.cuh"
#include "wf_team2048.cuh" // team::bar_sync

namespace wf {

namespace pair4096 {
constexpr int kFrame = 4096;
constexpr int kBins = 2048;  // packed complex points = bins
constexpr int kSub = 1024;  // sub-FFT length = bins per warp
constexpr int kPWarps = 16;
constexpr int kPairs = kPWarps / 2;
constexpr int kPBufBytes = fast::kWarpBufBytes; // TMA landing zone of this warp's half frame (8192 B) / transpose area
constexpr int kPStateBytes = fast::kStateBytes;     // EMA state of this warp's 1024 bins: [pair q][lane] -> (first, second)
constexpr int kPWarpBytes = kPBufBytes + kPStateBytes + 16;
// tables: twA[1024] (inter-pass, as N=2048) | tw1[1024] = W_2048^n (first stage; its first 512 entries are warp 0's split
// twiddles) | twPo[512] = W_4096^(2k'+1), k' = lane + 32 q (warp 1's split twiddles)
constexpr int kPTableBytes = (1024 + 1024 + 512) * 8;
constexpr int kPCtlBytes = 16; // per team: outs[2][2] bytes
constexpr int smem_bytes() { return kPTableBytes + kPWarps * kPWarpBytes + kPairs * kPCtlBytes; }
} // namespace pair4096

template<bool EXTRA>
__global__ void __launch_bounds__(pair4096::kPWarps * 32, 1) stft4096_pair_kernel(const __grid_constant__ KParams p)
{
    using namespace fast;
    using namespace pair4096;
    constexpr int B = kBins;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *s_twA = reinterpret_cast<float2 *>(smem_raw); // [k2][n1] = W_1024^(k2*n1)
    float2 *s_tw1 = s_twA + 1024;                         // [n] = W_2048^n, n < 1024
    float2 *s_twPo = s_tw1 + 1024;                        // [q][lane] = W_4096^(2(lane + 32 q) + 1), q < 16
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tm = warp >> 1;
    const int half = warp & 1;
    unsigned char *warps_base = reinterpret_cast<unsigned char *>(s_twPo + 512);
    unsigned char *wbase = warps_base + warp * kPWarpBytes;
    unsigned char *pbase = warps_base + (warp ^ 1) * kPWarpBytes;
    float2 *buf = reinterpret_cast<float2 *>(wbase);
    const pk::c64 *buf_lo = reinterpret_cast<const pk::c64 *>(warps_base + (2 * tm) * kPWarpBytes);     // z[n], n < 1024
    const pk::c64 *buf_hi = reinterpret_cast<const pk::c64 *>(warps_base + (2 * tm + 1) * kPWarpBytes); // z[n + 1024]
    float2 *sst = reinterpret_cast<float2 *>(wbase + kPBufBytes) + lane;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(wbase + kPBufBytes + kPStateBytes);
    uint64_t *mbar_peer = reinterpret_cast<uint64_t *>(pbase + kPBufBytes + kPStateBytes);
    volatile unsigned char *ctl = warps_base + kPWarps * kPWarpBytes + tm * kPCtlBytes; // [parity][half]
    const int bar_id = 1 + tm;

    for(int i = threadIdx.x; i < 1024; i += blockDim.x)
    {
        s_twA[i] = __ldg(p.tw_h + (((i >> 5) * (i & 31)) & 1023)); // W_1024^k table of the half size
        s_tw1[i] = __ldg(p.tw + i);                                 // W_2048^n
        if(i < 512)
            s_twPo[i] = __ldg(p.tw_post + 2 * i + 1);               // W_4096^(2k'+1), k' = i = lane + 32 q
    }
    if(lane == 0)
    {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int S = p.n_streams, T = p.n_frames;
    const int G = gridDim.x;
    const int n_local = (S > (int)blockIdx.x) ? (S - (int)blockIdx.x + G - 1) / G : 0;
    uint32_t phase = 0;
    const bool tsm = p.tsmooth != 0, gate = p.gate != 0;
    const pk::c64 *win = reinterpret_cast<const pk::c64 *>(p.window2s); // window pairs x (2/sum(w))/2, global (L1-resident)

    // bins of this lane, in units of k' (index inside this warp's sub-FFT): first of pair q -> k' = lane + 32 q
    // warp 0: second -> 1024 - k' = kb + 32 (31 - q)  (lane 0: 32 (32 - q); q == 0: k' = 512 rides in the unused slot)
    // warp 1: second -> 1023 - k' = (31 - lane) + 32 (31 - q)
    const int jp = half ? (31 - lane) : ((32 - lane) & 31);
    const int kb = half ? (31 - lane) : (((32 - lane) & 31) + (lane == 0 ? 32 : 0));
    auto second_of = [&](int q) -> int { return (!half && q == 0) ? ((lane == 0) ? 512 : (kb + 992)) : (kb + 32 * (31 - q)); };

    if(tm < n_local && lane == 0)
    {
        mbar_expect_tx(mbar, kBins * 4);
        tma_load_1d(buf, p.pcm + (size_t)(blockIdx.x + tm * G) * p.stream_stride + half * kBins, kBins * 4, mbar);
    }

    for(int li = tm; li < n_local; li += kPairs)
    {
        const int s = (int)blockIdx.x + li * G;
        const bool have_next_stream = (li + kPairs) < n_local;
        float *state_s = p.state + (size_t)s * B;
        float *hold_s = p.hold_db + (size_t)s * B;
        const float *pcm_s = p.pcm + (size_t)s * p.stream_stride;
        // ---- per-stream state: global (natural bin order, bin = 2 k' + half) -> shared ([pair][lane]) ----
#pragma unroll
        for(int q = 0; q < 16; ++q)
            sst[q * 32] = make_float2(state_s[2 * (lane + 32 * q) + half], state_s[2 * second_of(q) + half]);
        const unsigned char fl = p.flags[s];
        bool last_silent = (fl & 1u) != 0;
        bool pos_w = (fl & 2u) != 0; // this warp's share of "all outputs of the last producing tick <= floor-10 dB"
        const bool hold_lazy = (fl & 8u) != 0;
        bool last_from_state = false;
        if(lane == 0)
            ctl[2 + half] = pos_w ? 1 : 0; // slot of "tick -1" (parity 1)

#pragma unroll 1
        for(int t = 0; t < T; ++t)
        {
            // ---- both halves of the frame from shared (TMA-staged), window, radix-2 first stage ----
            mbar_wait(mbar, phase);
            mbar_wait(mbar_peer, phase);
            phase ^= 1u;
            pk::c64 v[32];
            unsigned long long nzbits = 0;
#pragma unroll
            for(int pidx = 0; pidx < 32; ++pidx)
            {
                const int n = lane + 32 * pidx;
                const pk::c64 za = buf_lo[n], zb = buf_hi[n];
                nzbits |= za | zb;
                const pk::c64 a = pk::mul(za, __ldg(win + n));
                if(half == 0)
                    v[pidx] = pk::fma(zb, __ldg(win + n + kSub), a);
                else
                {
                    const pk::c64 d = pk::sub(a, pk::mul(zb, __ldg(win + n + kSub)));
                    v[pidx] = pk::cmul(d, reinterpret_cast<const pk::c64 *>(s_tw1)[n]);
                }
            }
            const bool nz = __any_sync(0xffffffffu, (nzbits & 0x7fffffff7fffffffull) != 0ull);
            team::bar_sync(bar_id, 64); // both warps have read both halves: the buffers become transpose areas
            const bool pos = (ctl[2 * ((t + 1) & 1)] != 0) && (ctl[2 * ((t + 1) & 1) + 1] != 0); // the previous tick's flags

#pragma unroll
            for(int pass = 0; pass < 2; ++pass)
            {
                pk::dft_bitrev<32>(v);
                if(pass == 0)
                {
                    const pk::c64 *buf64 = reinterpret_cast<const pk::c64 *>(buf);
#pragma unroll
                    for(int k2 = 0; k2 < 32; ++k2)
                    {
                        pk::c64 a = v[bitrev<32>(k2)];
                        if(k2 > 0)
                            a = pk::cmul(a, reinterpret_cast<const pk::c64 *>(s_twA)[k2 * 32 + lane]);
                        reinterpret_cast<pk::c64 *>(buf)[lane * 33 + k2] = a;
                    }
                    __syncwarp();
#pragma unroll
                    for(int n1 = 0; n1 < 32; ++n1)
                        v[n1] = buf64[n1 * 33 + lane];
                    __syncwarp(); // this warp's buffer can take its half of the next frame
                    if(t + 1 == T && have_next_stream)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.state + (size_t)(s + kPairs * G) * B + half * kSub + lane * 32));
                    if(lane == 0)
                    {
                        const float *next = nullptr;
                        if(t + 1 < T)
                            next = pcm_s + (size_t)(t + 1) * p.hop;
                        else if(have_next_stream)
                            next = p.pcm + (size_t)(s + kPairs * G) * p.stream_stride;
                        if(next != nullptr)
                        {
                            fence_proxy_async();
                            mbar_expect_tx(mbar, kBins * 4);
                            tma_load_1d(buf, next + half * kBins, kBins * 4, mbar);
                        }
                    }
                }
            }
            // now X[2 (lane + 32 k1) + half] = v[bitrev(k1)]

            // ---- gate (src/source_generic.cpp:63-95): both warps take the same decisions from the same inputs ----
            const bool skip_all = EXTRA && (p.skip_mask != nullptr) && (p.skip_mask[(size_t)s * T + t] != 0);
            bool do_proc = !skip_all;
            if(!skip_all)
            {
                if(nz)
                    last_silent = false;
                else if(gate)
                {
                    if(last_silent)
                        do_proc = false;
                    else if(pos)
                    {
                        last_silent = true;
                        do_proc = false;
                    }
                }
            }
            float *odb = p.out_db + ((size_t)s * T + t) * B + half;
            float vc = 0.0f;
            if(EXTRA && p.normalize)
            {
                const float rms = (p.input_rms != nullptr) ? p.input_rms[(size_t)s * T + t] : 0.0f;
                vc = fminf(p.vol_target - dbfs(rms, p.db_min), p.max_gain);
            }
            bool outs = true;
            float peak = -INFINITY;
            const float2 gt = (EXTRA && p.g_tab != nullptr) ? __ldg(p.g_tab + t) : make_float2(p.g, p.g2);

            if(do_proc && !last_silent)
            {
#pragma unroll
                for(int q = 0; q < 16; ++q)
                {
                    const int k1 = 2 * (lane + 32 * q);  // + half through odb / table pointers
                    const int k2 = 2 * second_of(q);
                    const pk::c64 a = v[bitrev<32>(q)];
                    unsigned long long bp = __shfl_sync(0xffffffffu, v[bitrev<32>(31 - q)], jp);
                    if(!half && lane == 0)
                        bp = v[bitrev<32>((32 - q) & 31)];
                    const pk::c64 b = pk::conj(bp);
                    const pk::c64 sum = pk::add(a, b);
                    const pk::c64 o = pk::mul_neg_i(pk::sub(a, b));
                    const pk::c64 w = half ? reinterpret_cast<const pk::c64 *>(s_twPo)[q * 32 + lane]
                                           : reinterpret_cast<const pk::c64 *>(s_tw1)[q * 32 + lane];
                    const pk::c64 wo = pk::cmul(o, w);
                    const pk::c64 y1 = pk::add(sum, wo);
                    const pk::c64 y2 = pk::sub(sum, wo);
                    const pk::c64 s1 = pk::mul(y1, y1), s2 = pk::mul(y2, y2);
                    float p1 = pk::re(s1) + pk::im(s1);
                    float p2 = pk::re(s2) + pk::im(s2);
                    if(q == 0)
                    {
                        // warp 0, lane 0: the pair (0, 2048) has no bin 2048; its second slot carries bin 1024 (k' = 512)
                        const pk::c64 x512 = v[bitrev<32>(16)];
                        const pk::c64 sq = pk::mul(x512, x512);
                        const float p512 = 4.0f * (pk::re(sq) + pk::im(sq));
                        p2 = (!half && lane == 0) ? p512 : p2;
                    }
                    pk::c64 m = pk::make(sqrt_approx(p1), sqrt_approx(p2));
                    if(EXTRA && p.slope != nullptr)
                        m = pk::mul(m, pk::make(__ldg(p.slope + k1 + half), __ldg(p.slope + k2 + half)));
                    pk::c64 *sst64 = reinterpret_cast<pk::c64 *>(sst);
                    if(tsm)
                    {
                        pk::c64 old = sst64[q * 32];
                        if(EXTRA && p.fast_peaks)
                            old = pk::make(fmaxf(pk::re(m), pk::re(old)), fmaxf(pk::im(m), pk::im(old)));
                        m = pk::fma(pk::make(gt.x, gt.x), old, pk::mul(pk::make(gt.y, gt.y), m));
                    }
                    sst64[q * 32] = m;
                    float d1, d2;
                    pk::split(dbfs2(pk::re(m), pk::im(m), p.db_min), d1, d2);
                    if(EXTRA)
                    {
                        if(p.normalize)
                        {
                            if(k1 + half >= 1)
                                d1 += vc;
                            d2 += vc;
                        }
                        if(p.rolloff != nullptr)
                        {
                            if(k1 + half >= 1)
                                d1 = fmaxf(d1 - __ldg(p.rolloff + k1 + half), p.db_min);
                            d2 = fmaxf(d2 - __ldg(p.rolloff + k2 + half), p.db_min);
                        }
                        if(k1 + half >= 1)
                            peak = fmaxf(peak, d1);
                        peak = fmaxf(peak, d2);
                    }
                    outs &= !(d1 > p.floor_m10) & !(d2 > p.floor_m10);
                    stg_stream(odb + k1, d1);
                    stg_stream(odb + k2, d2);
                }
                last_from_state = true;
            }
            else
            {
                // tick returned early (hold) or the channel was skipped while the tick went on (stale dB re-converted)
                const float *prev_db = (t > 0) ? (odb - B) : (hold_s + half);
                const bool from_state = (t == 0) && hold_lazy;
#pragma unroll 1
                for(int q = 0; q < 16; ++q)
                {
                    const int k1 = 2 * (lane + 32 * q);
                    const int k2 = 2 * second_of(q);
                    float o1, o2;
                    if(from_state)
                    {
                        const float2 stv = sst[q * 32];
                        pk::split(dbfs2(stv.x, stv.y, p.db_min), o1, o2);
                    }
                    else
                    {
                        o1 = prev_db[k1];
                        o2 = prev_db[k2];
                    }
                    if(!last_silent)
                    {
                        o1 = dbfs(o1, p.db_min);
                        o2 = dbfs(o2, p.db_min);
                        if(EXTRA)
                        {
                            if(p.normalize)
                            {
                                if(k1 + half >= 1)
                                    o1 += vc;
                                o2 += vc;
                            }
                            if(p.rolloff != nullptr)
                            {
                                if(k1 + half >= 1)
                                    o1 = fmaxf(o1 - __ldg(p.rolloff + k1 + half), p.db_min);
                                o2 = fmaxf(o2 - __ldg(p.rolloff + k2 + half), p.db_min);
                            }
                        }
                    }
                    outs &= !(o1 > p.floor_m10) & !(o2 > p.floor_m10);
                    if(k1 + half >= 1)
                        peak = fmaxf(peak, o1);
                    peak = fmaxf(peak, o2);
                    odb[k1] = o1;
                    odb[k2] = o2;
                }
                last_from_state = false;
            }
            if(gate && !last_silent)
                pos_w = __all_sync(0xffffffffu, outs);
            if(lane == 0)
                ctl[2 * (t & 1) + half] = pos_w ? 1 : 0;
            if(p.out_silent != nullptr && half == 0 && lane == 0)
                p.out_silent[(size_t)s * T + t] = last_silent ? 1 : 0;
            if(EXTRA && p.out_peak != nullptr)
            {
                const float gm = group_max<32>(peak, nullptr);
                if(lane == 0)
                    atomic_max_float(p.out_peak + t, gm);
            }
        }

        // ---- state back to the engine; m_decibels mirror (left implicit when it equals dbfs(state)); flags ----
        team::bar_sync(bar_id, 64); // the last tick's flag bytes of both warps are visible
        const bool pos_all = (ctl[2 * ((T - 1) & 1)] != 0) && (ctl[2 * ((T - 1) & 1) + 1] != 0);
        {
            const float *last = p.out_db + ((size_t)s * T + (T - 1)) * B + half;
            const bool plain = !EXTRA || (!p.normalize && p.rolloff == nullptr);
            const bool lazy = last_from_state && plain && (p.lazy_hold != 0);
#pragma unroll
            for(int q = 0; q < 16; ++q)
            {
                const int k1 = 2 * (lane + 32 * q) + half;
                const int k2 = 2 * second_of(q) + half;
                const float2 stv = sst[q * 32];
                state_s[k1] = stv.x;
                state_s[k2] = stv.y;
                if(p.write_hold && !lazy)
                {
                    if(last_from_state && plain)
                    {
                        float h1, h2;
                        pk::split(dbfs2(stv.x, stv.y, p.db_min), h1, h2);
                        hold_s[k1] = h1;
                        hold_s[k2] = h2;
                    }
                    else
                    {
                        hold_s[k1] = last[k1 - half];
                        hold_s[k2] = last[k2 - half];
                    }
                }
            }
            if(half == 0 && lane == 0)
                p.flags[s] = (unsigned char)((last_silent ? 1u : 0u) | (pos_all ? 2u : 0u) | 4u | (lazy ? 8u : 0u));
        }
        team::bar_sync(bar_id, 64); // the flag bytes are re-initialised for the team's next stream only after both have read them
    }
}

} // namespace wf
