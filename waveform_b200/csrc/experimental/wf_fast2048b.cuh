// EXPERIMENT (not built by default; -DWF_BUILD_EXPERIMENTAL): measured 317 M spectra/s vs 381 M for wf_fast2048.cuh on
// B200 — the second shared-memory exchange makes it MIO/shared-memory bound (LDS+STS 342 vs 208 per frame,
// mio_throttle 3.5/issue; profiles/r01h_pair2048_first.txt).  Kept as the record of that design point.
//
// wf_fast2048b.cuh — second-generation kernel for the headline shape (N = 2048, one capture channel, aligned frames,
// spectrum output).  Same pipeline and semantics as wf_fast2048.cuh, different parallel decomposition:
//
//   * TWO warps (64 threads) own a stream and walk its frames; each thread holds only 16 complex points, so the
//     kernel needs ~70 registers instead of ~126 and an SM keeps up to 28 warps (14 frames) in flight instead of 16.
//     The kernel is latency-bound, not pipe-bound (profiles/r01_fast2048_final.txt: issue 53 %, FMA pipe 48 %), so
//     the extra warps are what buys throughput.
//   * packed 1024-point complex FFT = Stockham radix 16 x 16 x 4 (two shared-memory exchanges, 17-stride padding),
//     all complex arithmetic packed f32x2 as before; frame staged by TMA (cp.async.bulk + mbarrier) one frame ahead.
//   * split pass on pairs (k, 1024-k); the partner half crosses the two warps through shared memory.
//   * the 64 threads of a group synchronise with a named barrier (bar.sync id, 64).
#pragma once
#include "../wf_fast2048.cuh"

namespace wf {
namespace fastb {

constexpr int kGroupThreads = 64;
constexpr int kMaxGroups = 14;
constexpr int kBufElems = 1024 + 64;                      // padded: phys(i) = i + (i >> 4)
constexpr int kGroupBytes = kBufElems * 8 + 4096 + 32;     // exchange buffer + EMA state + mbarrier/flags
constexpr int kTableElems = 1024 + 256 + 768 + 512;        // window, pass-1 twiddles, pass-2 twiddles, split twiddles
constexpr int kTableBytes = kTableElems * 8;
constexpr int smem_bytes(int groups) { return kTableBytes + groups * kGroupBytes; }

__device__ __forceinline__ void group_bar(int id)
{
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

} // namespace fastb

template<bool TSM, bool GATE, bool EXTRA>
__global__ void __launch_bounds__(fastb::kMaxGroups * 64, 1) stft2048_pair_kernel(const __grid_constant__ KParams p)
{
    using namespace fast;
    using namespace fastb;
    using pk::c64;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    c64 *s_win = reinterpret_cast<c64 *>(smem_raw);   // [1024] window pairs x normalisation
    c64 *s_tw1 = s_win + 1024;                        // [t][g & 15]   = W_1024^(4 (g&15) t)
    c64 *s_tw2 = s_tw1 + 256;                         // [t-1][j]      = W_1024^(j t), j < 256, t = 1..3
    c64 *s_twP = s_tw2 + 768;                         // [k]           = W_2048^k, k < 512
    const int grp = threadIdx.x >> 6;
    const int g = threadIdx.x & 63;
    const int n_groups = blockDim.x >> 6;
    unsigned char *gbase = reinterpret_cast<unsigned char *>(s_twP + 512) + grp * kGroupBytes;
    c64 *buf = reinterpret_cast<c64 *>(gbase);
    c64 *sst = reinterpret_cast<c64 *>(gbase + kBufElems * 8) + g; // state column of this thread: [c][g]
    uint64_t *mbar = reinterpret_cast<uint64_t *>(gbase + kBufElems * 8 + 4096);
    volatile int *gflags = reinterpret_cast<volatile int *>(gbase + kBufElems * 8 + 4096 + 16); // [0..1] nz, [2..3] outs
    const int bar_id = 1 + grp;

    // ---- CTA prologue: tables -> shared, mbarriers ----
    for(int i = threadIdx.x; i < 1024; i += blockDim.x)
    {
        const float2 w = (p.window2 != nullptr) ? __ldg(p.window2 + i) : make_float2(1.0f, 1.0f);
        s_win[i] = pk::make(w.x * p.coef_half, w.y * p.coef_half);
        if(i < 256)
            s_tw1[i] = pk::from(__ldg(p.tw + ((4 * (i & 15) * (i >> 4)) & 1023)));
        if(i < 768)
            s_tw2[i] = pk::from(__ldg(p.tw + (((i & 255) * ((i >> 8) + 1)) & 1023)));
        if(i < 512)
            s_twP[i] = pk::from(__ldg(p.tw_post + i));
    }
    if(g == 0)
    {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int S = p.n_streams, T = p.n_frames;
    const int G = gridDim.x;
    const int n_local = (S > (int)blockIdx.x) ? (S - (int)blockIdx.x + G - 1) / G : 0;
    constexpr int B = kM;
    uint32_t phase = 0;

    // bins of pair c (c = 0..7): k1 = g + 64 (c & 3) + 256 (c >> 2);  k2 = 1024 - k1  (g == 0, c == 0: bin 512)
    const int pg = (g == 0) ? 64 : (64 - g);      // partner column in the exchange area (+64 = next row for thread 0)
    const int k2base = (g == 0) ? 1024 : (1024 - g);

    if(grp < n_local && g == 0)
    {
        mbar_expect_tx(mbar, kN * 4);
        tma_load_1d(buf, p.pcm + (size_t)(blockIdx.x + grp * G) * p.stream_stride, kN * 4, mbar);
    }

    for(int li = grp; li < n_local; li += n_groups)
    {
        const int s = (int)blockIdx.x + li * G;
        // ---- per-stream state: global (natural bin order) -> shared ([pair][thread]) ----
        {
            const float *sp = p.state + (size_t)s * B;
#pragma unroll
            for(int c = 0; c < 8; ++c)
            {
                const int k1 = g + 64 * (c & 3) + 256 * (c >> 2);
                const int k2 = (g == 0 && c == 0) ? 512 : (k2base - 64 * (c & 3) - 256 * (c >> 2));
                sst[c * 64] = pk::make(sp[k1], sp[k2]);
            }
        }
        const unsigned char fl = p.flags[s];
        bool last_silent = (fl & 1u) != 0;
        bool prev_out_silent = (fl & 2u) != 0;
        bool last_from_state = false;
        const float *pcm_s = p.pcm + (size_t)s * p.stream_stride;
        float *hold_s = p.hold_db + (size_t)s * B;

#pragma unroll 1
        for(int t = 0; t < T; ++t)
        {
            // ---- pass 0: frame (TMA-staged) x window, radix 16 over n2 (n = g + 64 n2) ----
            mbar_wait(mbar, phase);
            phase ^= 1u;
            c64 v[16];
            unsigned long long nzbits = 0;
#pragma unroll
            for(int i = 0; i < 16; ++i)
            {
                v[i] = buf[g + 64 * i];
                nzbits |= v[i];
            }
#pragma unroll
            for(int i = 0; i < 16; ++i)
                v[i] = pk::mul(v[i], s_win[g + 64 * i]);
            const bool nzw = __any_sync(0xffffffffu, (nzbits & 0x7fffffff7fffffffull) != 0ull);
            if((g & 31) == 0)
                gflags[g >> 5] = nzw ? 1 : 0;
            pk::dft_bitrev<16>(v);
            group_bar(bar_id); // everyone has read the frame; nz flags visible
            const bool nz = (gflags[0] | gflags[1]) != 0;
#pragma unroll
            for(int i = 0; i < 16; ++i)
                buf[17 * g + i] = v[bitrev<16>(i)];
            group_bar(bar_id);

            // ---- pass 1: radix 16, Ns = 16 ----
            {
                const int base = g + (g >> 4);
#pragma unroll
                for(int i = 0; i < 16; ++i)
                    v[i] = buf[base + 68 * i];
                group_bar(bar_id); // all reads done before the in-place stores
#pragma unroll
                for(int i = 1; i < 16; ++i)
                    v[i] = pk::cmul(v[i], s_tw1[i * 16 + (g & 15)]);
                pk::dft_bitrev<16>(v);
                const int ob = (g >> 4) * 272 + (g & 15);
#pragma unroll
                for(int i = 0; i < 16; ++i)
                    buf[ob + 17 * i] = v[bitrev<16>(i)];
                group_bar(bar_id);
            }

            // ---- pass 2: radix 4, Ns = 256: X[g + 64 b + 256 t2] ends up in v[b*4 + bitrev4(t2)] ----
            {
                const int base = g + (g >> 4);
#pragma unroll
                for(int b = 0; b < 4; ++b)
#pragma unroll
                    for(int i = 0; i < 4; ++i)
                        v[b * 4 + i] = buf[base + 68 * b + 272 * i];
                group_bar(bar_id); // all reads done: the buffer becomes the split-pass exchange area
#pragma unroll
                for(int b = 0; b < 4; ++b)
                {
#pragma unroll
                    for(int i = 1; i < 4; ++i)
                        v[b * 4 + i] = pk::cmul(v[b * 4 + i], s_tw2[(i - 1) * 256 + g + 64 * b]);
                    c64 x[4] = {v[b * 4 + 0], v[b * 4 + 1], v[b * 4 + 2], v[b * 4 + 3]};
                    pk::dft_bitrev<4>(x);
#pragma unroll
                    for(int i = 0; i < 4; ++i)
                        v[b * 4 + i] = x[i];
                }
            }
            // value of pair index c = b + 4 t2:  X_c = v[(c & 3) * 4 + bitrev4(c >> 2)]
#define WF_XC(c) v[((c) & 3) * 4 + bitrev<4>((c) >> 2)]

            // ---- split pass exchange: upper half (c >= 8) to the partner thread (64 - g) through shared memory ----
#pragma unroll
            for(int c = 8; c < 16; ++c)
                buf[(c - 8) * 64 + g] = WF_XC(c);
            if(g == 0)
                buf[8 * 64] = WF_XC(0); // X[0] doubles as "X[1024]" for the k = 0 pair
            group_bar(bar_id);
            c64 part[8];
#pragma unroll
            for(int c = 0; c < 8; ++c)
                part[c] = buf[(7 - c) * 64 + pg];
            group_bar(bar_id); // all generic-proxy accesses to buf are done: it can take the next frame

            // ---- prefetch the next frame (or the next stream's first frame) under the epilogue ----
            if(g == 0)
            {
                const float *next = nullptr;
                if(t + 1 < T)
                    next = pcm_s + (size_t)(t + 1) * p.hop;
                else if(li + n_groups < n_local)
                    next = p.pcm + (size_t)(s + n_groups * G) * p.stream_stride;
                if(next != nullptr)
                {
                    fence_proxy_async();
                    mbar_expect_tx(mbar, kN * 4);
                    tma_load_1d(buf, next, kN * 4, mbar);
                }
            }

            // ---- gate (src/source_generic.cpp:63-95), single capture channel ----
            if(GATE && t > 0 && !last_silent) // outs flags of the previous tick were published before its last barrier
                prev_out_silent = (gflags[2] & gflags[3]) != 0;
            const bool skip_all = EXTRA && (p.skip_mask != nullptr) && (p.skip_mask[(size_t)s * T + t] != 0);
            bool do_proc = !skip_all;
            if(!skip_all)
            {
                if(nz)
                    last_silent = false;
                else if(GATE)
                {
                    if(last_silent)
                        do_proc = false;
                    else if(prev_out_silent)
                    {
                        last_silent = true;
                        do_proc = false;
                    }
                }
            }

            float *odb = p.out_db + ((size_t)s * T + t) * B;
            float vc = 0.0f;
            if(EXTRA && p.normalize)
            {
                const float rms = (p.input_rms != nullptr) ? p.input_rms[(size_t)s * T + t] : 0.0f;
                vc = fminf(p.vol_target - dbfs(rms, p.db_min), p.max_gain);
            }
            bool outs = true;
            float peak = -INFINITY;

            if(do_proc && !last_silent)
            {
#pragma unroll
                for(int c = 0; c < 8; ++c)
                {
                    const int k1 = g + 64 * (c & 3) + 256 * (c >> 2);
                    const int k2 = (c == 0) ? ((g == 0) ? 512 : k2base) : (k2base - 64 * (c & 3) - 256 * (c >> 2));
                    const c64 a = WF_XC(c);
                    const c64 b = pk::conj(part[c]);
                    const c64 sum = pk::add(a, b);
                    const c64 o = pk::mul_neg_i(pk::sub(a, b));
                    const c64 wo = pk::cmul(o, s_twP[k1]);
                    const c64 y1 = pk::add(sum, wo);
                    const c64 y2 = pk::sub(sum, wo);
                    const c64 s1 = pk::mul(y1, y1), s2 = pk::mul(y2, y2);
                    float p1 = pk::re(s1) + pk::im(s1);
                    float p2 = pk::re(s2) + pk::im(s2);
                    if(c == 0)
                    {
                        const c64 x512 = WF_XC(8); // thread 0: bin 512 rides in the unused (0, 1024) slot
                        const c64 sq = pk::mul(x512, x512);
                        const float p512 = 4.0f * (pk::re(sq) + pk::im(sq));
                        p2 = (g == 0) ? p512 : p2;
                    }
                    c64 m = pk::make(sqrt_approx(p1), sqrt_approx(p2));
                    if(EXTRA && p.slope != nullptr)
                        m = pk::mul(m, pk::make(__ldg(p.slope + k1), __ldg(p.slope + k2)));
                    if(TSM)
                    {
                        c64 old = sst[c * 64];
                        if(EXTRA && p.fast_peaks)
                            old = pk::make(fmaxf(pk::re(m), pk::re(old)), fmaxf(pk::im(m), pk::im(old)));
                        m = pk::fma(pk::make(p.g, p.g), old, pk::mul(pk::make(p.g2, p.g2), m));
                    }
                    sst[c * 64] = m;
                    float d1, d2;
                    pk::split(dbfs2(pk::re(m), pk::im(m), p.db_min), d1, d2);
                    if(EXTRA)
                    {
                        if(p.normalize)
                        {
                            if(k1 >= 1)
                                d1 += vc;
                            d2 += vc;
                        }
                        if(p.rolloff != nullptr)
                        {
                            if(k1 >= 1)
                                d1 = fmaxf(d1 - __ldg(p.rolloff + k1), p.db_min);
                            d2 = fmaxf(d2 - __ldg(p.rolloff + k2), p.db_min);
                        }
                        if(k1 >= 1)
                            peak = fmaxf(peak, d1);
                        peak = fmaxf(peak, d2);
                    }
                    if(GATE)
                        outs &= !(d1 > p.floor_m10) & !(d2 > p.floor_m10);
                    stg_stream(odb + k1, d1);
                    stg_stream(odb + k2, d2);
                }
                last_from_state = true;
            }
            else
            {
                // rare path: hold (tick returned early) or stale-dB quirk (skipped channel), see wf_fast2048.cuh
                const float *prev_db = (t > 0) ? (odb - B) : hold_s;
#pragma unroll 1
                for(int k = g; k < B; k += 64)
                {
                    float o = prev_db[k];
                    if(!last_silent)
                    {
                        o = dbfs(o, p.db_min);
                        if(EXTRA && k >= 1)
                        {
                            if(p.normalize)
                                o += vc;
                            if(p.rolloff != nullptr)
                                o = fmaxf(o - __ldg(p.rolloff + k), p.db_min);
                        }
                    }
                    outs &= !(o > p.floor_m10);
                    if(k >= 1)
                        peak = fmaxf(peak, o);
                    odb[k] = o;
                }
                last_from_state = false;
            }
            if(GATE && !last_silent)
            {
                const bool ow = __all_sync(0xffffffffu, outs);
                if((g & 31) == 0)
                    gflags[2 + (g >> 5)] = ow ? 1 : 0; // read after the next tick's barriers (or the stream epilogue's)
            }
            if(p.out_silent != nullptr && g == 0)
                p.out_silent[(size_t)s * T + t] = last_silent ? 1 : 0;
            if(EXTRA)
            {
                if(p.out_peak != nullptr)
                {
                    const float gm = group_max<32>(peak, nullptr);
                    if((g & 31) == 0)
                        atomic_max_float(p.out_peak + t, gm);
                }
            }
#undef WF_XC
        }

        // ---- state back to the engine; m_decibels mirror for the next call's gate / hold paths ----
        group_bar(bar_id);
        if(GATE && T > 0 && !last_silent)
            prev_out_silent = (gflags[2] & gflags[3]) != 0;
        {
            float *sp = p.state + (size_t)s * B;
            const float *last = p.out_db + ((size_t)s * T + (T - 1)) * B;
            const bool plain = !EXTRA || (!p.normalize && p.rolloff == nullptr);
#pragma unroll
            for(int c = 0; c < 8; ++c)
            {
                const int k1 = g + 64 * (c & 3) + 256 * (c >> 2);
                const int k2 = (g == 0 && c == 0) ? 512 : (k2base - 64 * (c & 3) - 256 * (c >> 2));
                float s1v, s2v;
                pk::split(sst[c * 64], s1v, s2v);
                sp[k1] = s1v;
                sp[k2] = s2v;
                if(p.write_hold)
                {
                    if(last_from_state && plain)
                    {
                        float h1, h2;
                        pk::split(dbfs2(s1v, s2v, p.db_min), h1, h2);
                        hold_s[k1] = h1;
                        hold_s[k2] = h2;
                    }
                    else
                    {
                        hold_s[k1] = last[k1];
                        hold_s[k2] = last[k2];
                    }
                }
            }
            if(g == 0)
                p.flags[s] = (unsigned char)((last_silent ? 1u : 0u) | (prev_out_silent ? 2u : 0u) | 4u);
        }
        group_bar(bar_id);
    }
}

} // namespace wf
