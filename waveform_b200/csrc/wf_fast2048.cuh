// wf_fast2048.cuh — hand-specialised sm_100a kernel for the headline shape: N = 2048, one capture
// channel, 16-byte aligned frames (hop % 4 == 0), spectrum output only.
//
// One WARP owns one stream and walks its frames; per frame:
//   * the 8 KB PCM frame is staged HBM -> shared memory by a TMA bulk copy (cp.async.bulk + mbarrier),
//     issued one frame ahead (during the previous frame's epilogue) so the warp never waits on DRAM;
//   * the packed 1024-point complex FFT is two radix-32 register passes (32 points per lane) with ONE
//     padded shared-memory transpose between them (conflict-free 64-bit accesses);
//   * the real-FFT split pass processes bins k and N/2-k TOGETHER (one twiddle multiply for two bins),
//     after a half-size exchange so that each lane owns both bins of its 16 pairs;
//   * |X| via MUFU.SQRT, EMA with state in registers across the stream's frames, dB via MUFU.LG2,
//     32 coalesced 128-byte store instructions per frame.
// Tables (window, inter-pass twiddles, split twiddles: 20 KB) live in shared memory once per CTA.
//
// Semantics are those of wf_kernels.cuh / src/source_generic.cpp:26-180; differences are confined to the
// last-ulp behaviour of sqrt/log (well inside the 1e-5 parity bar).
#pragma once
#include "wf_kernels.cuh"

namespace wf {

namespace fast {

constexpr int kN = 2048;
constexpr int kM = 1024;
constexpr int kWarpBufBytes = 32 * 33 * 8; // padded transpose buffer, also the TMA landing zone (8192 B used)
constexpr int kWarpBytes = kWarpBufBytes + 16; // + mbarrier
constexpr int kTableBytes = (1024 + 1024 + 512) * 8;
constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kSmemBytes = kTableBytes + kWarpsPerCta * kWarpBytes;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted on the mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// dbfs with the MUFU.LG2 path: 20 log10(m) = (20 log10 2) log2(m); __log2f keeps subnormal inputs exact-ish
__device__ __forceinline__ float dbfs_fast(float mag, float db_min)
{
    const float l = __log2f(mag) * 6.02059991327962390f;
    return (mag > 0.0f) ? l : db_min;
}

} // namespace fast

template<bool WIN, bool TSM, bool GATE, bool EXTRA>
__global__ void __launch_bounds__(fast::kThreads, 2) stft2048_fast_kernel(const __grid_constant__ KParams p)
{
    using namespace fast;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *s_win = reinterpret_cast<float2 *>(smem_raw);
    float2 *s_twA = s_win + 1024; // [k2][n1] = W_1024^(k2*n1)
    float2 *s_twP = s_twA + 1024; // [q][lane] = W_2048^(lane + 32 q), q < 16
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    unsigned char *wbase = reinterpret_cast<unsigned char *>(s_twP + 512) + warp * kWarpBytes;
    float2 *buf = reinterpret_cast<float2 *>(wbase);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(wbase + kWarpBufBytes);

    // ---- CTA prologue: tables -> shared, mbarriers ----
    for(int i = threadIdx.x; i < 1024; i += kThreads)
    {
        if(WIN)
            s_win[i] = __ldg(p.window2 + i);
        s_twA[i] = __ldg(p.tw + (((i >> 5) * (i & 31)) & 1023));
        if(i < 512)
            s_twP[i] = __ldg(p.tw_post + i);
    }
    if(lane == 0)
    {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int total_warps = gridDim.x * kWarpsPerCta;
    const int gw = blockIdx.x * kWarpsPerCta + warp;
    const int S = p.n_streams, T = p.n_frames;
    constexpr int B = kM;
    uint32_t phase = 0;

    // bins owned by this lane: first output of pair q -> k1 = lane + 32 q; second -> k2 = kb + 32 (31 - q)
    const int jp = (32 - lane) & 31;
    const int kb = jp + (lane == 0 ? 32 : 0);
    const int k2_q0 = (lane == 0) ? 512 : (kb + 992);
    const int pbase = kb; // partner column (+32 selects the extra row for lane 0)

    if(gw < S && lane == 0)
    {
        mbar_expect_tx(mbar, kN * 4);
        tma_load_1d(buf, p.pcm + (size_t)gw * p.stream_stride, kN * 4, mbar);
    }

    for(int s = gw; s < S; s += total_warps)
    {
        // ---- per-stream state ----
        float st1[16], st2[16];
        {
            const float *sp = p.state + (size_t)s * B;
#pragma unroll
            for(int q = 0; q < 16; ++q)
            {
                st1[q] = sp[lane + 32 * q];
                st2[q] = sp[(q == 0) ? k2_q0 : (kb + 32 * (31 - q))];
            }
        }
        const unsigned char fl = p.flags[s];
        bool last_silent = (fl & 1u) != 0;
        bool prev_out_silent = (fl & 2u) != 0;
        const float *pcm_s = p.pcm + (size_t)s * p.stream_stride;
        float *hold_s = p.hold_db + (size_t)s * B;

        for(int t = 0; t < T; ++t)
        {
            // ---- frame from shared (TMA-staged), window in the load prologue ----
            mbar_wait(mbar, phase);
            phase ^= 1u;
            float2 v[32];
            unsigned nzbits = 0;
#pragma unroll
            for(int pidx = 0; pidx < 32; ++pidx)
            {
                v[pidx] = buf[lane + 32 * pidx];
                nzbits |= __float_as_uint(v[pidx].x) | __float_as_uint(v[pidx].y);
            }
            if(WIN)
            {
#pragma unroll
                for(int pidx = 0; pidx < 32; ++pidx)
                {
                    const float2 w = s_win[lane + 32 * pidx];
                    v[pidx].x *= w.x;
                    v[pidx].y *= w.y;
                }
            }
            const bool nz = __any_sync(0xffffffffu, (nzbits & 0x7fffffffu) != 0u);

            // ---- pass A: 32-point DFTs over n2, inter-pass twiddle, transpose through shared ----
            dft_bitrev<32>(v);
            __syncwarp(); // every lane has read the frame before the buffer becomes the transpose area
#pragma unroll
            for(int k2 = 0; k2 < 32; ++k2)
            {
                float2 a = v[bitrev<32>(k2)];
                if(k2 > 0)
                    a = cmul(a, s_twA[k2 * 32 + lane]);
                buf[lane * 33 + k2] = a;
            }
            __syncwarp();
#pragma unroll
            for(int n1 = 0; n1 < 32; ++n1)
                v[n1] = buf[n1 * 33 + lane];
            // ---- pass B: 32-point DFTs over n1 -> X[lane + 32 k1] in v[bitrev(k1)] ----
            dft_bitrev<32>(v);

            // ---- split pass on pairs (k, 1024-k): upper half goes to the partner lane ----
            __syncwarp();
#pragma unroll
            for(int q = 16; q < 32; ++q)
                buf[(q - 16) * 32 + lane] = v[bitrev<32>(q)];
            if(lane == 0)
                buf[16 * 32] = v[0]; // X[0] doubles as "X[1024]" for the k = 0 pair
            __syncwarp();
            float mag1[16], mag2[16];
#pragma unroll
            for(int q = 0; q < 16; ++q)
            {
                const float2 a = v[bitrev<32>(q)];
                float2 b = buf[(15 - q) * 32 + pbase];
                b.y = -b.y;
                const float2 sum = cadd(a, b);
                const float2 dif = csub(a, b);
                const float2 o = make_float2(dif.y, -dif.x);
                const float2 wo = cmul(o, s_twP[q * 32 + lane]);
                const float2 y1 = cadd(sum, wo);
                const float2 y2 = csub(sum, wo);
                float p1 = fmaf(y1.x, y1.x, y1.y * y1.y);
                float p2 = fmaf(y2.x, y2.x, y2.y * y2.y);
                if(q == 0)
                {
                    // lane 0: the pair (0, 1024) has no bin 1024; its second slot carries bin 512 = conj(X[512])
                    const float2 x512 = v[bitrev<32>(16)];
                    const float p512 = 4.0f * fmaf(x512.x, x512.x, x512.y * x512.y);
                    p2 = (lane == 0) ? p512 : p2;
                }
                mag1[q] = sqrt_approx(p1) * p.coef_half;
                mag2[q] = sqrt_approx(p2) * p.coef_half;
            }
            __syncwarp(); // all generic-proxy accesses to buf are done

            // ---- prefetch the next frame (or the next stream's first frame) while the epilogue runs ----
            if(lane == 0)
            {
                const float *next = nullptr;
                if(t + 1 < T)
                    next = pcm_s + (size_t)(t + 1) * p.hop;
                else if(s + total_warps < S)
                    next = p.pcm + (size_t)(s + total_warps) * p.stream_stride;
                if(next != nullptr)
                {
                    fence_proxy_async();
                    mbar_expect_tx(mbar, kN * 4);
                    tma_load_1d(buf, next, kN * 4, mbar);
                }
            }

            // ---- gate (src/source_generic.cpp:63-95), single capture channel ----
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
                        last_silent = true; // ++silent_channels >= 1
                        do_proc = false;
                    }
                }
            }

            // ---- slope, EMA (src/source_generic.cpp:121-134) ----
            if(do_proc)
            {
#pragma unroll
                for(int q = 0; q < 16; ++q)
                {
                    float m1 = mag1[q], m2 = mag2[q];
                    if(EXTRA && p.slope != nullptr)
                    {
                        m1 *= __ldg(p.slope + lane + 32 * q);
                        m2 *= __ldg(p.slope + ((q == 0) ? k2_q0 : (kb + 32 * (31 - q))));
                    }
                    if(TSM)
                    {
                        float o1 = st1[q], o2 = st2[q];
                        if(EXTRA && p.fast_peaks)
                        {
                            o1 = fmaxf(m1, o1);
                            o2 = fmaxf(m2, o2);
                        }
                        m1 = __fadd_rn(__fmul_rn(p.g, o1), __fmul_rn(p.g2, m1));
                        m2 = __fadd_rn(__fmul_rn(p.g, o2), __fmul_rn(p.g2, m2));
                    }
                    st1[q] = m1;
                    st2[q] = m2;
                }
            }

            // ---- dBFS, volume normalisation, roll-off, store (src/source_generic.cpp:138-179) ----
            float vc = 0.0f;
            if(EXTRA && p.normalize)
            {
                const float rms = (p.input_rms != nullptr) ? p.input_rms[(size_t)s * T + t] : 0.0f;
                vc = fminf(p.vol_target - dbfs(rms, p.db_min), p.max_gain);
            }
            float *odb = p.out_db + ((size_t)s * T + t) * B;
            const float *prev_db = (t > 0) ? (odb - B) : hold_s;
            bool outs = true;
            float peak = -INFINITY;
#pragma unroll
            for(int q = 0; q < 16; ++q)
            {
                const int k1 = lane + 32 * q;
                const int k2 = (q == 0) ? k2_q0 : (kb + 32 * (31 - q));
                float o1, o2;
                if(last_silent)
                {
                    o1 = prev_db[k1];
                    o2 = prev_db[k2];
                }
                else
                {
                    const float in1 = do_proc ? st1[q] : prev_db[k1];
                    const float in2 = do_proc ? st2[q] : prev_db[k2];
                    o1 = dbfs_fast(in1, p.db_min);
                    o2 = dbfs_fast(in2, p.db_min);
                    if(EXTRA)
                    {
                        if(p.normalize)
                        {
                            if(k1 >= 1)
                                o1 += vc;
                            o2 += vc;
                        }
                        if(p.rolloff != nullptr)
                        {
                            if(k1 >= 1)
                                o1 = fmaxf(o1 - __ldg(p.rolloff + k1), p.db_min);
                            o2 = fmaxf(o2 - __ldg(p.rolloff + k2), p.db_min);
                        }
                    }
                }
                if(GATE)
                    outs &= !(o1 > p.floor_m10) & !(o2 > p.floor_m10);
                if(EXTRA)
                {
                    if(k1 >= 1)
                        peak = fmaxf(peak, o1);
                    peak = fmaxf(peak, o2);
                }
                stg_stream(odb + k1, o1);
                stg_stream(odb + k2, o2);
            }
            if(GATE && !last_silent)
                prev_out_silent = __all_sync(0xffffffffu, outs);
            if(p.out_silent != nullptr && lane == 0)
                p.out_silent[(size_t)s * T + t] = last_silent ? 1 : 0;
            if(EXTRA)
            {
                if(p.out_peak != nullptr)
                {
                    const float gm = group_max<32>(peak, nullptr);
                    if(lane == 0)
                        atomic_max_float(p.out_peak + t, gm);
                }
            }
        }

        // ---- state back to the engine ----
        {
            float *sp = p.state + (size_t)s * B;
            const float *last = p.out_db + ((size_t)s * T + (T - 1)) * B;
#pragma unroll
            for(int q = 0; q < 16; ++q)
            {
                const int k1 = lane + 32 * q;
                const int k2 = (q == 0) ? k2_q0 : (kb + 32 * (31 - q));
                sp[k1] = st1[q];
                sp[k2] = st2[q];
                if(p.write_hold)
                {
                    hold_s[k1] = last[k1];
                    hold_s[k2] = last[k2];
                }
            }
            if(lane == 0)
                p.flags[s] = (unsigned char)((last_silent ? 1u : 0u) | (prev_out_silent ? 2u : 0u) | 4u);
        }
    }
}

} // namespace wf
