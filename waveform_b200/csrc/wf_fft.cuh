// wf_fft.cuh — register-resident DFT building blocks for the sm_100a STFT kernels.
//
// A thread holds R complex points in registers (R = 2..32) and performs a radix-2 decimation-in-
// frequency network on them with compile-time twiddles; the result of bin k ends up in register
// bitrev(k), which callers absorb into their (compile-time) store indices, so no data movement is
// spent on reordering.  These are the "warp-shuffle / register butterflies" of the design: all
// inter-thread exchange happens once per Stockham pass through padded shared memory.
//
// Replaces (together with wf_kernels.cuh) the FFTW codelets the reference executes through
// fftwf_execute (src/source_generic.cpp:106; deps/fftw-3.3.11/rdft/ct-hc2c.c:59-70).
#pragma once
#include <cuda_runtime.h>

namespace wf {

// ---- compile-time sin/cos (double precision Taylor with octant reduction) -------------------------
namespace cx {
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double sin_taylor(double x)
{ // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for(int n = 1; n < 12; ++n)
    {
        term *= -x2 / ((2.0 * n) * (2.0 * n + 1.0));
        sum += term;
    }
    return sum;
}
constexpr double cos_taylor(double x)
{ // |x| <= pi/4
    double x2 = x * x, term = 1.0, sum = 1.0;
    for(int n = 1; n < 12; ++n)
    {
        term *= -x2 / ((2.0 * n - 1.0) * (2.0 * n));
        sum += term;
    }
    return sum;
}
// cos(2 pi k / n), sin(2 pi k / n) for 0 <= k < n, exact symmetry handling by octant
constexpr double cos2pi(int k, int n)
{
    k %= n;
    if(8 * k <= n) return cos_taylor(2.0 * kPi * k / n);
    if(8 * k <= 3 * n) return -sin_taylor(2.0 * kPi * k / n - kPi / 2);
    if(8 * k <= 5 * n) return -cos_taylor(2.0 * kPi * k / n - kPi);
    if(8 * k <= 7 * n) return sin_taylor(2.0 * kPi * k / n - 3 * kPi / 2);
    return cos_taylor(2.0 * kPi * k / n - 2 * kPi);
}
constexpr double sin2pi(int k, int n)
{
    k %= n;
    if(8 * k <= n) return sin_taylor(2.0 * kPi * k / n);
    if(8 * k <= 3 * n) return cos_taylor(2.0 * kPi * k / n - kPi / 2);
    if(8 * k <= 5 * n) return -sin_taylor(2.0 * kPi * k / n - kPi);
    if(8 * k <= 7 * n) return -cos_taylor(2.0 * kPi * k / n - 3 * kPi / 2);
    return sin_taylor(2.0 * kPi * k / n - 2 * kPi);
}
} // namespace cx

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
    return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}

// multiply by the compile-time twiddle W_L^J = exp(-2 pi i J / L)
template<int J, int L>
__device__ __forceinline__ float2 mul_tw(float2 a)
{
    if constexpr(J == 0)
        return a;
    else if constexpr(4 * J == L) // -i
        return make_float2(a.y, -a.x);
    else if constexpr(8 * J == L) // (1 - i)/sqrt2
    {
        constexpr float h = 0.70710678118654752440f;
        return make_float2((a.x + a.y) * h, (a.y - a.x) * h);
    }
    else if constexpr(8 * J == 3 * L) // (-1 - i)/sqrt2
    {
        constexpr float h = 0.70710678118654752440f;
        return make_float2((a.y - a.x) * h, -(a.x + a.y) * h);
    }
    else
    {
        constexpr float c = (float)cx::cos2pi(J, L);
        constexpr float s = -(float)cx::sin2pi(J, L);
        return make_float2(fmaf(a.x, c, -a.y * s), fmaf(a.x, s, a.y * c));
    }
}

template<int R>
__host__ __device__ constexpr int bitrev(int k)
{
    int r = 0;
    for(int b = 1; b < R; b <<= 1)
    {
        r = (r << 1) | (k & 1);
        k >>= 1;
    }
    return r;
}

// One DIF stage on the sub-block [BASE, BASE+L): v[b+j] = a+c, v[b+j+L/2] = (a-c) W_L^j
template<int L, int BASE, int J, typename V>
__device__ __forceinline__ void dif_bfly(V &v)
{
    if constexpr(J < L / 2)
    {
        const float2 a = v[BASE + J];
        const float2 c = v[BASE + J + L / 2];
        v[BASE + J] = cadd(a, c);
        v[BASE + J + L / 2] = mul_tw<J, L>(csub(a, c));
        dif_bfly<L, BASE, J + 1>(v);
    }
}

template<int L, int BASE, typename V>
__device__ __forceinline__ void dif_block(V &v)
{
    if constexpr(L >= 2)
    {
        dif_bfly<L, BASE, 0>(v);
        dif_block<L / 2, BASE>(v);
        dif_block<L / 2, BASE + L / 2>(v);
    }
}

// In-place forward DFT of v[0..R): afterwards X[k] == v[bitrev<R>(k)].
template<int R>
__device__ __forceinline__ void dft_bitrev(float2 (&v)[R])
{
    dif_block<R, 0>(v);
}

} // namespace wf
