// Microbenchmark: issue throughput of scalar FFMA/FADD vs packed fma.rn.f32x2 / add.rn.f32x2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template<int MODE> __global__ void k(float *out, float a, float b)
{
    // 16 independent accumulator pairs per thread
    float x[32];
#pragma unroll
    for(int i = 0; i < 32; ++i) x[i] = a + i + threadIdx.x;
    unsigned long long w = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    unsigned long long c = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
    for(int it = 0; it < ITERS; ++it)
    {
        if(MODE == 0)
        {
#pragma unroll
            for(int i = 0; i < 32; ++i) x[i] = fmaf(x[i], b, a);
        }
        else if(MODE == 1)
        {
#pragma unroll
            for(int i = 0; i < 32; i += 2)
            {
                unsigned long long v = ((unsigned long long)__float_as_uint(x[i + 1]) << 32) | __float_as_uint(x[i]);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(w), "l"(c));
                x[i] = __uint_as_float((unsigned)v);
                x[i + 1] = __uint_as_float((unsigned)(v >> 32));
            }
        }
        else if(MODE == 2)
        {
#pragma unroll
            for(int i = 0; i < 32; ++i) x[i] = x[i] + b;
        }
        else if(MODE == 3)
        {
#pragma unroll
            for(int i = 0; i < 32; i += 2)
            {
                unsigned long long v = ((unsigned long long)__float_as_uint(x[i + 1]) << 32) | __float_as_uint(x[i]);
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
                x[i] = __uint_as_float((unsigned)v);
                x[i + 1] = __uint_as_float((unsigned)(v >> 32));
            }
        }
        else if(MODE == 4)
        { // mixed: 16 FFMA + 16 FADD scalar
#pragma unroll
            for(int i = 0; i < 16; ++i) x[i] = fmaf(x[i], b, a);
#pragma unroll
            for(int i = 16; i < 32; ++i) x[i] = x[i] + b;
        }
    }
    float s = 0;
#pragma unroll
    for(int i = 0; i < 32; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int MODE> void run(const char *name, float *d)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(d, 1.0001f, 0.9999f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(d, 1.0001f, 0.9999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = 148.0 * 4 * 256 * (double)ITERS * 32; // scalar-equivalent ops
    printf("%-28s %8.3f ms  %8.2f T scalar-op/s  (%.1f ops/clk/SM @1.9GHz)\n", name, ms, lane_ops / ms / 1e9, lane_ops / (ms*1e-3) / 148 / 1.9e9);
}
int main()
{
    float *d; cudaMalloc(&d, 148 * 4 * 256 * 4);
    run<0>("FFMA scalar", d); run<1>("fma.rn.f32x2", d); run<2>("FADD scalar", d); run<3>("add.rn.f32x2", d); run<4>("FFMA+FADD scalar mix", d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
