// wf_v3_impl.cuh — launcher template shared by the two instantiation units (wf_v3_c1.cu / wf_v3_c2.cu)
#pragma once
#include <cuda_runtime.h>

#include "wf_v3.cuh"
#include "wf_v3.hpp"

namespace wf {
namespace v3impl {

template<int N, int CC, int R, int EXTRA>
cudaError_t launch_one(const KParams &kp, const v3::Tw3 &tw, cudaStream_t st, bool display, int device)
{
    const size_t smem = v3::smem_bytes<N>(kp.dch, kp.scratch_q, display, CC, R);
    static thread_local size_t configured[64] = {0};
    const int dev = device & 63;
    if(smem > 48 * 1024 && configured[dev] < smem)
    {
        cudaError_t err = cudaFuncSetAttribute(stft_v3_kernel<N, CC, R, EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem);
        if(err != cudaSuccess)
            return err;
        configured[dev] = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(kp.n_streams * R));
    cfg.blockDim = dim3((unsigned)v3::Geo3<N>::TN);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = R;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (R > 1) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, stft_v3_kernel<N, CC, R, EXTRA>, kp, tw);
}

template<int N, int CC, int EXTRA>
cudaError_t launch_r(int R, const KParams &kp, const v3::Tw3 &tw, cudaStream_t st, bool display, int device)
{
    switch(R)
    {
    case 1:
        if constexpr(N <= 8192)
            return launch_one<N, CC, 1, EXTRA>(kp, tw, st, display, device);
        else
            return cudaErrorInvalidValue;
    case 2: return launch_one<N, CC, 2, EXTRA>(kp, tw, st, display, device);
    case 4: return launch_one<N, CC, 4, EXTRA>(kp, tw, st, display, device);
    case 8: return launch_one<N, CC, 8, EXTRA>(kp, tw, st, display, device);
    default: return cudaErrorInvalidValue;
    }
}

template<int CC>
cudaError_t launch_cc(int N, int R, int extra, const KParams &kp, const v3::Tw3 &tw, cudaStream_t st, bool display, int device)
{
    switch(N)
    {
    case 1024:
        return (extra == 0)   ? launch_r<1024, CC, 0>(R, kp, tw, st, display, device)
               : (extra == 1) ? launch_r<1024, CC, 1>(R, kp, tw, st, display, device)
                              : launch_r<1024, CC, 3>(R, kp, tw, st, display, device);
    case 2048:
        return (extra == 0)   ? launch_r<2048, CC, 0>(R, kp, tw, st, display, device)
               : (extra == 1) ? launch_r<2048, CC, 1>(R, kp, tw, st, display, device)
                              : launch_r<2048, CC, 3>(R, kp, tw, st, display, device);
    case 4096:
        return (extra == 0)   ? launch_r<4096, CC, 0>(R, kp, tw, st, display, device)
               : (extra == 1) ? launch_r<4096, CC, 1>(R, kp, tw, st, display, device)
                              : launch_r<4096, CC, 3>(R, kp, tw, st, display, device);
    case 8192:
        return (extra == 0)   ? launch_r<8192, CC, 0>(R, kp, tw, st, display, device)
               : (extra == 1) ? launch_r<8192, CC, 1>(R, kp, tw, st, display, device)
                              : launch_r<8192, CC, 3>(R, kp, tw, st, display, device);
    case 16384:
        return (extra == 0)   ? launch_r<16384, CC, 0>(R, kp, tw, st, display, device)
               : (extra == 1) ? launch_r<16384, CC, 1>(R, kp, tw, st, display, device)
                              : launch_r<16384, CC, 3>(R, kp, tw, st, display, device);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace v3impl
} // namespace wf
