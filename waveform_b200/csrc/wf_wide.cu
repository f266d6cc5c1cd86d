// wf_wide.cu — instantiations + launcher of the cluster kernel (wf_wide.cuh); a separate translation unit so that it
// compiles in parallel with wf_engine.cu.
#include <cuda_runtime.h>

#include "wf_wide.cuh"
#include "wf_wide.hpp"

namespace wf {

namespace {

template<int N, int CC, int R>
cudaError_t launch_one(const KParams &kp, cudaStream_t st, bool display, int device)
{
    const size_t smem = wide::smem_bytes<N>(kp.dch, kp.scratch_q, display);
    static thread_local size_t configured[64] = {0};
    const int dev = device & 63;
    if(smem > 48 * 1024 && configured[dev] < smem)
    {
        cudaError_t err =
            cudaFuncSetAttribute(stft_wide_kernel<N, CC, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if(err != cudaSuccess)
            return err;
        configured[dev] = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(kp.n_streams * R));
    cfg.blockDim = dim3((unsigned)Geo<N>::TN);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = R;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, stft_wide_kernel<N, CC, R>, kp);
}

template<int N, int CC>
cudaError_t launch_r(int R, const KParams &kp, cudaStream_t st, bool display, int device)
{
    switch(R)
    {
    case 2: return launch_one<N, CC, 2>(kp, st, display, device);
    case 4: return launch_one<N, CC, 4>(kp, st, display, device);
    case 8: return launch_one<N, CC, 8>(kp, st, display, device);
    default: return cudaErrorInvalidValue;
    }
}

template<int CC>
cudaError_t launch_n(int N, int R, const KParams &kp, cudaStream_t st, bool display, int device)
{
    switch(N)
    {
    case 4096: return launch_r<4096, CC>(R, kp, st, display, device);
    case 8192: return launch_r<8192, CC>(R, kp, st, display, device);
    case 16384: return launch_r<16384, CC>(R, kp, st, display, device);
    case 32768: return launch_r<32768, CC>(R, kp, st, display, device);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace

bool wide_supported(int N) { return N == 4096 || N == 8192 || N == 16384 || N == 32768; }

size_t wide_smem_bytes(int N, int dch, int n_points, bool display)
{
    switch(N)
    {
    case 4096: return wide::smem_bytes<4096>(dch, n_points, display);
    case 8192: return wide::smem_bytes<8192>(dch, n_points, display);
    case 16384: return wide::smem_bytes<16384>(dch, n_points, display);
    case 32768: return wide::smem_bytes<32768>(dch, n_points, display);
    default: return 0;
    }
}

cudaError_t wide_launch(int N, int cc, int R, const KParams &kp, cudaStream_t st, bool display, int device)
{
    return (cc == 2) ? launch_n<2>(N, R, kp, st, display, device) : launch_n<1>(N, R, kp, st, display, device);
}

} // namespace wf
