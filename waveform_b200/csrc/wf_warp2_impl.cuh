// wf_warp2_impl.cuh — launcher template shared by the wf_warp2_*.cu instantiation units
#pragma once
#include "wf_warp2.cuh"
#include "wf_warp2.hpp"

namespace wf {
namespace warp2 {

template<int L, int P, bool EXTRA>
cudaError_t launch(const KParams &kp, int grid, int warps, cudaStream_t st, bool pdl, int device)
{
    using G = Geo<L, P>;
    static thread_local bool configured[64] = {false};
    const int dev = device & 63;
    if(!configured[dev])
    {
        cudaError_t err = cudaFuncSetAttribute(stft_warp2_kernel<L, P, EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               G::smem_bytes(G::kWarps));
        if(err != cudaSuccess)
            return err;
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)(warps * 32));
    cfg.dynamicSmemBytes = G::smem_bytes(warps);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, stft_warp2_kernel<L, P, EXTRA>, kp);
}

#define WF_WARP2_CASE(NN, LL, PP_)                                                                                       \
    case NN:                                                                                                             \
        static_assert(2 * LL * PP_ == NN, "plan");                                                                      \
        *name = "stft_warp2_kernel<" #LL "," #PP_ ">";                                                                   \
        return extra ? launch<LL, PP_, true>(kp, grid, warps, st, pdl, device) : launch<LL, PP_, false>(kp, grid, warps, st, pdl, device);

} // namespace warp2
} // namespace wf
