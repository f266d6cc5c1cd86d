// wf_pair4096.cu — instantiations + launcher of stft4096_pair_kernel (its own translation unit)
#include "wf_team2048.cuh"
#include "wf_pair4096.cuh"
#include "wf_pair4096.hpp"

namespace wf {

template<bool EXTRA>
static cudaError_t launch(const KParams &kp, int grid, cudaStream_t st, bool pdl, int device)
{
    static thread_local bool configured[64] = {false};
    const int dev = device & 63;
    if(!configured[dev])
    {
        cudaError_t err = cudaFuncSetAttribute(stft4096_pair_kernel<EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               pair4096::smem_bytes());
        if(err != cudaSuccess)
            return err;
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(pair4096::kPWarps * 32);
    cfg.dynamicSmemBytes = pair4096::smem_bytes();
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, stft4096_pair_kernel<EXTRA>, kp);
}

cudaError_t pair4096_launch(bool extra, const KParams &kp, int grid, cudaStream_t st, bool pdl, int device)
{
    return extra ? launch<true>(kp, grid, st, pdl, device) : launch<false>(kp, grid, st, pdl, device);
}

} // namespace wf
