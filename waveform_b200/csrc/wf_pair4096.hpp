// wf_pair4096.hpp — host interface of the two-warps-per-stream N=4096 kernel (wf_pair4096.cuh)
#pragma once
#include <cuda_runtime.h>

namespace wf {
struct KParams;
// grid = CTAs (one per SM at most, 8 pairs each); extra = slope / fast peaks / skip mask / volume / roll-off / peak output /
// per-tick gravity in use
cudaError_t pair4096_launch(bool extra, const KParams &kp, int grid, cudaStream_t st, bool pdl, int device);
} // namespace wf
