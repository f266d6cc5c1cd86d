// wf_warp2.hpp — host interface of the warp-per-stream kernel for fft sizes N = 2*L*P (wf_warp2.cuh)
#pragma once
#include <cuda_runtime.h>

namespace wf {
struct KParams;
// true when a compiled (L, P) plan exists for this fft size
bool warp2_supported(int N);
// power-of-two sizes with a plan (routed here only for display outputs)
bool warp2_pow2_supported(int N);
// *warps = warps per CTA (1..16; lowered if the display scratch does not fit), grid = CTAs; extra = slope / fast peaks / skip
// mask / volume / roll-off / peak output in use; disp = display outputs (points / pixels / minimum) requested
cudaError_t warp2_launch(int N, bool extra, bool disp, const KParams &kp, int grid, int *warps, cudaStream_t st, bool pdl, int device,
                         const char **name);
} // namespace wf
