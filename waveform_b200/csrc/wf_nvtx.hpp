// wf_nvtx.hpp — NVTX ranges around the C-ABI's processing calls (SURVEY §5 tracing hooks): header-only NVTX3, a no-op unless a
// profiler (Nsight Systems / Compute) is attached.  One range per call, named after the entry point; the kernel a call
// dispatched to is available through wf_last_kernel_name().
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace wf {
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
} // namespace wf
