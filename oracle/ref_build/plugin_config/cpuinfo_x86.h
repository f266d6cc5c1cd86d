// Stand-in for google/cpu_features' cpuinfo_x86.h (submodule not needed by the oracle).
// The reference only reads features.avx2/avx/fma3 (src/source.cpp:34-39). Test infrastructure only.
#pragma once
namespace cpu_features {
struct X86Features { int avx2, avx, fma3; };
struct X86Info { X86Features features; };
static inline X86Info GetX86Info()
{
    X86Info i{};
    i.features.avx2 = __builtin_cpu_supports("avx2") ? 1 : 0;
    i.features.avx = __builtin_cpu_supports("avx") ? 1 : 0;
    i.features.fma3 = __builtin_cpu_supports("fma") ? 1 : 0;
    return i;
}
}
