// Small helpers shared by host and device code of libgcdyn_b200.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GCDYN_HD __host__ __device__ __forceinline__
#else
#define GCDYN_HD inline
#endif

namespace gcdyn {

// Step counts are quantised to the ladder {4,5,6,8,10,12,16,20,24,32,...} (×1.25, ×1.2, ×1.33 …) so that
// pset-independent forest statistics (the moment tables of the totals path) exist per rung, not per value.
GCDYN_HD int ladder_ceil(long long x) {
    if (x <= 4) return 4;
    long long base = 4;
    while (2 * base < x) base *= 2;          // base < x <= 2·base
    if (x <= base) return (int)base;
    if (x <= base + base / 4) return (int)(base + base / 4);
    if (x <= base + base / 2) return (int)(base + base / 2);
    return (int)(2 * base);
}

GCDYN_HD int ladder_next(int s) { return ladder_ceil((long long)s + 1); }

// h·rate targets.  `safe` was calibrated on the hardest case found (ρ=1, σ=0, γ=0,
// λ≫μ: every cell needs the full order); `start` is the optimistic first guess of the in-kernel
// adaptive loop, which restarts with more steps as soon as the error indicator exceeds tol.
GCDYN_HD double hrate_target_safe(int M) {
    return M <= 8 ? 0.055 : M <= 12 ? 0.20 : M <= 16 ? 0.36 : M <= 20 ? 0.50 : 0.62;
}
GCDYN_HD double hrate_target_start(int M) {
    return M <= 8 ? 0.15 : M <= 12 ? 0.55 : M <= 16 ? 1.0 : M <= 20 ? 1.4 : 1.8;
}

// Widest cell, as h·rate, the a-posteriori width control of traj_kernel may take: beyond ≈ M/3 the last two Taylor terms
// of a mode e^{±rate·τ} are no longer decreasing and stop being a usable estimate of the truncation error.
GCDYN_HD double hrate_cap(int M) { return 0.35 * (double)M; }

}  // namespace gcdyn
