// Device-resident random-walk Metropolis over the parameters of a sigmoidal branching process (SURVEY §8(f)-1,
// BASELINE config 4).  The reference ships no sampler (README.md:3: its analyses moved to another repository); this is
// the consumer the batched likelihood serves: C chains advance in lock-step, one batched likelihood evaluation of C
// parameter sets per iteration, proposals and accept/reject decided on the device so that an iteration costs no
// host round trip.  θ = (log xscale, xshift, log yscale, log yshift, log μ, log δ) of `SigmoidalBranchingProcess`
// (src/types.jl:65-110); flat prior on a box; Γ, ρ, σ and the type space are fixed.
//
// Random numbers are counter based — a hash of (seed, iteration, chain, slot) — so every rank of a multi-GPU run
// draws identical proposals without communication, and the host mirror (gcdyn/mcmc.py) reproduces the stream.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gcdyn {

constexpr int MH_DIM = 6;

__host__ __device__ __forceinline__ uint64_t mh_mix(uint64_t z) {      // splitmix64 finaliser
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
// uniform in (0, 1): 53 bits, never 0
__host__ __device__ __forceinline__ double mh_uniform(uint64_t seed, uint64_t it, uint64_t chain, uint64_t slot) {
    const uint64_t h = mh_mix(mh_mix(mh_mix(seed) ^ it) ^ (chain * 16 + slot));
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

struct MhView {
    int C, K;
    double* theta;          // [C][6] current state
    double* prop;           // [C][6] proposal of this iteration
    double* ll;             // [C] current log-likelihood
    double* ll_prop;        // [C] log-likelihood of the proposals (all-reduced over ranks)
    int* inside;            // [C] proposal inside the prior box
    long long* accepted;    // [C]
    double* xscale; double* xshift; double* yscale; double* yshift; double* mu; double* delta;   // [C] parameter blob
    double lower[MH_DIM], upper[MH_DIM];
    double step;
    uint64_t seed;
    // device-side bookkeeping, so that ONE captured CUDA graph serves every iteration: the number of completed
    // iterations (the kernels derive their random-number counter from it) and where history rows go
    unsigned long long* it_dev;     // [1]
    int* tick_count;                // [1] CTAs of the closing kernel that are done with this iteration (fused form)
    struct Hist { double* theta; double* ll; long long base; long long rows; }* hist;   // [1]
};

// Coordinate j of chain c's Gaussian random-walk proposal (Box–Muller on slots 2j, 2j+1): writes the proposal and the
// canonical parameter of the (clipped) proposal into the likelihood's parameter blob; returns whether it lies in the box.
__device__ __forceinline__ bool mh_propose_coord(const MhView& v, uint64_t it, int c, int j) {
    const double u1 = mh_uniform(v.seed, it, c, 2 * j), u2 = mh_uniform(v.seed, it, c, 2 * j + 1);
    const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    const double x = v.theta[c * MH_DIM + j] + v.step * z;
    v.prop[c * MH_DIM + j] = x;
    // proposals outside the box are rejected without looking at their likelihood, but still evaluated
    // (clipped), so that every iteration is one batch of exactly C parameter sets
    const double q = fmin(fmax(x, fmax(v.lower[j], -50.0)), fmin(v.upper[j], 50.0));
    double* dst = j == 0 ? v.xscale : j == 1 ? v.xshift : j == 2 ? v.yscale : j == 3 ? v.yshift : j == 4 ? v.mu : v.delta;
    dst[c] = j == 1 ? q : exp(q);
    return (x >= v.lower[j]) && (x <= v.upper[j]);
}

// Metropolis accept/reject of chain c (flat prior: the ratio is the likelihood ratio); history rows are optional.
__device__ __forceinline__ void mh_accept_chain(const MhView& v, uint64_t it, int c) {
    const long long hist_row = (long long)(it - 1) - v.hist->base;
    double* __restrict__ hist_theta = (hist_row >= 0 && hist_row < v.hist->rows) ? v.hist->theta : nullptr;
    double* __restrict__ hist_ll = v.hist->ll;
    const double logu = log(mh_uniform(v.seed, it, c, 2 * MH_DIM));
    const double llp = v.ll_prop[c], llc = v.ll[c];
    const bool acc = v.inside[c] && (logu < llp - llc);        // NaN compares false: never accepted
    if (acc) {
#pragma unroll
        for (int j = 0; j < MH_DIM; ++j) v.theta[c * MH_DIM + j] = v.prop[c * MH_DIM + j];
        v.ll[c] = llp;
        v.accepted[c] += 1;
    }
    if (hist_theta) {
#pragma unroll
        for (int j = 0; j < MH_DIM; ++j) hist_theta[((size_t)hist_row * v.C + c) * MH_DIM + j] = v.theta[c * MH_DIM + j];
        hist_ll[(size_t)hist_row * v.C + c] = v.ll[c];
    }
}

// The same steps as stand-alone kernels (one thread per chain), for forests that take the item sweep.
__global__ void mh_propose_kernel(MhView v) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= v.C) return;
    const uint64_t it = *v.it_dev + 1;
    bool in = true;
#pragma unroll
    for (int j = 0; j < MH_DIM; ++j) in = mh_propose_coord(v, it, c, j) && in;
    v.inside[c] = in ? 1 : 0;
}

__global__ void mh_accept_kernel(MhView v) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= v.C) return;
    mh_accept_chain(v, *v.it_dev + 1, c);
}

// one iteration completed (its own launch: every block of mh_accept_kernel has read the counter by then)
__global__ void mh_tick_kernel(MhView v) { *v.it_dev += 1; }

// parameters of the CURRENT state into the blob (initial likelihood)
__global__ void mh_state_params_kernel(MhView v) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= v.C) return;
    const double* t = v.theta + c * MH_DIM;
    v.xscale[c] = exp(t[0]);
    v.xshift[c] = t[1];
    v.yscale[c] = exp(t[2]);
    v.yshift[c] = exp(t[3]);
    v.mu[c] = exp(t[4]);
    v.delta[c] = exp(t[5]);
}

}  // namespace gcdyn
