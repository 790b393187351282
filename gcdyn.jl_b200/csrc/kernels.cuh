// Device kernels of libgcdyn_b200 (sm_100a, FP64).  See DESIGN.md §3-§5.
//
// Pipeline of one batched ODE likelihood evaluation (P parameter sets × T trees):
//   traj_kernel   one CTA of four warps per pset.  One warp (two for 32 < K ≤ 64) integrates the extinction ODE
//                 (dp_dt!, mbp.jl:266-278) with a fixed-order Taylor series on an adaptive grid of cells and
//                 publishes every accepted cell through a shared-memory ring; the other warps compute rates and
//                 node-term logs (λ, μ, γ of mbp.jl:27-97), build per cell and type the polynomial of
//                 F_x(τ) = ∫(−Λ_x + 2 λ_x p_x) — the branch density of dp_logq_dt!, mbp.jl:285-300 — write that
//                 table, and evaluate F at the Chebyshev nodes — those values are the pset's row of GEMM coefficients
//                 (cheb_gemm.cuh keeps the forest's moments in the nodal basis) — and log(1 − p_x(T)) (mbp.jl:244-245)
//   gemm_kernel   (cheb_gemm.cuh) LL[p][t] = Σ_j A[p][j]·Mt[t][j] with DMMA, split over the column range
//   finish_kernel (cheb_gemm.cuh) combines the splits, applies −Inf/status rules, recomputes flagged psets exactly
//                 from the table, and sums over trees → out_pset[p]   (mbp.jl:258)
// Small batches without cached moments, or GCDYN_FLAG_SWEEP_ONLY:
//   sweep_kernel  persistent CTAs over (pset, tree-chunk) units; the pset's table is staged into shared
//                 memory by the bulk-copy engine (cp.async.bulk + mbarrier); a warp per tree sums
//                 w·F_type(τ) + node terms (mbp.jl:167-191) over the tree's lookup items,
//                 warp-shuffle reduction → out_tree[p*T + t]
//   reduce_kernel per pset: fixed-order sum over trees → out_pset[p]
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "common.h"
#include "metropolis.cuh"

namespace gcdyn {

constexpr int EV_ROOT = 0, EV_BIRTH = 1, EV_SDEATH = 2, EV_UDEATH = 3, EV_TC = 4, EV_SSURV = 5, EV_USURV = 6;
constexpr int ST_SELF_LOOP = 1, ST_SOLVER = 2, ST_DOMAIN = 4;
constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// Device views.  All arrays are plain device pointers; K = number of types.
// ------------------------------------------------------------------------------------------------
struct ParamsView {              // raw parameter batch as uploaded (one contiguous blob)
    int P, K, response, gamma_stride;
    const double* lambda;        // TABLE [P*K] | CONSTANT [P] | SIGMOID nullptr
    const double *xscale, *xshift, *yscale, *yshift, *type_space;
    const double *mu, *delta, *rho, *sigma;
    const double* Gamma;         // column-major K×K, 1 or P matrices
};

struct PsetPack {                // derived per-pset quantities (workspace)
    int ntw;                     // row width of `nt` = K + K*K + 2
    double* lam;                 // [P*K]  λ_i
    double* Lam;                 // [P*K]  Λ_i = λ_i + μ + γ_i,  γ_i = −δ Γ[i,i]      (mbp.jl:67-75)
    double* nt;                  // [P*ntw] log λ_x | log(δΓ[x,y]) row-major, −Inf on the diagonal | log μ + log σ | log ρ
    double* logcond;             // [P*K]  log(1 − p_i(T)) (ODE) or log(1 − p_stadler,i) (STADLER)
    double* st_c;                // [P*K]  Stadler c_x                                   (extra_likelihoods.jl:52)
    double* st_A;                // [P*K]  y + λ(1−ρ)
    double* st_B;                // [P*K]  x + λ(1−ρ)
    int* pstatus;                // [P]
    int* S;                      // [P]    Taylor cells used for this pset
    double* perr;                // [P]    Taylor error indicator
};

struct ForestView {
    int64_t n_trees, n_nodes, n_items;
    int K;
    // as flattened by the host (post-order SoA)
    const int64_t* tree_off;
    const uint8_t* event;
    const int32_t* type_idx;
    const int32_t* btype_idx;
    const double* t_start;
    const double* t_end;
    const uint8_t* live;
    const int32_t* root_type;
    const double* root_time;
    const uint8_t* self_loop;
    // derived lookup items of the regrouped ODE sum (tree-major CSR)
    const int64_t* item_off;     // [T+1]
    const double* it_time;       // [n_items] node time t (τ = present_time − t)
    const int2* it_meta;         // [n_items] .x = type | (int8 weight << 24), .y = index into nt row (−1 = none)
    const int32_t* n_death;      // [T]
    const int32_t* n_surv;       // [T]
};

constexpr int TRAJ_THREADS = 128;   // warp 0 integrates; all four warps run the Chebyshev stage

struct TrajCfg {
    double T;                    // present_time
    double tol;                  // accepted error indicator
    int S_fixed;                 // > 0: use exactly this many cells for every pset (no adaptivity)
    int S_cap;                   // table capacity in cells (adaptive mode never exceeds it)
    int grid_doubles;            // doubles reserved per pset for [τ_s | 1/h_s] ahead of the rows (even)
    long long tab_stride;        // doubles per pset in the table (grid block + rows)
    // trajectory sharding (trees AND parameter sets split over ranks): this launch covers psets p_base + blockIdx.x;
    // only_missing: a second launch over all psets in which a CTA works only for a pset OUTSIDE [own_lo, own_hi) that its
    // owner flagged for the exact fallback — the fallback sums need the Taylor table, which is not exchanged
    int p_base, only_missing, own_lo, own_hi;
    int n_ctas;                  // grid size of this launch (0: one CTA per pset of the batch)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ double warp_max_nan(double v) {   // NaN-propagating maximum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(FULL, v, o);
        if (!(other <= v)) v = other;
    }
    return v;
}

__device__ __forceinline__ double count_times_log(int n, double lg) {
    return n == 0 ? 0.0 : (double)n * lg;      // 0 · (−Inf) must stay 0: the reference never forms that term
}

// λ(model, type) for type index i of pset p (mbp.jl:27-43; sigmoid at :4-7)
__device__ __forceinline__ double birth_rate(const ParamsView& pv, int p, int i) {
    if (pv.response == 0) return pv.lambda[(size_t)p * pv.K + i];
    if (pv.response == 1) return pv.lambda[p];
    const double e = exp(-(pv.xscale[p] * (pv.type_space[i] - pv.xshift[p])));
    return pv.yscale[p] * (1.0 / (1.0 + e)) + pv.yshift[p];
}

// Rates and node-term logs of one pset, computed cooperatively by `nthr` threads (tid = 0..nthr-1).
__device__ __forceinline__ void prep_tc_logs(const ParamsView& pv, const PsetPack& pk, int p, int tid, int nthr) {
    const int K = pv.K;
    const double delta = pv.delta[p];
    const double* G = pv.Gamma + (size_t)p * pv.gamma_stride;   // column-major: Γ[i,j] = G[i + j*K]
    double* nt = pk.nt + (size_t)p * pk.ntw;
    for (int e = tid; e < K * K; e += nthr) {
        const int x = e / K, y = e - x * K;
        nt[K + e] = (x == y) ? -CUDART_INF : log(delta * G[x + (size_t)y * K]);   // γ(model, from, to), mbp.jl:84-97
    }
}

__device__ __forceinline__ void prep_pset(const ParamsView& pv, const PsetPack& pk, int p, int method, int tid, int nthr,
                                          bool init_status = true, bool tc_logs = true) {
    const int K = pv.K;
    const double mu = pv.mu[p], delta = pv.delta[p], rho = pv.rho[p], sigma = pv.sigma[p];
    const double* G = pv.Gamma + (size_t)p * pv.gamma_stride;   // column-major: Γ[i,j] = G[i + j*K]
    double* nt = pk.nt + (size_t)p * pk.ntw;
    for (int i = tid; i < K; i += nthr) {
        const double lam = birth_rate(pv, p, i);
        const double gam = delta * -G[i + (size_t)i * K];
        const double Lam = lam + mu + gam;
        pk.lam[(size_t)p * K + i] = lam;
        pk.Lam[(size_t)p * K + i] = Lam;
        nt[i] = log(lam);
        if (method == 2 || method == 3) {   // Stadler per-type constants, extra_likelihoods.jl:50-57
            const double c = sqrt(Lam * Lam - 4.0 * mu * (1.0 - sigma) * lam);
            const double x = (-Lam - c) / 2.0, y = (-Lam + c) / 2.0;
            pk.st_c[(size_t)p * K + i] = c;
            pk.st_A[(size_t)p * K + i] = y + lam * (1.0 - rho);
            pk.st_B[(size_t)p * K + i] = x + lam * (1.0 - rho);
        }
    }
    if (tc_logs) prep_tc_logs(pv, pk, p, tid, nthr);
    if (tid == 0) {
        nt[K + K * K] = log(mu) + log(sigma);   // mbp.jl:171
        nt[K + K * K + 1] = log(rho);           // mbp.jl:168
        if (init_status) {
            pk.pstatus[p] = 0;
            pk.perr[p] = 0.0;
            pk.S[p] = 0;
        }
    }
}

__global__ void prep_kernel(ParamsView pv, PsetPack pk, int method) {
    prep_pset(pv, pk, blockIdx.x, method, threadIdx.x, blockDim.x);
}

// Stadler conditioning: log(1 − p) for a root of each type, extra_likelihoods.jl:82-101.
__global__ void stadler_cond_kernel(ParamsView pv, PsetPack pk, double present_time) {
    const int p = blockIdx.x, K = pv.K;
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const size_t o = (size_t)p * K + i;
        const double lam = pk.lam[o], c = pk.st_c[o], A = pk.st_A[o], B = pk.st_B[o];
        const double Lam = pk.Lam[o];
        const double x = (-Lam - c) / 2.0, y = (-Lam + c) / 2.0;
        const double e = exp(-c * present_time);
        const double pr = -1.0 / lam * (A * x * e - y * B) / (A * e - B);
        pk.logcond[o] = log(1.0 - pr);
    }
}

struct MomentView {
    int K;
    int C1;                // 2K + 4: start of the nodal block
    int D;                 // degree of the Chebyshev–Gauss–Lobatto grid: nodes k = 0..D, τ_k = T(1 + cos(πk/D))/2
    long long J1;          // end of the nodal block, rounded up to a multiple of 4
    long long Jtc;         // start of the K² type-change block
    long long Jst;         // row stride of Mt and A (doubles), multiple of 4
    const double* Mt;      // [T][Jst]
    const double* colsum;  // [Jst] forest-level column sums
};

// Trajectory sharding: the CTA that finished a parameter set's coefficient row stores it straight into the row windows of
// the other ranks (P2P stores over NVLink, fenced at system scope before the CTA retires); row_exchange_kernel
// (peer_allreduce.cuh) then only publishes the metadata and the flags.  n == 0: off.
constexpr int ROW_PUSH_MAX = 15;
struct RowPush {
    int n;                       // peers to push to
    int w2;                      // double2 per row that the contraction uses
    double* dst[ROW_PUSH_MAX];   // the peers' coefficient buffers (same layout as ChebCfg::A)
};

struct ChebCfg {
    double T;
    double tol;            // accepted size of the interpolant's last three Chebyshev coefficients (absolute, on F)
    int D;                 // = mv.D; 0: no Chebyshev stage (item sweep only)
    int compact_tc;        // 1: Γ shared — log δ column + per-tree constant; 0: K² type-change columns
    MomentView mv;
    double* A;             // [P][Jst] coefficient rows of the GEMM: count columns | F_x(τ_k), node-major | K² block
    int* need_sweep;       // [P] fallback flags: 1 = interpolant unresolved, 2 = non-finite node-term logs
    int* nflag;            // [1] psets of this batch that were unresolved (reset by traj_kernel, counted by finish_kernel)
    // GCDYN_FLAG_TOTALS: out_pset[p] = Σ_j A[p][j]·colsum[j] + tc_total straight from traj_kernel (nullptr = off)
    double* totals;        // [P]
    const double* tc_total;   // [1] Σ_t tree_const[t] (shared Γ), or nullptr
    int any_self_loop;     // the forest holds a tree the reference gives -Inf for
    // fused Metropolis iteration (metropolis.cuh): the CTA of pset p draws chain p's proposal before it integrates and,
    // when nothing is left to combine (one rank, pset resolved), decides accept/reject right after its sum
    int mh_on, mh_sharded;
    MhView mh;
    RowPush push;          // trajectory sharding: send the finished row to the peers
};

// Cell containing τ on a pset's (non-uniform) grid: largest s < S with grid[s] ≤ τ.
__device__ __forceinline__ int find_cell(const double* __restrict__ grid, int S, double tau) {
    int lo = 0, hi = S - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (grid[mid] <= tau) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// F_x(τ) from a table row of L = M+2 coefficients: even/odd split halves the dependent FMA chain.
// COHERENT: the row was written earlier by this same kernel (ld.global.cg instead of the read-only path).
template <int M, bool SMEM, bool COHERENT = false>
__device__ __forceinline__ double eval_row(const double* __restrict__ row, double th) {
    constexpr int L = M + 2;
    double co[L];
#pragma unroll
    for (int n = 0; n < L; n += 2) {
        double2 v;
        if constexpr (SMEM) v = *reinterpret_cast<const double2*>(row + n);
        else if constexpr (COHERENT) v = __ldcg(reinterpret_cast<const double2*>(row + n));
        else v = __ldg(reinterpret_cast<const double2*>(row + n));
        co[n] = v.x;
        co[n + 1] = v.y;
    }
    const double t2 = th * th;
    double ev = co[L - 2], od = co[L - 1];
#pragma unroll
    for (int n = L - 4; n >= 0; n -= 2) {
        ev = fma(ev, t2, co[n]);
        od = fma(od, t2, co[n + 1]);
    }
    return fma(od, th, ev);
}

// ------------------------------------------------------------------------------------------------
// Bulk-copy (TMA) helpers: one elected thread arms an mbarrier with the byte count and issues
// cp.async.bulk global→shared; everybody waits on the barrier's phase parity.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Programmatic dependent launch: a kernel launched with the stream-serialization attribute may become resident while
// its predecessor still runs; pdl_wait() blocks until the predecessor has completed and its writes are visible,
// pdl_launch_dependents() lets the successor start becoming resident.  Both are no-ops in an ordinary launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#ifndef GCDYN_CONS_POLL
#define GCDYN_CONS_POLL 0
#endif
// A whole warp waits for a phase: variant 0 every lane polls; 1 lane 0 polls with a short sleep between attempts and the
// others park at the warp barrier; 2 lane 0 polls back to back.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
#if GCDYN_CONS_POLL == 0
    mbar_wait(bar, parity);
#else
    if ((threadIdx.x & 31) == 0) {
        for (;;) {
            uint32_t ok;
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
            if (ok) break;
#if GCDYN_CONS_POLL == 1
            __nanosleep(100);
#endif
        }
    }
    __syncwarp();
#endif
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// traj_kernel: one CTA of 4 warps per pset.
//
// Integration (NPW warps: one for K ≤ 32 — lane l owns type l; two for 32 < K ≤ 64 — thread t of the pair owns type t;
// one for K > 64 — lane l owns types l, l+32, …, δΓ staged in shared memory).  The ODE is polynomial
// (dp_dt!, mbp.jl:266-278), so the DERIVATIVES d_n = p^(n)(τ_s) at the start of a cell follow a recurrence that does not
// involve the cell width h:
//     d_1     = G d_0 + (λ d_0 − a) d_0 + b                          a = λ + μ,  b = μ(1 − σ),  G = δΓ
//     d_{n+1} = G d_n + κ d_n + λ n! R_n,     κ = 2λ d_0 − a,   R_n = Σ_{m=1}^{n−1} A_m A_{n−m},   A_n = d_n / n!
// Per Taylor stage the critical path is: broadcast d_n through shared memory → K-term dot product whose first chain is
// seeded with κ d_n + λ n! R_n (both known before the broadcast returns) → two DADDs.  The convolution, the factorials
// and the publication of A_n sit in the shadow of the broadcast.  Only AFTER the M derivatives are known is the width
// chosen — the largest h with |A_M| h^M + |A_{M−1}| h^{M−1} ≤ tol/2 over the pset's types (high-word log2, one MUFU.EX2),
// capped at hrate_cap(M)/rate — so no cell is ever rejected or recomputed, and p(τ_s + h) = Σ A_n h^n is one Estrin
// evaluation.  A result outside [0, 1] (isoutofdomain, mbp.jl:146,204,231) halves h and re-evaluates the same
// polynomial: the reference's reject-and-shrink at no extra derivative work; a width below 1e-10·T is a solver failure
// (the reference's dt < dtmin → non-success retcode → -Inf, mbp.jl:154-158).
//
// The other warps: rates and node-term logs first (λ, μ, γ of mbp.jl:27-97), then per published cell (A_0..A_M per
// type, τ_s, h through a 6-slot shared-memory ring guarded by mbarriers) the polynomial of
// F_x(τ) = ∫(−Λ_x + 2 λ_x p_x) — the branch density of dp_logq_dt!, mbp.jl:285-300 — on the cell,
//     row[0] = F(τ_s),  row[1] = h(2λA_0 − Λ),  row[n+1] = 2λ A_n h^{n+1}/(n+1),     θ = (τ − τ_s)/h,
// written to the table in global memory (exact item sums of flagged psets) and evaluated at the Chebyshev–Gauss–Lobatto
// nodes inside the cell.  Those node values ARE the pset's GEMM coefficients (the forest's moments are kept in the
// nodal basis, cheb_gemm.cuh), so nothing but a resolution check and the count columns remains after the last cell.
// ------------------------------------------------------------------------------------------------
#ifdef GCDYN_TRAJ_TRACE
__device__ long long g_trace_out[1024 * 24];
#define TRACE_STAMP(i) do { if ((threadIdx.x & 31) == 0) g_trace[i] = clock64(); } while (0)
#else
#define TRACE_STAMP(i) do { } while (0)
#endif

#ifndef GCDYN_RING_SLOTS
#define GCDYN_RING_SLOTS 6
#endif
// (the K > 64 layout keeps δΓ — K² doubles — in shared memory as well: two cells in flight there, or K = 128 would not fit)
constexpr int RING_SLOTS_WIDE = 2;
__host__ __device__ constexpr int ring_slots(int KPL) { return KPL == 1 ? GCDYN_RING_SLOTS : RING_SLOTS_WIDE; }
constexpr int RING_SLOTS = GCDYN_RING_SLOTS;                 // cells in flight between the integrating warp(s) and the consumer warps
constexpr int CELL_LAST = 1, CELL_FAIL = 2;   // flags of a published cell

__host__ __device__ constexpr double factd(int n) { double f = 1.0; for (int i = 2; i <= n; ++i) f *= (double)i; return f; }

template <int N>
__device__ __forceinline__ double powi(double h) {
    if constexpr (N == 0) return 1.0;
    else if constexpr (N & 1) return h * powi<N - 1>(h);
    else { const double t = powi<N / 2>(h); return t * t; }
}

// Σ_{n ≤ M} A[n] h^n by Estrin's scheme (depth ⌈log2(M+1)⌉ instead of M dependent FMAs)
template <int LEN, int N>
__device__ __forceinline__ void estrin_level(double (&cur)[N], double pw) {
#pragma unroll
    for (int i = 0; 2 * i < LEN; ++i) cur[i] = (2 * i + 1 < LEN) ? fma(cur[2 * i + 1], pw, cur[2 * i]) : cur[2 * i];
    if constexpr ((LEN + 1) / 2 > 1) estrin_level<(LEN + 1) / 2, N>(cur, pw * pw);
}
template <int M>
__device__ __forceinline__ double estrin(const double (&A)[M + 1], double h) {
    double cur[M + 1];
#pragma unroll
    for (int i = 0; i <= M; ++i) cur[i] = A[i];
    estrin_level<M + 1, M + 1>(cur, h);
    return cur[0];
}

struct TrajSmem {                      // carve-up of the kernel's dynamic shared memory behind the broadcast buffers
    double* costab;                    // [2D]  cos(π m / D)
    double* tauk;                      // [D+2] τ_k = T(1 + cos(πk/D))/2
    double* V;                         // [(D+1)*K] node values F_x(τ_k)
    double* chk;                       // [8][3][K] partial sums of a_D, a_{D−1}, a_{D−2} of the interpolant (resolution check)
    double* meta;                      // [RING_SLOTS][4]  τ_s, τ_{s+1}, h_s, flags of the cells in flight
    double* ring;                      // [RING_SLOTS][(M+1)*KS]  A_n[x], n-major (conflict-free for lanes = types)
    double* rowbuf;                    // [2][K*L] table rows of the cell being evaluated at the Chebyshev nodes
    int KS;
    __device__ TrajSmem(double* sm, int K, int D, int M, int slots) {
        KS = (K + 1) & ~1;
        costab = sm;
        tauk = costab + 2 * D;
        V = tauk + (D > 0 ? ((D + 3) & ~1) : 0);
        chk = V + (D > 0 ? ((((size_t)(D + 1) * K) + 1) & ~(size_t)1) : 0);
        meta = chk + (D > 0 ? 24 * (size_t)K : 0);
        ring = meta + slots * 4;
        rowbuf = ring + (size_t)slots * (M + 1) * KS;
    }
};

// Count columns of the pset's coefficient row that do not depend on the trajectory (everything but −log(1 − p_x(T))),
// written by the consumer warps while the integration runs.  A non-finite log the forest actually uses flags the pset
// (bit 2): finish_kernel then sums its items exactly, where 0·log 0 terms are never formed.
__device__ __forceinline__ void early_count_columns(const ParamsView& pv, const PsetPack& pk, const ChebCfg& cfg, int p,
                                                    int tid, int nthr, int* s_flag) {
    const MomentView& mv = cfg.mv;
    const int K = mv.K, C1 = mv.C1;
    double* Arow = cfg.A + (size_t)p * mv.Jst;
    const double* nt = pk.nt + (size_t)p * pk.ntw;
    for (int c = tid; c < C1; c += nthr) {
        if (c >= K + 2 && c < 2 * K + 2) continue;                     // −log(1 − p_x(T)): the integrating lanes, at the end
        double v = 0.0;
        if (c < K) v = nt[c];                                         // log λ_x
        else if (c == K) v = nt[K + K * K];                           // log μ + log σ
        else if (c == K + 1) v = nt[K + K * K + 1];                   // log ρ
        else if (c == 2 * K + 2) v = cfg.compact_tc ? log(pv.delta[p]) : 0.0;
        if (!isfinite(v)) {
            if (mv.colsum[c] != 0.0) atomicOr(s_flag, 2);
            v = 0.0;
        }
        Arow[c] = v;
    }
    if (!cfg.compact_tc) {
        for (int e = tid; e < K * K; e += nthr) {
            double v = nt[K + e];
            const bool diag = (e / K) == (e % K);          // self-loop column: trees using it are −Inf anyway
            if (!isfinite(v)) {
                if (!diag && mv.colsum[mv.Jtc + e] != 0.0) atomicOr(s_flag, 2);
                v = 0.0;
            }
            Arow[mv.Jtc + e] = v;
        }
    }
    // the GEMM rounds its column ranges up to multiples of 4: keep the padding finite
    if (tid < 4) {
        const long long j = C1 + (long long)(mv.D + 1) * K + tid;
        if (j < mv.Jtc) Arow[j] = 0.0;
        const long long j2 = mv.Jtc + (long long)K * K + tid;
        if (j2 < mv.Jst) Arow[j2] = 0.0;
    }
}

// The warps that do not integrate (ct = 0..nct-1).  Node k of the Chebyshev grid sits at τ_k = T(1 + cos(πk/D))/2, so
// cells (increasing τ) meet the nodes in decreasing k; a node belongs to the largest cell with τ_s ≤ τ_k (find_cell's
// rule), the last cell takes the rest.
template <int M, int SLOTS>
__device__ __forceinline__ void cell_consume(const ParamsView& pv, const TrajCfg& tc, const ChebCfg& cfg, bool cheb, int p,
                                             double* __restrict__ tab, const TrajSmem& ts, uint64_t* full, uint64_t* empty,
                                             int ct, int nct, const PsetPack& pk, int* s_flag) {
    constexpr int L = M + 2;
    constexpr int TPT = 2;                              // types per consumer thread: K ≤ 128, nct ≥ 64
    const int K = pv.K, D = cfg.D, KS = ts.KS;
    const double mu = pv.mu[p], delta = pv.delta[p];
    const double* G = pv.Gamma + (size_t)p * pv.gamma_stride;
    double lamx[TPT], Lamx[TPT], Fc[TPT];
#pragma unroll
    for (int j = 0; j < TPT; ++j) {
        const int x = ct + j * nct;
        lamx[j] = x < K ? birth_rate(pv, p, x) : 0.0;
        const double gam = x < K ? delta * -G[x + (size_t)x * K] : 0.0;
        Lamx[j] = lamx[j] + mu + gam;
        Fc[j] = 0.0;
    }
    if (cheb) {
        for (int m = ct; m < 2 * D; m += nct) {
            const double c = cospi((double)m / (double)D);
            ts.costab[m] = c;
            if (m <= D) ts.tauk[m] = cfg.T * (1.0 + c) * 0.5;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(nct) : "memory");      // cos table and the node-term logs of prep_pset are complete
        early_count_columns(pv, pk, cfg, p, ct, nct, s_flag);
    }
    double* tabp = tab + (size_t)p * tc.tab_stride;
    double* gridp = tabp;
    double* invhp = tabp + (tc.S_cap + 1);
    double* rowsp = tabp + tc.grid_doubles;
    double* Anod = cheb ? cfg.A + (size_t)p * cfg.mv.Jst + cfg.mv.C1 : nullptr;
    int kn = D;
    for (int sc = 0;; ++sc) {
        const int slot_i = sc % SLOTS;
        mbar_wait_warp(full + slot_i, (uint32_t)(sc / SLOTS) & 1u);
        const double* meta = ts.meta + slot_i * 4;
        const int flag = (int)meta[3];
        if (flag & CELL_FAIL) {
            if (ct == 0) gridp[sc] = tc.T;
            break;
        }
        const double t0 = meta[0], t1 = meta[1], hc = meta[2];
        const double ih = 1.0 / hc;
        const double* slot = ts.ring + (size_t)slot_i * (M + 1) * KS;
        double* rb = ts.rowbuf + (size_t)(sc & 1) * K * L;
#pragma unroll
        for (int j = 0; j < TPT; ++j) {
            const int x = ct + j * nct;
            if (x < K) {
                double row[L];
                const double a0 = slot[x];
                const double l2 = 2.0 * lamx[j];
                row[0] = Fc[j];
                row[1] = hc * (l2 * a0 - Lamx[j]);
                double hp = hc;                               // h^(n+1)
#pragma unroll
                for (int n = 1; n <= M; ++n) {
                    hp *= hc;
                    row[n + 1] = (l2 * (1.0 / (double)(n + 1))) * (slot[(size_t)n * KS + x] * hp);
                }
                double fs[4] = {0.0, 0.0, 0.0, 0.0};          // smallest terms first, four partial sums
#pragma unroll
                for (int n = M; n >= 1; --n) fs[n & 3] += row[n + 1];
                Fc[j] += ((fs[0] + fs[1]) + (fs[2] + fs[3])) + row[1];
                double2* dst = reinterpret_cast<double2*>(rowsp + ((size_t)sc * K + x) * L);
#pragma unroll
                for (int n = 0; n < L; n += 2) dst[n >> 1] = make_double2(row[n], row[n + 1]);
                if (cheb) {
                    double2* d2 = reinterpret_cast<double2*>(rb + (size_t)x * L);
#pragma unroll
                    for (int n = 0; n < L; n += 2) d2[n >> 1] = make_double2(row[n], row[n + 1]);
                }
            }
        }
        mbar_arrive(empty + slot_i);                    // this thread has read what it needs from the slot
        if (ct == 0) {
            gridp[sc] = t0;
            invhp[sc] = ih;
            if (flag & CELL_LAST) gridp[sc + 1] = tc.T;
        }
        if (cheb) {
            // rows of this cell are complete in rb; rb alternates, and the barrier of the next cell keeps the cell
            // after next from overwriting rows that are still being read
            asm volatile("bar.sync 1, %0;" ::"r"(nct) : "memory");
            int klo = kn;                               // nodes klo+1..kn lie in this cell
            if (flag & CELL_LAST) klo = -1;
            else while (klo >= 0 && ts.tauk[klo] < t1) --klo;
#ifdef GCDYN_EXP_NO_NODES
            const int nitems = 0;
#else
            const int nitems = (kn - klo) * K;
#endif
            for (int it = ct; it < nitems; it += nct) {
                const int k = kn - it / K, x = it % K;
                const double v = eval_row<M, true>(rb + (size_t)x * L, (ts.tauk[k] - t0) * ih);
                ts.V[k * K + x] = v;
                Anod[k * K + x] = v;
            }
            kn = klo;
        }
        if (flag & CELL_LAST) break;
    }
}

// After the last cell, all threads: resolution check of the degree-D interpolant (its last three Chebyshev coefficients
// must be below tol — otherwise the pset is flagged and finish_kernel sums its items exactly from the table) and the
// fallback flag.
__device__ __forceinline__ void nodal_finish(const ChebCfg& cfg, int p, const TrajSmem& ts, int* s_flag, int pst,
                                             unsigned long long mh_it) {
    const MomentView& mv = cfg.mv;
    const int K = mv.K, C1 = mv.C1, D = cfg.D, NV = (D + 1) * K;
    double* V = ts.V;
    double* Arow = cfg.A + (size_t)p * mv.Jst;
    if (pst & ST_SOLVER) {                               // CTA-uniform: a failed trajectory has no usable nodes; it is
        for (int idx = threadIdx.x; idx < NV; idx += blockDim.x) Arow[C1 + idx] = 0.0;   // -Inf through its status
        if (threadIdx.x == 0) {
            cfg.need_sweep[p] = 0;
            if (cfg.totals) {
                cfg.totals[p] = -CUDART_INF;
                if (cfg.mh_on && !cfg.mh_sharded) mh_accept_chain(cfg.mh, mh_it, p);
            }
        }
        return;
    }
    // a_{D−j} = (2/D) Σ''_k V_k (−1)^k cos(π k j / D) for j = 0, 1, 2 (a_D halved).  NSEG threads per type share the
    // node range; the partial sums are combined in segment order.
    const int NSEG = min(8, max(1, (int)blockDim.x / K));
    const int per = (D + NSEG) / NSEG;                   // nodes per segment (32-bit arithmetic only: this is the tail)
    for (int w = threadIdx.x; w < NSEG * K; w += blockDim.x) {
        const int seg = w / K, x = w - seg * K;
        const int k0 = seg * per, k1 = min(D + 1, k0 + per);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        int i2 = 2 * k0;                                 // (2k) mod 2D
        if (i2 >= 2 * D) i2 -= 2 * D;
        for (int k = k0; k < k1; ++k) {
            const double wk = (k == 0 || k == D) ? 0.5 : 1.0;
            const double v = (k & 1) ? -wk * V[k * K + x] : wk * V[k * K + x];
            s0 += v;
            s1 = fma(v, ts.costab[k], s1);
            s2 = fma(v, ts.costab[i2], s2);
            i2 += 2;
            if (i2 >= 2 * D) i2 -= 2 * D;
        }
        ts.chk[(seg * 3 + 0) * K + x] = s0;
        ts.chk[(seg * 3 + 1) * K + x] = s1;
        ts.chk[(seg * 3 + 2) * K + x] = s2;
    }
    __syncthreads();
    bool unresolved = false;
    for (int x = threadIdx.x; x < K; x += blockDim.x) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int seg = 0; seg < NSEG; ++seg) {
            a0 += ts.chk[(seg * 3 + 0) * K + x];
            a1 += ts.chk[(seg * 3 + 1) * K + x];
            a2 += ts.chk[(seg * 3 + 2) * K + x];
        }
        const double last3 = (2.0 / (double)D) * (0.5 * fabs(a0) + fabs(a1) + fabs(a2));
        if (!(last3 <= cfg.tol)) unresolved = true;
    }
    const int any_unres = __syncthreads_or(unresolved ? 1 : 0);
    const int flag = *s_flag | (any_unres ? 1 : 0);
    if (threadIdx.x == 0) cfg.need_sweep[p] = flag;
    if (cfg.totals && !flag) {
        // Σ_j A[p][j]·colsum[j]: the count columns (read back: other threads of this CTA wrote them before the barrier),
        // the node values still in shared memory, the K² block when every pset has its own Γ.  Fixed order.
        __shared__ double tpart[TRAJ_THREADS / 32];
        double acc = 0.0;
        for (int c = threadIdx.x; c < C1; c += blockDim.x) acc = fma(__ldcg(Arow + c), mv.colsum[c], acc);
        for (int idx = threadIdx.x; idx < NV; idx += blockDim.x) acc = fma(V[idx], mv.colsum[C1 + idx], acc);
        if (!cfg.compact_tc)
            for (int e = threadIdx.x; e < K * K; e += blockDim.x) acc = fma(__ldcg(Arow + mv.Jtc + e), mv.colsum[mv.Jtc + e], acc);
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) tpart[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double sum = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += tpart[w];
            if (cfg.tc_total) sum += *cfg.tc_total;
            cfg.totals[p] = cfg.any_self_loop ? -CUDART_INF : sum;
            if (cfg.mh_on && !cfg.mh_sharded) mh_accept_chain(cfg.mh, mh_it, p);   // totals IS the sampler's ll_prop
        }
    }
}

template <int M, int KP, int KPL, int NPW>
__global__ void __launch_bounds__(TRAJ_THREADS, 2) traj_kernel(ParamsView pv, PsetPack pk, TrajCfg cfg, ChebCfg cc,
                                                            double* __restrict__ tab) {
    extern __shared__ __align__(16) double smem[];
    constexpr int NPROD = 32 * NPW, NCONS = TRAJ_THREADS - NPROD;   // integrating / consumer threads
    static_assert(NPW == 1 || (NPW == 2 && KPL == 1), "two integrating warps own one type per lane");
    const int p = cfg.p_base + blockIdx.x;
    const int K = pv.K;
    if (cfg.only_missing && ((p >= cfg.own_lo && p < cfg.own_hi) || !cc.need_sweep[p] || (pk.pstatus[p] & ST_SOLVER))) return;
#ifdef GCDYN_TRAJ_TRACE
    __shared__ long long g_trace[16];
#endif
    TRACE_STAMP(0);
#ifdef GCDYN_TRAJ_TRACE
    if (threadIdx.x == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_trace[14] = gt; }
#endif
    __shared__ __align__(8) uint64_t s_full[RING_SLOTS], s_empty[RING_SLOTS];   // ring of cells → consumer warps
    __shared__ int s_flag;                              // fallback reasons collected while the integration runs
    __shared__ int s_pst;                               // the pset's status, for the tail (no trip through global memory)
    const bool cheb = cc.D > 0;                         // CTA-uniform
    const int KSB = KPL == 1 ? KP : ((K + 3) & ~3);     // broadcast buffer length (multiple of 4)
    const size_t traj_doubles = 2 * (size_t)KSB + (KPL == 1 ? 0 : (((size_t)K * K + 1) & ~(size_t)1));
    constexpr int SLOTS = ring_slots(KPL);              // cells in flight between the integrating and the consumer warps
    const TrajSmem ts(smem + traj_doubles, K, cheb ? cc.D : 0, M, SLOTS);
    // The integrating warp is chosen from the hardware warp slots so that CTAs sharing an SM integrate on different
    // sub-partitions (two integrations on one FP64 pipe: +32 % per Taylor stage, tools/traj_micro.cu).  Slot groups
    // of four are the usual allocation for a 4-warp CTA; any other allocation only loses the benefit.
    __shared__ int s_wid[TRAJ_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        unsigned wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        if (lane == 0) s_wid[warp] = (int)wid;
    }
    if (threadIdx.x == 0) {
        s_flag = 0;
        if (blockIdx.x == 0 && cheb) *cc.nflag = 0;
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, NCONS); }
    }
    // fused Metropolis iteration: six threads draw the six coordinates of chain p's proposal and write its parameters,
    // which everybody reads behind the barrier
    unsigned long long mh_it = 0;
    if (cc.mh_on) {
        mh_it = *cc.mh.it_dev + 1;
        int in = 1;
        if (threadIdx.x < MH_DIM) in = mh_propose_coord(cc.mh, mh_it, p, threadIdx.x) ? 1 : 0;
        const int all_in = __syncthreads_and(in);
        if (threadIdx.x == 0) cc.mh.inside[p] = all_in;
    } else {
        __syncthreads();
    }
#ifndef GCDYN_EXP_LATE_PDL
    pdl_launch_dependents();                            // the contraction's CTAs may become resident behind this grid
#endif
    // roles: prod_rank = 0..NPW−1 for the integrating warps (−1 otherwise), cons_rank = 0.. for the others
    int prod_rank = -1, cons_rank = -1;
    {
        int lo = s_wid[0];
#pragma unroll
        for (int w = 1; w < TRAJ_THREADS / 32; ++w) lo = min(lo, s_wid[w]);
        bool isp[TRAJ_THREADS / 32];
        int np = 0;
        if (NPW == 1) {
            int pw = 0;
            const int want = (lo >> 2) & 3;
#pragma unroll
            for (int w = TRAJ_THREADS / 32 - 1; w >= 0; --w) if ((s_wid[w] & 3) == want) pw = w;
#pragma unroll
            for (int w = 0; w < TRAJ_THREADS / 32; ++w) isp[w] = (w == pw);
            np = 1;
        } else {
            // two CTAs on an SM: one integrates on sub-partitions {0,1}, the other on {2,3}
            const int want = ((lo >> 2) & 1) * 2;
#pragma unroll
            for (int w = 0; w < TRAJ_THREADS / 32; ++w) { isp[w] = ((s_wid[w] & 3) >> 1) == (want >> 1); np += isp[w] ? 1 : 0; }
            if (np != NPW) {                            // irregular slot allocation: any two warps will do
#pragma unroll
                for (int w = 0; w < TRAJ_THREADS / 32; ++w) isp[w] = w < NPW;
            }
        }
        int pr = 0, cr = 0;
#pragma unroll
        for (int w = 0; w < TRAJ_THREADS / 32; ++w) {
            if (w == warp) { if (isp[w]) prod_rank = pr; else cons_rank = cr; }
            if (isp[w]) ++pr; else ++cr;
        }
    }
    if (prod_rank < 0) {
        const int ct = cons_rank * 32 + lane;
        prep_pset(pv, pk, p, 0, ct, NCONS, false, !(cheb && cc.compact_tc));
        cell_consume<M, SLOTS>(pv, cfg, cc, cheb, p, tab, ts, s_full, s_empty, ct, NCONS, pk, &s_flag);
#ifdef GCDYN_TRAJ_TRACE
        if (ct == 0) g_trace[6] = clock64();
#endif
    }
    if (prod_rank >= 0) {
    const int pt = prod_rank * 32 + lane;               // index among the integrating threads
    // exchange between the integrating warps (NPW == 2): named barrier 2 and a double-buffered pair of slots
    __shared__ unsigned s_xch[2][2][2];
    int xtog = 0;
    auto prod_sync = [&]() {
        if (NPW == 1) __syncwarp();
        else asm volatile("bar.sync 2, %0;" ::"r"(NPROD) : "memory");
    };
    auto prod_max2 = [&](unsigned& v0, unsigned& v1) {      // v0, v1 are already warp-uniform
        if (NPW == 1) return;
        if (lane == 0) { s_xch[xtog][0][prod_rank] = v0; s_xch[xtog][1][prod_rank] = v1; }
        asm volatile("bar.sync 2, %0;" ::"r"(NPROD) : "memory");
        v0 = max(s_xch[xtog][0][0], s_xch[xtog][0][1]);
        v1 = max(s_xch[xtog][1][0], s_xch[xtog][1][1]);
        xtog ^= 1;
    };
    double* cb = smem;                                  // 2 × KSB
    double* GT = smem + 2 * KSB;                        // KPL > 1 only: GT[j*K + i] = δ Γ[i,j]

    // ---- set-up.  Every global load is issued before the first value is used, so that a cold cache costs ONE memory
    //      round trip, not one per dependent group (scalars → birth rates → diagonal → row of δΓ).
    const double* G = pv.Gamma + (size_t)p * pv.gamma_stride;
    const double mu = pv.mu[p], delta = pv.delta[p], rho = pv.rho[p], sigma = pv.sigma[p];
    bool own[KPL];
    double in0[KPL], gdiag[KPL];                        // birth-rate input (λ itself, or the type value) and Γ[i,i]
    double rs0 = 0.0, rs1 = 0.0, rs2 = 0.0, rs3 = 0.0;  // sigmoid parameters
    if (pv.response == 2) { rs0 = pv.xscale[p]; rs1 = pv.xshift[p]; rs2 = pv.yscale[p]; rs3 = pv.yshift[p]; }
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int i = pt + NPROD * q;
        own[q] = i < K;
        const int ii = own[q] ? i : 0;
        in0[q] = pv.response == 0 ? pv.lambda[(size_t)p * K + ii] : (pv.response == 1 ? pv.lambda[p] : pv.type_space[ii]);
        gdiag[q] = G[ii + (size_t)ii * K];
    }
    double g[KPL == 1 ? KP : 1];
    if constexpr (KPL == 1) {
        const int ii = own[0] ? pt : 0;
#pragma unroll
        for (int j = 0; j < KP; ++j) g[j] = G[ii + (size_t)(j < K ? j : 0) * K];
    }
    const double b = mu * (1.0 - sigma);
    double lam[KPL], a[KPL];
    double rate = 0.0;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        double l = in0[q];                              // λ(model, type), mbp.jl:27-43
        if (pv.response == 2) l = rs2 * (1.0 / (1.0 + exp(-(rs0 * (in0[q] - rs1))))) + rs3;
        lam[q] = own[q] ? l : 0.0;
        const double gam = own[q] ? delta * -gdiag[q] : 0.0;
        a[q] = lam[q] + mu;
        const double r = own[q] ? fabs(lam[q]) + fabs(mu) + 2.0 * fabs(gam) : 0.0;
        if (!(r <= rate)) rate = r;
    }
    if constexpr (KPL == 1) {
#pragma unroll
        for (int j = 0; j < KP; ++j) g[j] = (own[0] && j < K) ? delta * g[j] : 0.0;
    } else {
        for (int e = lane; e < K * K; e += 32) GT[e] = delta * G[e];   // G[i + j*K] → GT[j*K + i]
    }
    rate = warp_max_nan(rate);
    if (NPW > 1) {                                      // non-negative doubles (and NaN) order like their bit patterns
        unsigned rh = (unsigned)__double2hiint(fabs(rate)), rl = (unsigned)__double2loint(fabs(rate));
        // compare (hi, lo) lexicographically through two rounds: the hi words first
        unsigned h0 = rh, dummy = 0u;
        prod_max2(h0, dummy);
        unsigned l0 = (rh == h0) ? rl : 0u;
        dummy = 0u;
        prod_max2(l0, dummy);
        rate = __hiloint2double((int)h0, (int)l0);
    }
    for (int e = pt; e < 2 * KSB; e += NPROD) cb[e] = 0.0;
    prod_sync();

    const bool adaptive = cfg.S_fixed <= 0;
    double pcur[KPL];
#pragma unroll
    for (int q = 0; q < KPL; ++q) pcur[q] = own[q] ? 1.0 - rho : 0.0;   // p(0) = 1 − ρ, mbp.jl:130,141,226
    // widest cell the controller may take: beyond h·rate ≈ M/3 the last two Taylor terms stop being a usable estimate
    double hcap = hrate_cap(M) / rate;
    if (!(hcap <= cfg.T)) hcap = cfg.T;                 // also rate == 0, NaN
    double tau = 0.0;
    // Width control works on the HIGH WORDS of |A_M| and |A_{M−1}| (non-negative doubles order like their bit patterns,
    // NaN sorts above +Inf): one integer REDUX each gives the maxima over the types, log2(x) ≈ (hi − 0x3ff00000)/2^20
    // (exponent + linear mantissa, never above the true value by construction of the bound below, short by < 0.0861),
    // so h = min((tol/4/max|A_M|)^(1/M), (tol/4/max|A_{M−1}|)^(1/(M−1))) keeps the indicator below 0.53·tol.  Every
    // lane sees the same values: decisions are warp-uniform and deterministic.
    const unsigned tol_hi = (unsigned)__double2hiint(cfg.tol);
    const float l2aim = (float)((int)(unsigned)__double2hiint(0.25 * cfg.tol) - 0x3ff00000) * (1.0f / 1048576.0f);
    unsigned errmax_hi = 0u;
    bool failed = false;
    int s = 0;
    TRACE_STAMP(2);
#ifdef GCDYN_TRAJ_TRACE
    long long tr_wait = 0, tr_stage = 0, tr_epi = 0, tr_t = clock64();
#define TRACE_ACC(v) do { const long long t__ = clock64(); v += t__ - tr_t; tr_t = t__; } while (0)
#else
#define TRACE_ACC(v) do { } while (0)
#endif
    for (;;) {
        const int slot_i = s % SLOTS;
        if (s >= cfg.S_cap) { failed = true; break; }
        if (s >= SLOTS) mbar_wait(s_empty + slot_i, (uint32_t)(s / SLOTS - 1) & 1u);   // consumers left it
        TRACE_ACC(tr_wait);
        double* slot = ts.ring + (size_t)slot_i * (M + 1) * ts.KS;
        double A[KPL][M + 1];
        double dn[KPL], kappa[KPL], seed[KPL];
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            A[q][0] = dn[q] = pcur[q];
            kappa[q] = fma(2.0 * lam[q], dn[q], -a[q]);
            seed[q] = fma(dn[q], fma(lam[q], dn[q], -a[q]), b);       // stage 0: λ d_0² − a d_0 + b
        }
#pragma unroll
        for (int n = 0; n < M; ++n) {
            double* buf = cb + (n & 1) * KSB;
#pragma unroll
            for (int q = 0; q < KPL; ++q)
                if (own[q]) buf[pt + NPROD * q] = dn[q];
            prod_sync();
            double dnew[KPL];
            if constexpr (KPL == 1) {
                double d0 = seed[0], d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                for (int j = 0; j < KP; j += 4) {
                    const double2 v01 = *reinterpret_cast<const double2*>(buf + j);
                    const double2 v23 = *reinterpret_cast<const double2*>(buf + j + 2);
                    d0 = fma(g[j], v01.x, d0);
                    d1 = fma(g[j + 1], v01.y, d1);
                    d2 = fma(g[j + 2], v23.x, d2);
                    d3 = fma(g[j + 3], v23.y, d3);
                }
                if (own[0]) slot[(size_t)n * ts.KS + pt] = A[0][n];      // publication rides behind the loads
                dnew[0] = (d0 + d1) + (d2 + d3);
            } else {
                double d[KPL][2];
#pragma unroll
                for (int q = 0; q < KPL; ++q) { d[q][0] = seed[q]; d[q][1] = 0.0; }
                int irow[KPL];
#pragma unroll
                for (int q = 0; q < KPL; ++q) irow[q] = own[q] ? pt + NPROD * q : 0;
                int j = 0;
#pragma unroll 4
                for (; j + 1 < K; j += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(buf + j);
                    const double* r0 = GT + (size_t)j * K;
#pragma unroll
                    for (int q = 0; q < KPL; ++q) {
                        d[q][0] = fma(r0[irow[q]], v.x, d[q][0]);
                        d[q][1] = fma(r0[K + irow[q]], v.y, d[q][1]);
                    }
                }
                if (j < K) {
                    const double v0 = buf[j];
#pragma unroll
                    for (int q = 0; q < KPL; ++q) d[q][0] = fma(GT[(size_t)j * K + irow[q]], v0, d[q][0]);
                }
#pragma unroll
                for (int q = 0; q < KPL; ++q) {
                    if (own[q]) slot[(size_t)n * ts.KS + pt + NPROD * q] = A[q][n];
                    dnew[q] = d[q][0] + d[q][1];
                }
            }
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                // λ (n+1)! R_{n+1}: needs A_1..A_n only, i.e. nothing of this stage — it fills the broadcast's shadow
                double add = 0.0;
                if (n >= 1) {
                    const int k = n + 1;
                    double r0 = 0.0, r1 = 0.0;
#pragma unroll
                    for (int m = 1; 2 * m < k; ++m) {
                        if (m & 1) r1 = fma(A[q][m], A[q][k - m], r1);
                        else r0 = fma(A[q][m], A[q][k - m], r0);
                    }
                    double R = 2.0 * (r0 + r1);
                    if ((k & 1) == 0) R = fma(A[q][k / 2], A[q][k / 2], R);
                    add = (lam[q] * factd(k)) * R;
                }
                dn[q] = dnew[q];
                A[q][n + 1] = dnew[q] * (1.0 / factd(n + 1));
                seed[q] = fma(kappa[q], dnew[q], add);
            }
        }
#pragma unroll
        for (int q = 0; q < KPL; ++q)
            if (own[q]) slot[(size_t)M * ts.KS + pt + NPROD * q] = A[q][M];
        TRACE_ACC(tr_stage);
        // ---- cell width from the last two coefficients
        unsigned hiM = 0u, hiM1 = 0u;
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            if (own[q]) {
                hiM = max(hiM, (unsigned)__double2hiint(fabs(A[q][M])));
                hiM1 = max(hiM1, (unsigned)__double2hiint(fabs(A[q][M - 1])));
            }
        }
        hiM = __reduce_max_sync(FULL, hiM);
        hiM1 = __reduce_max_sync(FULL, hiM1);
        prod_max2(hiM, hiM1);
        double h;
        bool last = false;
        if (adaptive) {
            const float lM = (float)((int)hiM - 0x3ff00000) * (1.0f / 1048576.0f);
            const float lM1 = (float)((int)hiM1 - 0x3ff00000) * (1.0f / 1048576.0f);
            float lh = fminf((l2aim - lM) * (1.0f / (float)M), (l2aim - lM1) * (1.0f / (float)(M - 1)));
            lh = fminf(lh, 100.0f);
            h = (double)exp2f(lh);
            if (!(h > cfg.T * 1e-10)) { failed = true; break; }          // non-finite coefficients land here too
            if (!(h <= hcap)) h = hcap;
            const double rem = cfg.T - tau;
            if (h >= rem * (1.0 - 1e-3)) { h = rem; last = true; }
            else if (h > 0.5 * rem) h = 0.5 * rem;                       // two even cells instead of a sliver at the end
        } else {
            last = s + 1 >= cfg.S_fixed;
            h = last ? cfg.T - tau : cfg.T / (double)cfg.S_fixed;
        }
        // ---- p(τ_s + h); outside [0, 1]: shrink the cell and re-evaluate (isoutofdomain, mbp.jl:146,204,231)
        double pn[KPL];
        for (int shrinks = 0;; ++shrinks) {
            bool bad = false;
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                pn[q] = estrin<M>(A[q], h);
                if (own[q] && !(pn[q] >= -1e-9 && pn[q] <= 1.0 + 1e-9)) bad = true;
            }
            unsigned anyb = __any_sync(FULL, bad) ? 1u : 0u, zero = 0u;
            prod_max2(anyb, zero);
            if (!anyb) break;
            h *= 0.5;
            last = false;
            if (!adaptive || shrinks >= 60 || !(h > cfg.T * 1e-10)) { failed = true; break; }
        }
        if (failed) break;
        // error indicator of the accepted cell (diagnostics; decides success on a fixed grid)
        {
            const double hM1 = powi<M - 1>(h), hM = hM1 * h;
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                const double ind = fma(fabs(A[q][M]), hM, fabs(A[q][M - 1]) * hM1);
                if (own[q]) errmax_hi = max(errmax_hi, (unsigned)__double2hiint(ind));
            }
        }
        // ---- publish the cell: A_n are in the slot already; τ_s, τ_{s+1}, h and the flags follow
        if (pt == 0) {
            double* meta = ts.meta + slot_i * 4;
            meta[0] = tau; meta[1] = last ? cfg.T : tau + h; meta[2] = h; meta[3] = last ? (double)CELL_LAST : 0.0;
        }
        prod_sync();
        if (pt == 0) mbar_arrive(s_full + slot_i);
#pragma unroll
        for (int q = 0; q < KPL; ++q) pcur[q] = pn[q];
        ++s;
        TRACE_ACC(tr_epi);
        if (last) break;
        tau += h;
    }
#ifdef GCDYN_TRAJ_TRACE
    if (lane == 0) { g_trace[7] = tr_wait; g_trace[8] = tr_stage; g_trace[9] = tr_epi; }
#endif
    if (failed) {                                       // tell the consumers to stop (the regular exit flagged its last cell)
        const int slot_i = s % SLOTS;
        if (s >= SLOTS) mbar_wait(s_empty + slot_i, (uint32_t)(s / SLOTS - 1) & 1u);
        if (pt == 0) { ts.meta[slot_i * 4 + 3] = (double)CELL_FAIL; }
        prod_sync();
        if (pt == 0) mbar_arrive(s_full + slot_i);
    }
    TRACE_STAMP(3);
    // conditioning term and diagnostics
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int i = pt + NPROD * q;
        if (own[q]) {
            const double lc = log1p(-pcur[q]);                         // log(1 − p_i(T)), mbp.jl:244-245
            pk.logcond[(size_t)p * K + i] = lc;
            if (cheb) {                                                // its column of the coefficient row
                double v = -lc;
                if (!isfinite(v)) {
                    if (cc.mv.colsum[K + 2 + i] != 0.0) atomicOr(&s_flag, 2);
                    v = 0.0;
                }
                cc.A[(size_t)p * cc.mv.Jst + K + 2 + i] = v;
            }
        }
    }
    unsigned emax = __reduce_max_sync(FULL, errmax_hi), zero = 0u;
    prod_max2(emax, zero);
    const bool unresolved = failed || !(emax < tol_hi);
    if (pt == 0) {
        pk.perr[p] = __hiloint2double((int)min(emax + 1u, 0x7ff00000u), 0);   // upper bound of the largest accepted indicator
        pk.S[p] = s > 0 ? s : 1;
        pk.pstatus[p] = unresolved ? ST_SOLVER : 0;
        s_pst = unresolved ? ST_SOLVER : 0;
    }
    }   // integrating warps
#ifdef GCDYN_EXP_LATE_PDL
    pdl_launch_dependents();
#endif
    if (!cheb) return;
    __syncthreads();                                    // node values, status of this pset are complete and visible
    TRACE_STAMP(4);
    nodal_finish(cc, p, ts, &s_flag, s_pst, mh_it);
    if (cc.push.n > 0) {                                // CTA-uniform
        __syncthreads();                                // every column of the row has been written
        const double2* src = reinterpret_cast<const double2*>(cc.A + (size_t)p * cc.mv.Jst);
        const size_t at = (size_t)p * (size_t)(cc.mv.Jst / 2);
        for (int i0 = threadIdx.x; i0 < cc.push.w2; i0 += 4 * TRAJ_THREADS) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * TRAJ_THREADS;
                if (i < cc.push.w2) v[u] = __ldcg(src + i);
            }
            for (int r = 0; r < cc.push.n; ++r) {
                double2* dst = reinterpret_cast<double2*>(cc.push.dst[r]) + at;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * TRAJ_THREADS;
                    if (i < cc.push.w2) dst[i] = v[u];
                }
            }
        }
        __threadfence_system();                         // performed at the peers before this CTA counts as finished
    }
#ifdef GCDYN_TRAJ_TRACE
    __syncthreads();
    TRACE_STAMP(5);
    if (threadIdx.x == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_trace[15] = gt; }
    if (threadIdx.x == 0 && p < 1024) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        long long* o = g_trace_out + (size_t)p * 24;
        for (int i = 0; i < 16; ++i) o[i] = g_trace[i];
        o[16] = smid; o[18] = pk.S[p];
    }
#endif
}

template <int M, bool SMEM>
__device__ __forceinline__ double tree_items_sum(const ForestView& fv, const double* __restrict__ tp,
                                                 const double* __restrict__ nt, int K, int S, int S_cap, int grid_doubles,
                                                 double T, int64_t i0, int64_t i1, int lane) {
    // tp = the pset's block: [τ_s | 1/h_s | pad][rows]
    constexpr int L = M + 2;
    double acc0 = 0.0, acc1 = 0.0;
    for (int64_t i = i0 + lane; i < i1; i += 64) {
        const int64_t j = i + 32;
        const bool has2 = j < i1;
        const double tm0 = fv.it_time[i];
        const int2 me0 = fv.it_meta[i];
        const double tm1 = has2 ? fv.it_time[j] : tm0;
        const int2 me1 = has2 ? fv.it_meta[j] : me0;
        const double ta0 = T - tm0, ta1 = T - tm1;
        const int s0 = find_cell(tp, S, ta0), s1 = find_cell(tp, S, ta1);
        const double* invh = tp + (S_cap + 1);
        const double* rows = tp + grid_doubles;
        const double f0 = eval_row<M, SMEM>(rows + ((size_t)s0 * K + (me0.x & 0xffffff)) * L, (ta0 - tp[s0]) * invh[s0]);
        const double f1 = eval_row<M, SMEM>(rows + ((size_t)s1 * K + (me1.x & 0xffffff)) * L, (ta1 - tp[s1]) * invh[s1]);
        acc0 = fma((double)(me0.x >> 24), f0, acc0);
        if (me0.y >= 0) acc0 += __ldg(nt + me0.y);
        if (has2) {
            acc1 = fma((double)(me1.x >> 24), f1, acc1);
            if (me1.y >= 0) acc1 += __ldg(nt + me1.y);
        }
    }
    return warp_sum(acc0 + acc1);
}

// ------------------------------------------------------------------------------------------------
// sweep_kernel: persistent CTAs, each owning a contiguous range of (pset, chunk) units in pset-major
// order, so the pset's table is (re)staged into shared memory only when the pset changes.
//   LL(t,p) = Σ_items [ w·F_type(τ) + nt[aux] ] + n_death·(log μ + log σ) + n_surv·log ρ − logcond[root_type]
// ------------------------------------------------------------------------------------------------
constexpr int SWEEP_THREADS = 768;   // 24 warps share one staged table; one CTA per SM

template <int M>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) sweep_kernel(ForestView fv, PsetPack pk, const double* __restrict__ tab,
                                                    long long tab_stride, int S_cap, int grid_doubles,
                                                    const int64_t* __restrict__ chunk_off,
                                                    int n_chunks, int P, double T, int smem_doubles,
                                                    const int* __restrict__ only_flagged,
                                                    double* __restrict__ out_tree, int32_t* __restrict__ status) {
    // only_flagged != nullptr: fallback mode — redo just the psets the Chebyshev/GEMM path could not take
    extern __shared__ __align__(128) double stab[];
    __shared__ __align__(8) uint64_t mbar;
    constexpr int L = M + 2;
    const int K = fv.K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const long long U = (long long)P * n_chunks;
    const long long u_lo = U * blockIdx.x / gridDim.x, u_hi = U * (blockIdx.x + 1) / gridDim.x;
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    __syncthreads();
    uint32_t parity = 0;
    int cur_p = -1, S = 1, pst = 0, flag_p = -1, flag_v = 0;
    bool in_smem = false;
    const double* tp = nullptr;
    const double* nt = nullptr;
    double lmusig = 0.0, lrho = 0.0;
    for (long long u = u_lo; u < u_hi; ++u) {
        const int p = (int)(u / n_chunks);
        const int ch = (int)(u - (long long)p * n_chunks);
        if (only_flagged) {                                  // CTA-uniform; one flag load per pset, not per unit
            if (p != flag_p) { flag_p = p; flag_v = only_flagged[p]; }
            if (!flag_v) { u = (long long)(p + 1) * n_chunks - 1; continue; }   // jump to the next pset's units
        }
        if (p != cur_p) {
            __syncthreads();                       // every warp is done with the previous table
            cur_p = p;
            S = pk.S[p];
            pst = pk.pstatus[p];
            nt = pk.nt + (size_t)p * pk.ntw;
            lmusig = nt[K + K * K];
            lrho = nt[K + K * K + 1];
            const double* src = tab + (size_t)p * tab_stride;
            const long long n = grid_doubles + (long long)S * K * L;      // grid block + the cells actually used
            in_smem = n <= smem_doubles && !(pst & ST_SOLVER);
            if (in_smem) {
                if (threadIdx.x == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // order prior generic reads before async writes
                    const uint32_t bytes = (uint32_t)(n * 8);
                    mbar_expect_tx(&mbar, bytes);
                    uint32_t off = 0;
                    while (off < bytes) {
                        const uint32_t len = bytes - off < 32768u ? bytes - off : 32768u;
                        bulk_g2s(reinterpret_cast<char*>(stab) + off, reinterpret_cast<const char*>(src) + off, len, &mbar);
                        off += len;
                    }
                }
                mbar_wait(&mbar, parity);
                parity ^= 1u;
                tp = stab;
            } else {
                tp = src;
            }
        }
        const int64_t t_lo = chunk_off[ch], t_hi = chunk_off[ch + 1];
        for (int64_t t = t_lo + warp; t < t_hi; t += nwarps) {
            const int64_t i0 = fv.item_off[t], i1 = fv.item_off[t + 1];
            double acc;
            if (pst & ST_SOLVER) acc = 0.0;
            else if (in_smem) acc = tree_items_sum<M, true>(fv, tp, nt, K, S, S_cap, grid_doubles, T, i0, i1, lane);
            else acc = tree_items_sum<M, false>(fv, tp, nt, K, S, S_cap, grid_doubles, T, i0, i1, lane);
            if (lane == 0) {
                const int rt = fv.root_type[t];
                double r = acc + count_times_log(fv.n_death[t], lmusig) + count_times_log(fv.n_surv[t], lrho)
                           - (rt >= 0 ? pk.logcond[(size_t)p * K + rt] : CUDART_NAN);
                int st = pst;
                if (fv.self_loop[t]) st |= ST_SELF_LOOP;
                if (st & (ST_SELF_LOOP | ST_SOLVER)) r = -CUDART_INF;
                else if (r != r) st |= ST_DOMAIN;
                out_tree[(size_t)p * fv.n_trees + t] = r;
                if (status) status[(size_t)p * fv.n_trees + t] = st;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// alt_kernel: the closed-form alternates, evaluated per node exactly as written.
//   method 1 naive (extra_likelihoods.jl:10-34), 2 stadler (:45-102), 3 stadler_unconditioned (:115-151)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) alt_kernel(ForestView fv, ParamsView pv, PsetPack pk,
                                                  const int64_t* __restrict__ chunk_off, double present_time,
                                                  int method, double* __restrict__ out_tree,
                                                  int32_t* __restrict__ status) {
    const int p = blockIdx.y;
    const int K = fv.K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const double* nt = pk.nt + (size_t)p * pk.ntw;
    const double* lamv = pk.lam + (size_t)p * K;
    const double* Lamv = pk.Lam + (size_t)p * K;
    const double mu = pv.mu[p], rho = pv.rho[p], sigma = pv.sigma[p];
    const double lrho = nt[K + K * K + 1];
    const int64_t t_lo = chunk_off[blockIdx.x], t_hi = chunk_off[blockIdx.x + 1];
    for (int64_t t = t_lo + warp; t < t_hi; t += nwarps) {
        const int64_t i0 = fv.tree_off[t], i1 = fv.tree_off[t + 1];
        double acc = 0.0;
        if (fv.self_loop[t]) {                  // log γ(x→x) = log 0 (extra_likelihoods.jl:21-23 with mbp.jl:84-97): −Inf, and the
            if (lane == 0) {                    // records behind the self-loop may hold placeholder indices
                out_tree[(size_t)p * fv.n_trees + t] = -CUDART_INF;
                if (status) status[(size_t)p * fv.n_trees + t] = 0;
            }
            continue;
        }
        for (int64_t i = i0 + lane; i < i1; i += 32) {
            const int ev = fv.event[i];
            const int x = fv.btype_idx[i];
            const double lam = lamv[x], Lam = Lamv[x];
            double term;
            if (method == 1) {
                const double dt = fv.t_end[i] - fv.t_start[i];
                const double theta = 1.0 / Lam;
                // Distributions.jl: logpdf(Exponential(θ), x) = −(log θ + x/θ) for x ≥ 0, −Inf otherwise
                const double lpdf = dt >= 0.0 ? -(log(theta) + dt / theta) : -CUDART_INF;
                if (ev == EV_BIRTH)        term = lpdf + log(lam / Lam);
                else if (ev == EV_SDEATH)  term = lpdf + log(mu / Lam);
                else if (ev == EV_TC)      term = lpdf + nt[K + x * K + fv.type_idx[i]] - log(Lam);
                else {
                    // log(1 − cdf(Exponential(θ), x)), cdf = −expm1(−max(x,0)/θ): kept as written
                    const double z = fmax(dt, 0.0) / theta;
                    const double lccdf = log(1.0 - (-expm1(-z)));
                    term = lccdf + (ev == EV_SSURV ? lrho : log(1.0 - rho));
                }
            } else {
                const size_t o = (size_t)p * K + x;
                const double c = pk.st_c[o], A = pk.st_A[o], B = pk.st_B[o];
                const double ts = present_time - fv.t_start[i], te = present_time - fv.t_end[i];
                const double hs = A * exp(-c * ts) - B, he = A * exp(-c * te) - B;
                term = c * (te - ts) + 2.0 * (log(he) - log(hs));
                if (ev == EV_BIRTH)        term += nt[x];
                else if (ev == EV_SDEATH)  term += log(sigma) + log(mu);
                else if (ev == EV_TC)      term += nt[K + x * K + fv.type_idx[i]];
                else                       term += lrho;
            }
            acc += term;
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            double r = acc;
            if (method == 2) {
                const int rt = fv.root_type[t];
                r -= rt >= 0 ? pk.logcond[(size_t)p * K + rt] : CUDART_NAN;
            }
            int st = 0;
            if (r != r) st |= ST_DOMAIN;
            out_tree[(size_t)p * fv.n_trees + t] = r;
            if (status) status[(size_t)p * fv.n_trees + t] = st;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// reduce_kernel: out_pset[p] = Σ_t out_tree[p*T + t], one CTA per pset, fixed order (deterministic).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reduce_kernel(const double* __restrict__ out_tree, int64_t T,
                                                     double* __restrict__ out_pset) {
    __shared__ double part[8];
    const int p = blockIdx.x;
    const double* row = out_tree + (size_t)p * T;
    double acc = 0.0;
    for (int64_t t = threadIdx.x; t < T; t += blockDim.x) acc += row[t];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
        out_pset[p] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// event_count_kernel: per-tree histogram of event codes (integer parity with the reference's events).
// ------------------------------------------------------------------------------------------------
__global__ void event_count_kernel(ForestView fv, int64_t* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= fv.n_trees) return;
    int cnt[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = fv.tree_off[t] + lane; i < fv.tree_off[t + 1]; i += 32) {
        const int ev = fv.event[i];
#pragma unroll
        for (int e = 0; e < 7; ++e) cnt[e] += (ev == e);
    }
#pragma unroll
    for (int e = 0; e < 7; ++e) {
        int v = cnt[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) out[t * 7 + e] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// fp64 peak probe: 8 independent DFMA chains per thread, register resident.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456) out[0] = r;   // never true; keeps the loop alive
}

// dependent-chain DFMA latency probe: one warp, one chain; cycles per DFMA = (clock delta) / iters
__global__ void dfma_latency_kernel(double* out, long long* cycles, int iters, double seed) {
    double a = seed;
    const double m = 1.0000001, c = 1e-9;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, m, c);
    const long long t1 = clock64();
    if (threadIdx.x == 0) { cycles[0] = t1 - t0; out[0] = a; }
}

}  // namespace gcdyn
