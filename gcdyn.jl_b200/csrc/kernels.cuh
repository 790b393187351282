// Device kernels of libgcdyn_b200 (sm_100a, FP64).  See DESIGN.md §3-§5.
//
// Pipeline of one batched ODE likelihood evaluation (P parameter sets × T trees):
//   traj_kernel   one CTA of four warps per pset.  One warp (two for 32 < K ≤ 64) integrates the extinction ODE
//                 (dp_dt!, mbp.jl:266-278) with a fixed-order Taylor series on an adaptive grid of cells and
//                 publishes every accepted cell through a shared-memory ring; the other warps compute rates and
//                 node-term logs (λ, μ, γ of mbp.jl:27-97), build per cell and type the polynomial of
//                 F_x(τ) = ∫(−Λ_x + 2 λ_x p_x) — the branch density of dp_logq_dt!, mbp.jl:285-300 — write that
//                 table, and evaluate F at the Chebyshev nodes; the CTA then produces the pset's row of GEMM
//                 coefficients (cheb_gemm.cuh) and log(1 − p_x(T)) (mbp.jl:244-245)
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
__device__ __forceinline__ void prep_pset(const ParamsView& pv, const PsetPack& pk, int p, int method, int tid, int nthr,
                                          bool init_status = true) {
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
    for (int e = tid; e < K * K; e += nthr) {
        const int x = e / K, y = e - x * K;
        nt[K + e] = (x == y) ? -CUDART_INF : log(delta * G[x + (size_t)y * K]);   // γ(model, from, to), mbp.jl:84-97
    }
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
    int C1;                // 2K + 4: start of the Chebyshev block
    int Dcap;              // moments exist for d = 0..Dcap
    long long Jtc;         // start of the K² type-change block
    long long Jst;         // row stride of Mt and A (doubles), multiple of 4
    const double* Mt;      // [T][Jst]
    const double* colsum;  // [Jst] forest-level column sums (only count columns are consulted)
};

struct ChebCfg {
    double T;
    double tol;            // accepted Chebyshev tail (absolute, on F)
    int D;                 // interpolation degree (≤ Dcap), even
    int compact_tc;        // 1: Γ shared — log δ column + per-tree constant; 0: K² type-change columns
    int S_cap, grid_doubles;   // layout of a pset's table block: [τ_s | 1/h_s | pad][rows]
    long long tab_stride;
    MomentView mv;
    double* A;             // [P][Jst] coefficient rows of the GEMM
    int* need_sweep;       // [P] fallback flags
    int* d_common;         // [1] max effective degree over the psets the GEMM takes
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
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// Chebyshev stage of one pset inside traj_kernel.  While warp 0 integrates, warps 1..3 (`cheb_consume`) build
// the cosine tables and evaluate F_x at the D+1 Chebyshev–Gauss–Lobatto nodes cell by cell, as soon as warp 0
// publishes a cell (release/acquire on a shared-memory counter; table rows are read back through L2).  After the
// CTA barrier `cheb_finish` runs on all threads: DCT-I → a[d][x]; effective degree; count-column coefficients;
// fallback flag.
// ------------------------------------------------------------------------------------------------
#ifdef GCDYN_TRAJ_TRACE
__device__ long long g_trace_out[1024 * 24];
#define TRACE_STAMP(i) do { if ((threadIdx.x & 31) == 0) g_trace[i] = clock64(); } while (0)
#define TRACE_ARG , long long* g_trace
#define TRACE_PASS , g_trace
#else
#define TRACE_STAMP(i) do { } while (0)
#define TRACE_ARG
#define TRACE_PASS
#endif

constexpr int CHEB_RING_SLOTS = 3;   // cells in flight between the integrating warp and the consumer warps
constexpr int CELL_LAST = 1, CELL_FAIL = 2;   // flags of a published cell

struct ChebSmem {                      // carve-up of the stage's shared-memory scratch
    double* costab;                    // [2D]  cos(π m / D)
    double* V;                         // [(D+1)*K] node values, later the coefficients a[d][x]
    double* EO;                        // [2][(H+1)*K] folded node values
    double* cosm;                      // [H+1][D+4]  cos(π k d / D), zero for d > D
    double* sgrid;                     // [S_cap+1] τ_s as published by warp 0
    double* shc;                       // [S_cap]   h_s
    int* sflag;                        // [S_cap+1] CELL_LAST / CELL_FAIL of a published cell
    double* ring;                      // [CHEB_RING_SLOTS][K*L] Taylor coefficients ĉ_0..ĉ_M of the cells in flight
    double* rowbuf;                    // [2][K*L] table rows of the cell being evaluated at the Chebyshev nodes
    __device__ ChebSmem(double* sm, int K, int D, int S_cap, int L) {
        const int H = D / 2;
        costab = sm;
        V = costab + 2 * D;
        if (D > 0) {
            cosm = V + ((((size_t)(D + 1) * K) + 1) & ~(size_t)1);
            sgrid = cosm + (size_t)(H + 1) * (D + 4);
        } else {                                        // no Chebyshev stage: only the ring and its grid/flags exist
            cosm = sgrid = sm;
        }
        shc = sgrid + (S_cap + 1);
        sflag = reinterpret_cast<int*>(sgrid + 2 * S_cap + 1);
        ring = sgrid + 2 * (S_cap + 1) + (((size_t)S_cap + 4) / 2 & ~(size_t)1);   // even offset: 16-byte aligned rows
        rowbuf = ring + (size_t)CHEB_RING_SLOTS * K * L;
        EO = ring;     // the folded halves live where the ring and the row buffers were: those are dead once the
                       // trajectory is complete (the host sizes the region as the larger of the two uses)
    }
};

// The three warps that do not integrate (ct = 0..nct-1).  For every cell the integrating warp publishes (ĉ_0..ĉ_M per
// type, τ_s, h_s) they build the table rows of F — row[0] = F(τ_s), row[1] = h(2λĉ_0 − Λ), row[n+1] = 2hλ ĉ_n/(n+1) —
// keep the running F(τ_s) per type, write rows and grid to the table in global memory and, when the Chebyshev stage
// is on, evaluate F_x at the Chebyshev–Gauss–Lobatto nodes inside the cell.  Node k sits at τ_k = T(1 + cos(πk/D))/2,
// so cells (increasing τ) meet the nodes in decreasing k; a node belongs to the largest cell with τ_s ≤ τ_k
// (find_cell's rule), the last cell takes the rest.
template <int M>
__device__ __forceinline__ void cell_consume(const ParamsView& pv, const TrajCfg& tc, const ChebCfg& cfg, bool cheb, int p,
                                             double* __restrict__ tab, const ChebSmem& cs, uint64_t* full, uint64_t* empty,
                                             int ct, int nct) {
    constexpr int L = M + 2;
    constexpr int TPT = 2;                              // types per consumer thread: K ≤ 128 (KPL ≤ 4), nct = 96
    const int K = pv.K, D = cfg.D, H = D / 2, CW = D + 4;
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
        for (int m = ct; m < 2 * D; m += nct) cs.costab[m] = cospi((double)m / (double)D);
        asm volatile("bar.sync 1, %0;" ::"r"(nct) : "memory");
        for (int idx = ct; idx < (H + 1) * CW; idx += nct) {
            const int k = idx / CW, d = idx - k * CW;
            cs.cosm[idx] = d <= D ? cs.costab[(k * d) % (2 * D)] : 0.0;
        }
    }
    double* tabp = tab + (size_t)p * tc.tab_stride;
    double* gridp = tabp;
    double* invhp = tabp + (tc.S_cap + 1);
    double* rowsp = tabp + tc.grid_doubles;
    int kn = D;
    for (int sc = 0;; ++sc) {
        const int slot_i = sc % CHEB_RING_SLOTS;
        mbar_wait(full + slot_i, (uint32_t)(sc / CHEB_RING_SLOTS) & 1u);
        const int flag = cs.sflag[sc];
        if (flag & CELL_FAIL) {
            if (ct == 0) gridp[sc] = tc.T;
            break;
        }
        const double t0 = cs.sgrid[sc], t1 = cs.sgrid[sc + 1], hc = cs.shc[sc];
        const double ih = 1.0 / hc;
        const double* slot = cs.ring + (size_t)slot_i * K * L;
        double* rb = cs.rowbuf + (size_t)(sc & 1) * K * L;
#pragma unroll
        for (int j = 0; j < TPT; ++j) {
            const int x = ct + j * nct;
            if (x < K) {
                double c[L], row[L];
#pragma unroll
                for (int n = 0; n < L; n += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(slot + (size_t)x * L + n);
                    c[n] = v.x;
                    c[n + 1] = v.y;
                }
                row[0] = Fc[j];
                row[1] = hc * (2.0 * lamx[j] * c[0] - Lamx[j]);
                const double h2l = 2.0 * hc * lamx[j];
                double fs[4] = {0.0, 0.0, 0.0, 0.0};         // smallest terms first, four partial sums
#pragma unroll
                for (int n = M; n >= 1; --n) {
                    row[n + 1] = (h2l * (1.0 / (double)(n + 1))) * c[n];
                    fs[n & 3] += row[n + 1];
                }
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
            else while (klo >= 0 && cfg.T * (1.0 + cs.costab[klo]) * 0.5 < t1) --klo;
            const int nitems = (kn - klo) * K;
            for (int it = ct; it < nitems; it += nct) {
                const int k = kn - it / K, x = it % K;
                const double tau = cfg.T * (1.0 + cs.costab[k]) * 0.5;
                cs.V[k * K + x] = eval_row<M, true>(rb + (size_t)x * L, (tau - t0) * ih);
            }
            kn = klo;
        }
        if (flag & CELL_LAST) break;
    }
}

template <int M>
__device__ __forceinline__ void cheb_finish(const ParamsView& pv, const PsetPack& pk, const ChebCfg& cfg, int p,
                                            const ChebSmem& cs TRACE_ARG) {
    const MomentView& mv = cfg.mv;
    double* __restrict__ A = cfg.A;
    int* __restrict__ need_sweep = cfg.need_sweep;
    int* __restrict__ d_common = cfg.d_common;
    const int K = mv.K, C1 = mv.C1, D = cfg.D;
    const int H = D / 2, CW = D + 4, NV = (D + 1) * K;
    double* V = cs.V;
    double* Acoef = V;
    double* EO = cs.EO;
    const double* cosm = cs.cosm;
    __shared__ int s_deff, s_flag;
    const int pst = pk.pstatus[p];
    if (threadIdx.x == 0) { s_deff = 8; s_flag = (pst & ST_SOLVER) ? 1 : 0; }
    if (pst & ST_SOLVER) {                               // CTA-uniform: a failed trajectory has no usable nodes
        for (int idx = threadIdx.x; idx < NV; idx += blockDim.x) V[idx] = 0.0;
        __syncthreads();
    }
    TRACE_STAMP(6);
    // DCT-I with the symmetry cos(π(D−k)d/D) = (−1)^d cos(πkd/D): fold the node values into an even and an
    // odd half, E_k = w_k (V_k + V_{D−k}), O_k = w_k (V_k − V_{D−k}), k = 0..D/2, w_0 = w_{D/2} = 1/2.
    for (int idx = threadIdx.x; idx < (H + 1) * K; idx += blockDim.x) {
        const int k = idx / K, x = idx - k * K;
        const double w = (k == 0 || k == H) ? 0.5 : 1.0;
        const double a = V[k * K + x], b = V[(D - k) * K + x];
        EO[idx] = w * (a + b);
        EO[(H + 1) * K + idx] = w * (a - b);
    }
    __syncthreads();
    TRACE_STAMP(7);
    // a thread owns four consecutive degrees 4g..4g+3 of one type: even degrees contract E, odd degrees O
    {
        const int NG = D / 4 + 1;
        const double scale = 2.0 / (double)D;
        for (int w = threadIdx.x; w < NG * K; w += blockDim.x) {
            const int g = w / K, x = w - g * K;
            const double* E = EO + x;
            const double* O = EO + (size_t)(H + 1) * K + x;
            const double* cg = cosm + 4 * g;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 5
            for (int k = 0; k <= H; ++k) {
                const double e = E[k * K], o = O[k * K];
                const double2 c01 = *reinterpret_cast<const double2*>(cg + (size_t)k * CW);
                const double2 c23 = *reinterpret_cast<const double2*>(cg + (size_t)k * CW + 2);
                a0 = fma(e, c01.x, a0);
                a1 = fma(o, c01.y, a1);
                a2 = fma(e, c23.x, a2);
                a3 = fma(o, c23.y, a3);
            }
            const int d0 = 4 * g;
            Acoef[d0 * K + x] = (d0 == 0 || d0 == D) ? 0.5 * scale * a0 : scale * a0;
            if (d0 + 1 <= D) Acoef[(d0 + 1) * K + x] = scale * a1;
            if (d0 + 2 <= D) Acoef[(d0 + 2) * K + x] = scale * a2;
            if (d0 + 3 <= D) Acoef[(d0 + 3) * K + x] = (d0 + 3 == D) ? 0.5 * scale * a3 : scale * a3;
        }
    }
    __syncthreads();
    TRACE_STAMP(8);
    // effective degree (a thread per type): the largest d whose tail Σ_{k≥d}|a_k| exceeds tol.  The sum runs downwards
    // from D and stops at the first excess — for a resolved series that is after the D − d_eff smallest coefficients.
    // The series counts as unresolved if its last three terms are not small.
    for (int x = threadIdx.x; x < K; x += blockDim.x) {
        const double last3 = fabs(Acoef[D * K + x]) + fabs(Acoef[(D - 1) * K + x]) + fabs(Acoef[(D - 2) * K + x]);
        if (!(last3 <= cfg.tol)) atomicExch(&s_flag, 1);
        double tail = 0.0;
        int found = 8;
        for (int d = D; d > 8; --d) {
            tail += fabs(Acoef[d * K + x]);
            if (!(tail <= cfg.tol)) { found = d; break; }
        }
        if (found > 8) atomicMax(&s_deff, found);
    }
    __syncthreads();
    TRACE_STAMP(9);
    const int deff = s_deff;
    double* Arow = A + (size_t)p * mv.Jst;
    for (int idx = threadIdx.x; idx < (D + 1) * K; idx += blockDim.x) {
        Arow[C1 + idx] = Acoef[idx];          // all interpolation coefficients are valid; the GEMM stops at d_common
    }
    // the GEMM rounds its column ranges up to multiples of 4: keep the padding finite
    if (threadIdx.x < 4) {
        const long long j = C1 + (long long)(D + 1) * K + threadIdx.x;
        if (j < mv.Jtc) Arow[j] = 0.0;
        const long long j2 = mv.Jtc + (long long)K * K + threadIdx.x;
        if (j2 < mv.Jst) Arow[j2] = 0.0;
    }
    // count columns
    const double* nt = pk.nt + (size_t)p * pk.ntw;
    for (int c = threadIdx.x; c < C1; c += blockDim.x) {
        double v = 0.0;
        if (c < K) v = nt[c];                                         // log λ_x
        else if (c == K) v = nt[K + K * K];                           // log μ + log σ
        else if (c == K + 1) v = nt[K + K * K + 1];                   // log ρ
        else if (c < 2 * K + 2) v = -pk.logcond[(size_t)p * K + (c - K - 2)];
        else if (c == 2 * K + 2) v = cfg.compact_tc ? log(pv.delta[p]) : 0.0;
        if (!isfinite(v)) {
            if (mv.colsum[c] != 0.0) atomicExch(&s_flag, 1);
            v = 0.0;
        }
        Arow[c] = v;
    }
    if (!cfg.compact_tc) {
        for (int e = threadIdx.x; e < K * K; e += blockDim.x) {
            double v = nt[K + e];
            const bool diag = (e / K) == (e % K);          // self-loop column: trees using it are −Inf anyway
            if (!isfinite(v)) {
                if (!diag && mv.colsum[mv.Jtc + e] != 0.0) atomicExch(&s_flag, 1);
                v = 0.0;
            }
            Arow[mv.Jtc + e] = v;
        }
    }
    __syncthreads();
    TRACE_STAMP(10);
    if (threadIdx.x == 0) {
        need_sweep[p] = s_flag;
        if (!s_flag) atomicMax(d_common, deff);
    }
}

// ------------------------------------------------------------------------------------------------
// traj_kernel: one CTA of 4 warps per pset.  NPW warps integrate — one (K ≤ 32: lane l owns type l; K > 64: lane l
// owns types l, l+32, …, KPL per lane, δΓ in shared memory) or two (32 < K ≤ 64: thread t of the pair owns type t and
// keeps its row of δΓ in registers; the pair exchanges ĉ_n through shared memory and a named barrier per Taylor
// stage) — the other warps build the table and the Chebyshev node values behind them; then the whole CTA finishes the
// Chebyshev stage.
//
// Taylor recurrence for the h-scaled coefficients ĉ_n of p(τ_s + θh) = Σ ĉ_n θ^n:
//   ĉ_{n+1} = h/(n+1) · [ −(λ+μ) ĉ_n + [n=0] μ(1−σ) + λ Σ_{m≤n} ĉ_m ĉ_{n−m} + δ Γ ĉ_n ]
// which is dp_dt! (mbp.jl:266-278) expanded order by order.  Cells are NON-uniform: after every cell the
// width is re-sized from the error indicator (|ĉ_M| + |ĉ_{M−1}| ~ h^M), a cell above tol is redone smaller.
// Table row for (cell s, type x), L = M+2, θ = (τ − τ_s)/h_s:
//   row[0] = F(τ_s),  row[1] = h(2λĉ_0 − Λ),  row[n+1] = 2hλ ĉ_n/(n+1).
// KPL == 1 (K ≤ KP ≤ 32): the lane's row of δΓ lives in registers, ĉ_n is broadcast through shared memory.
// KPL  > 1: δΓ is staged transposed in shared memory (lane-contiguous rows).
// ------------------------------------------------------------------------------------------------
template <int M, int KP, int KPL, int NPW>
__global__ void __launch_bounds__(TRAJ_THREADS) traj_kernel(ParamsView pv, PsetPack pk, TrajCfg cfg, ChebCfg cc,
                                                            double* __restrict__ tab) {
    extern __shared__ __align__(16) double smem[];
    constexpr int L = M + 2;
    constexpr int NPROD = 32 * NPW, NCONS = TRAJ_THREADS - NPROD;   // integrating / consumer threads
    static_assert(NPW == 1 || (NPW == 2 && KPL == 1), "two integrating warps own one type per lane");
    const int p = blockIdx.x;
    const int K = pv.K;
#ifdef GCDYN_TRAJ_TRACE
    __shared__ long long g_trace[16];
    __shared__ int g_attempts;
#endif
    TRACE_STAMP(0);
#ifdef GCDYN_TRAJ_TRACE
    if (threadIdx.x == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_trace[14] = gt; }
#endif
    __shared__ __align__(8) uint64_t s_full[CHEB_RING_SLOTS], s_empty[CHEB_RING_SLOTS];   // ring of cells → Chebyshev warps
    const bool cheb = cc.D > 0;                         // CTA-uniform
    const size_t traj_doubles = 2 * (size_t)(KPL == 1 ? KP : ((K + 3) & ~3)) + (KPL == 1 ? 0 : (((size_t)K * K + 1) & ~(size_t)1));
    const ChebSmem cs(smem + traj_doubles, K, cheb ? cc.D : 0, cfg.S_cap, L);
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
#pragma unroll
        for (int i = 0; i < CHEB_RING_SLOTS; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, NCONS); }
    }
    __syncthreads();
    pdl_launch_dependents();                            // the contraction's CTAs may become resident behind this grid
    // roles: prod_rank = 0..NPW−1 for the integrating warps (−1 otherwise), cons_rank = 0.. for the others
    int pw = 0, prod_rank = -1, cons_rank = -1;
    {
        int lo = s_wid[0];
#pragma unroll
        for (int w = 1; w < TRAJ_THREADS / 32; ++w) lo = min(lo, s_wid[w]);
        bool isp[TRAJ_THREADS / 32];
        int np = 0;
        if (NPW == 1) {
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
        prep_pset(pv, pk, p, 0, ct, NCONS, false);
        cell_consume<M>(pv, cfg, cc, cheb, p, tab, cs, s_full, s_empty, ct, NCONS);
    }
    TRACE_STAMP(1);
    if (prod_rank >= 0) {
    const int pt = prod_rank * 32 + lane;               // index among the integrating threads
    // exchange between the integrating warps (NPW == 2): named barrier 2 and a double-buffered pair of slots
    __shared__ unsigned long long s_xch[2][2];
    int xtog = 0;
    auto prod_sync = [&]() {
        if (NPW == 1) __syncwarp();
        else asm volatile("bar.sync 2, %0;" ::"r"(NPROD) : "memory");
    };
    auto prod_max_u64 = [&](unsigned long long v) -> unsigned long long {      // v is already warp-uniform
        if (NPW == 1) return v;
        if (lane == 0) s_xch[xtog][prod_rank] = v;
        asm volatile("bar.sync 2, %0;" ::"r"(NPROD) : "memory");
        const unsigned long long a0 = s_xch[xtog][0], a1 = s_xch[xtog][1];
        xtog ^= 1;
        return a0 > a1 ? a0 : a1;
    };
    const int KS = KPL == 1 ? KP : ((K + 3) & ~3);     // broadcast buffer length (multiple of 4)
    double* cb = smem;                                  // 2 × KS
    double* GT = smem + 2 * KS;                         // KPL > 1 only: GT[j*K + i] = δ Γ[i,j]

    const double mu = pv.mu[p], delta = pv.delta[p], rho = pv.rho[p], sigma = pv.sigma[p];
    const double* G = pv.Gamma + (size_t)p * pv.gamma_stride;
    const double b = mu * (1.0 - sigma);

    double lam[KPL], a[KPL], Lam[KPL];
    bool own[KPL];
    double rate = 0.0;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int i = pt + NPROD * q;
        own[q] = i < K;
        lam[q] = own[q] ? birth_rate(pv, p, i) : 0.0;
        const double gam = own[q] ? delta * -G[i + (size_t)i * K] : 0.0;
        Lam[q] = lam[q] + mu + gam;
        a[q] = lam[q] + mu;
        const double r = own[q] ? fabs(lam[q]) + fabs(mu) + 2.0 * fabs(gam) : 0.0;
        if (!(r <= rate)) rate = r;
    }
    rate = warp_max_nan(rate);
    if (NPW > 1) {                                      // non-negative doubles (and NaN) order like their bit patterns
        rate = __longlong_as_double((long long)prod_max_u64((unsigned long long)__double_as_longlong(fabs(rate))));
    }

    double g[KPL == 1 ? KP : 1];
    if constexpr (KPL == 1) {
#pragma unroll
        for (int j = 0; j < KP; ++j) g[j] = (own[0] && j < K) ? delta * G[pt + (size_t)j * K] : 0.0;
    } else {
        for (int e = lane; e < K * K; e += 32) GT[e] = delta * G[e];   // G[i + j*K] → GT[j*K + i]
    }
    for (int e = pt; e < 2 * KS; e += NPROD) cb[e] = 0.0;
    prod_sync();

    const bool adaptive = cfg.S_fixed <= 0;
    // (the table block of this pset — [τ_s | 1/h_s | pad][rows (cell, type) of L doubles] — is written by the consumers)
    double pcur[KPL];
#pragma unroll
    for (int q = 0; q < KPL; ++q) pcur[q] = own[q] ? 1.0 - rho : 0.0;   // p(0) = 1 − ρ, mbp.jl:130,141,226
    // first cell width: optimistic guess from the fastest rate; the controller below corrects it
    double h = adaptive ? hrate_target_start(M) / rate : cfg.T / (double)cfg.S_fixed;
    if (!(h > 0.0) || !(h <= 0.25 * cfg.T)) h = adaptive ? 0.25 * cfg.T : cfg.T / (double)cfg.S_fixed;
    double tau = 0.0;
    // Step control works on the HIGH WORD of the indicator (sign, exponent, 20 mantissa bits): non-negative doubles
    // order like their bit patterns and NaN sorts above +Inf, so one integer REDUX gives the warp maximum, the accept
    // test is an integer compare (at most 1e-6 stricter than ind ≤ tol) and log2(ind) ≈ (hi − 0x3ff00000)/2^20
    // (exponent plus linear mantissa, error < 0.09) is all the accuracy (tol/ind)^(1/M) needs.  Every lane sees the
    // same value, so accept/reject decisions are warp-uniform and deterministic.
    const unsigned tol_hi = (unsigned)__double2hiint(cfg.tol);
    const float tol_l2 = (float)((int)tol_hi - 0x3ff00000) * (1.0f / 1048576.0f);
    const float decade = exp2f(__log2f(0.1f) * (1.0f / (float)M));      // 0.1^(1/M)
    unsigned errmax_hi = 0u;
    bool bad = false, failed = false;
    int s = 0, rejects = 0;
    TRACE_STAMP(2);
#ifdef GCDYN_TRAJ_TRACE
    long long tr_a = 0, tr_b = 0, tr_c = 0, tr_t = clock64();
#endif
    for (;;) {
        double hc = h;
        bool last = false;
        if (adaptive) { const double rem = cfg.T - tau; if (hc >= rem * (1.0 - 1e-3)) { hc = rem; last = true; } }
        else if (s + 1 >= cfg.S_fixed) { hc = cfg.T - tau; last = true; }
        double c[M + 1][KPL];
#pragma unroll
        for (int q = 0; q < KPL; ++q) c[0][q] = pcur[q];
#pragma unroll
        for (int n = 0; n < M; ++n) {
            double* buf = cb + (n & 1) * KS;
#pragma unroll
            for (int q = 0; q < KPL; ++q)
                if (own[q]) buf[pt + NPROD * q] = c[n][q];
            prod_sync();
            double dot[KPL];
            if constexpr (KPL == 1) {
                double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                for (int j = 0; j < KP; j += 4) {
                    const double2 v01 = *reinterpret_cast<const double2*>(buf + j);
                    const double2 v23 = *reinterpret_cast<const double2*>(buf + j + 2);
                    d0 = fma(g[j], v01.x, d0);
                    d1 = fma(g[j + 1], v01.y, d1);
                    d2 = fma(g[j + 2], v23.x, d2);
                    d3 = fma(g[j + 3], v23.y, d3);
                }
                dot[0] = (d0 + d1) + (d2 + d3);
            } else {
                double d[KPL][2];
#pragma unroll
                for (int q = 0; q < KPL; ++q) d[q][0] = d[q][1] = 0.0;
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
                for (int q = 0; q < KPL; ++q) dot[q] = d[q][0] + d[q][1];
            }
            const double hn = hc * (1.0 / (double)(n + 1));
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                // Σ_{m≤n} ĉ_m ĉ_{n−m} = 2 Σ_{m<n/2} ĉ_m ĉ_{n−m} (+ ĉ_{n/2}² when n is even)
                double cv0 = 0.0, cv1 = 0.0;
#pragma unroll
                for (int m = 0; 2 * m < n; ++m) {
                    if (m & 1) cv1 = fma(c[m][q], c[n - m][q], cv1);
                    else cv0 = fma(c[m][q], c[n - m][q], cv0);
                }
                double conv = 2.0 * (cv0 + cv1);
                if ((n & 1) == 0) conv = fma(c[n / 2][q], c[n / 2][q], conv);
                double rhs = fma(lam[q], conv, dot[q]);
                rhs = fma(-a[q], c[n][q], rhs);
                if (n == 0) rhs += b;
                c[n + 1][q] = rhs * hn;
            }
        }
#ifdef GCDYN_TRAJ_TRACE
        if (s == 0 && rejects == 0) TRACE_STAMP(11);
        if (s > 0) { const long long t = clock64(); tr_a += t - tr_t; tr_t = t; }
#endif
        // error indicator of this cell: the last two Taylor terms, maximum over the pset's types
        double ind = 0.0;
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const double v = fabs(c[M][q]) + fabs(c[M - 1][q]);
            if (own[q] && !(v <= ind)) ind = v;
        }
        const unsigned ind_hi = (unsigned)prod_max_u64(__reduce_max_sync(FULL, (unsigned)__double2hiint(ind)));
        const float ind_l2 = (float)((int)ind_hi - 0x3ff00000) * (1.0f / 1048576.0f);
        const float ratio = exp2f((tol_l2 - ind_l2) * (1.0f / (float)M));     // ≈ (tol/ind)^(1/M)
        if (adaptive && !(ind_hi < tol_hi)) {
            // reject: the indicator scales like h^M — shrink the cell accordingly and redo it
            float f = 0.9f * ratio;
            if (!(f >= 0.2f)) f = 0.2f;
            if (f > 0.9f) f = 0.9f;
            h = hc * (double)f;
            if (++rejects > 200 || !(h > cfg.T * 1e-10)) { failed = true; break; }
            continue;
        }
        if (s >= cfg.S_cap) { failed = true; break; }
#ifdef GCDYN_TRAJ_TRACE
        if (s == 0 && rejects == 0) TRACE_STAMP(12);
        if (s > 0) { const long long t = clock64(); tr_b += t - tr_t; tr_t = t; }
#endif
        // accept: publish ĉ_0..ĉ_M, τ_s and h_s to the consumer warps (they build and store the table rows of F, which
        // the integration itself never needs) and advance p to the next grid point
        {
            const int slot_i = s % CHEB_RING_SLOTS;
            if (s >= CHEB_RING_SLOTS) mbar_wait(s_empty + slot_i, (uint32_t)(s / CHEB_RING_SLOTS - 1) & 1u);   // consumers left it
            double* slot = cs.ring + (size_t)slot_i * K * L;
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                if (own[q]) {
                    double2* dst = reinterpret_cast<double2*>(slot + (size_t)(pt + NPROD * q) * L);
#pragma unroll
                    for (int n = 0; n < M; n += 2) dst[n >> 1] = make_double2(c[n][q], c[n + 1][q]);
                    dst[M >> 1] = make_double2(c[M][q], 0.0);
                }
            }
            if (pt == 0) {
                cs.sgrid[s] = tau; cs.sgrid[s + 1] = last ? cfg.T : tau + hc; cs.shc[s] = hc;
                cs.sflag[s] = last ? CELL_LAST : 0;
            }
            prod_sync();
            if (pt == 0) mbar_arrive(s_full + slot_i);
        }
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            double ps[4] = {0.0, 0.0, 0.0, 0.0};             // Σ_{n≥1} ĉ_n, smallest terms first, four partial sums
#pragma unroll
            for (int n = M; n >= 1; --n) ps[n & 3] += c[n][q];
            pcur[q] = ((ps[0] + ps[1]) + (ps[2] + ps[3])) + c[0][q];
            if (own[q] && !(pcur[q] >= -1e-9 && pcur[q] <= 1.0 + 1e-9)) bad = true;   // isoutofdomain, mbp.jl:146,204,231
        }
        if (ind_hi > errmax_hi) errmax_hi = ind_hi;
#ifdef GCDYN_TRAJ_TRACE
        if (s == 0 && rejects == 0) TRACE_STAMP(13);
        { const long long t = clock64(); if (s > 0) tr_c += t - tr_t; tr_t = t; }
#endif
        ++s;
        if (last) break;
        tau += hc;
        if (adaptive) {
            // next cell: aim a decade below tol; never more than double, never less than half
            float f = 0.9f * decade * ratio;
            if (!(f <= 2.0f)) f = 2.0f;            // also ind == 0 → +Inf, NaN
            if (f < 0.5f) f = 0.5f;
            h = hc * (double)f;
        }
    }
    if (failed) {                                       // tell the consumers to stop (the regular exit flagged its last cell)
        const int slot_i = s % CHEB_RING_SLOTS;
        if (s >= CHEB_RING_SLOTS) mbar_wait(s_empty + slot_i, (uint32_t)(s / CHEB_RING_SLOTS - 1) & 1u);
        if (pt == 0) { cs.sflag[s] = CELL_FAIL; mbar_arrive(s_full + slot_i); }
    }
    TRACE_STAMP(3);
#ifdef GCDYN_TRAJ_TRACE
    if (lane == 0) { g_attempts = s + rejects; g_trace[1] = tr_a; g_trace[5] = tr_b; g_trace[11] = tr_c; }
#endif
    // conditioning term and diagnostics
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int i = pt + NPROD * q;
        if (own[q]) pk.logcond[(size_t)p * K + i] = log1p(-pcur[q]);   // log(1 − p_i(T)), mbp.jl:244-245
    }
    const bool unresolved = failed || !(errmax_hi < tol_hi);
    const bool anybad = prod_max_u64(__any_sync(FULL, bad) ? 1ULL : 0ULL) != 0ULL || unresolved;
    if (pt == 0) {
        pk.perr[p] = __hiloint2double((int)min(errmax_hi + 1u, 0x7ff00000u), 0);   // upper bound of the largest accepted indicator
        pk.S[p] = s > 0 ? s : 1;
        pk.pstatus[p] = anybad ? ST_SOLVER : 0;
    }
    }   // integrating warps
    if (!cheb) return;
    __syncthreads();                                    // node values, status of this pset are complete and visible
    TRACE_STAMP(4);
    cheb_finish<M>(pv, pk, cc, p, cs TRACE_PASS);
#ifdef GCDYN_TRAJ_TRACE
    __syncthreads();
    if (threadIdx.x == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_trace[15] = gt; }
    if (threadIdx.x == 0 && p < 1024) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        long long* o = g_trace_out + (size_t)p * 24;
        for (int i = 0; i < 16; ++i) o[i] = g_trace[i];
        o[16] = smid; o[17] = pw; o[18] = pk.S[p]; o[19] = g_attempts; o[20] = s_wid[0]; o[21] = s_wid[1]; o[22] = s_wid[2]; o[23] = s_wid[3];
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
