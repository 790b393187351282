// Chebyshev-moment formulation of the per-tree sums.
//
// F_x(τ) is smooth on [0, T], so for every parameter set it is represented by ONE degree-D polynomial per type — the
// interpolant through the D+1 Chebyshev–Gauss–Lobatto nodes τ_k = T(1 + cos(πk/D))/2,
//     F_x(τ) ≈ Σ_k F_x(τ_k) · ℓ_k(u),  u = 2τ/T − 1,   ℓ_k(u) = (2/D) w_k Σ''_d cos(πkd/D) T_d(u),
// and the per-tree sum over lookup items becomes a dense contraction with pset-independent per-tree statistics in the
// NODAL basis, Mt[t][k][x] = Σ_{items of tree t, type x} w · ℓ_k(u_item) (built once per forest from the Chebyshev
// moments Σ w·T_d(u_item) by one cosine transform), against the node values the trajectory kernel produces anyway:
//     LL[p][t] = Σ_j A[p][j] · Mt[t][j]
// i.e. a [P × J] · [J × T] FP64 GEMM executed with DMMA (mma.sync.m8n8k4.f64).  Column space j:
//     [0, K)            births by branch type            ×  log λ_x                      (mbp.jl:175)
//     K, K+1            sampled deaths, sampled survivors ×  log μ + log σ,  log ρ        (mbp.jl:168,171)
//     [K+2, 2K+2)       root type one-hot                 ×  −log(1 − p_x(T))             (mbp.jl:244-245)
//     2K+2              number of type changes            ×  log δ   (only when Γ is shared by all psets)
//     2K+3              padding
//     [C1, C1+(D+1)K)   nodal moments, node-major           ×  F_x(τ_k)
//     [Jtc, Jtc+K²)     type changes x→y                  ×  log δΓ_xy  (only when every pset has its own Γ;
//                                                            with a shared Γ the Σ count·log Γ_xy part is a
//                                                            per-tree constant added in the GEMM epilogue)
// Psets whose series is not resolved, or whose node-term logs are not finite where the forest needs them,
// are re-done exactly (item sums from the Taylor table) inside finish_kernel.
#pragma once
#include "kernels.cuh"
#include "peer_allreduce.cuh"

namespace gcdyn {

// ------------------------------------------------------------------------------------------------
// counts_kernel: one CTA per tree (run once per forest and present_time) — the count columns, the K² type-change
// block and zero padding of the tree's row of Mt.  Integer atomics in shared memory: exact, order-independent.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) counts_kernel(ForestView fv, int Dcap, long long Jtc, long long Jst,
                                                     double* __restrict__ Mt) {
    extern __shared__ __align__(16) int cnt[];
    const int K = fv.K;
    const int C1 = 2 * K + 4;
    const int NC = C1 + K * K;
    const int64_t t = blockIdx.x;
    for (int i = threadIdx.x; i < NC; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    // (a tree with a self-loop is -Inf through its status; its records may hold placeholder indices: leave its row zero)
    const int64_t i_end = fv.self_loop[t] ? fv.tree_off[t] : fv.tree_off[t + 1];
    for (int64_t i = fv.tree_off[t] + threadIdx.x; i < i_end; i += blockDim.x) {
        if (!fv.live[i]) continue;
        const int ev = fv.event[i], x = fv.btype_idx[i], y = fv.type_idx[i];
        if (ev == EV_BIRTH) atomicAdd(&cnt[x], 1);
        else if (ev == EV_TC) { atomicAdd(&cnt[C1 + x * K + y], 1); atomicAdd(&cnt[2 * K + 2], 1); }
        else if (ev == EV_SDEATH) atomicAdd(&cnt[K], 1);
        else if (ev == EV_SSURV) atomicAdd(&cnt[K + 1], 1);
    }
    if (threadIdx.x == 0) {
        const int rt = fv.root_type[t];
        if (rt >= 0 && !fv.self_loop[t]) cnt[K + 2 + rt] = 1;
    }
    __syncthreads();
    double* row = Mt + (size_t)t * Jst;
    const long long cheb_end = C1 + (long long)(Dcap + 1) * K;
    for (long long i = threadIdx.x; i < Jst; i += blockDim.x) {
        if (i < C1) row[i] = (double)cnt[i];
        else if (i < cheb_end) continue;                                   // written by moments_kernel
        else if (i >= Jtc && i < Jtc + (long long)K * K) row[i] = (double)cnt[C1 + (i - Jtc)];
        else row[i] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// moments_kernel: one WARP per (tree, type): Mt[t][C1 + d*K + x] = Σ_{items of t with type x} w·T_d(u), d ≤ Dcap.
// The warp gathers the tree's items of its type 32 at a time (ballot compaction keeps tree order), every lane
// runs the three-term recurrence for its item, and each degree is summed with a fixed-order shuffle tree —
// deterministic, and parallel over trees × types × items.  Lane d%32 keeps the accumulator of degree d.
// ------------------------------------------------------------------------------------------------
constexpr int MOM_DBLK = 5;   // ceil((Dcap+1)/32) for Dcap = 128 … 159

__global__ void __launch_bounds__(256) moments_kernel(ForestView fv, double T, int Dcap, long long Jst, int C1,
                                                      double* __restrict__ Mt) {
    const int K = fv.K;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= fv.n_trees * K) return;
    const int64_t t = wid / K;
    const int x = (int)(wid - t * K);
    double acc[MOM_DBLK];
#pragma unroll
    for (int b = 0; b < MOM_DBLK; ++b) acc[b] = 0.0;
    const int64_t i0 = fv.item_off[t], i1 = fv.item_off[t + 1];
    double u_mine = 0.0, w_mine = 0.0;    // the item this lane holds in the current batch (w = 0: empty slot)
    int filled = 0;
    auto flush = [&]() {
        double tm = 1.0, tc = u_mine;
#pragma unroll
        for (int b = 0; b < MOM_DBLK; ++b) {
            for (int dd = 0; dd < 32; ++dd) {
                const int d = b * 32 + dd;
                if (d > Dcap) break;
                const double td = d == 0 ? 1.0 : tc;
                const double ssum = warp_sum(w_mine * td);
                if (lane == dd) acc[b] += ssum;
                if (d >= 1) { const double tn = fma(2.0 * u_mine, tc, -tm); tm = tc; tc = tn; }
            }
        }
        w_mine = 0.0; u_mine = 0.0; filled = 0;
    };
    for (int64_t base = i0; base < i1; base += 32) {
        const int64_t i = base + lane;
        int2 me = make_int2(0, 0);
        double tm_i = 0.0;
        bool match = false;
        if (i < i1) { me = fv.it_meta[i]; tm_i = fv.it_time[i]; match = (me.x & 0xffffff) == x; }
        unsigned mask = __ballot_sync(FULL, match);
        while (mask) {
            // move as many matches as fit into the free slots [filled, 32), in tree order
            const int room = 32 - filled;
            const int rank = __popc(mask & ((1u << lane) - 1u));          // my position among the remaining matches
            const bool take = match && ((mask >> lane) & 1u) && rank < room;
            const int n_take = min(__popc(mask), room);
            // destination slot for taker `rank` is filled + rank; gather with shuffles: slot s reads the lane whose rank == s − filled
            const unsigned takers = __ballot_sync(FULL, take);
            const int want = lane - filled;                                 // which taker this slot receives
            int src = -1;
            if (want >= 0 && want < n_take) {
                // lane index of the (want)-th set bit of `takers`
                unsigned m = takers;
                for (int r = 0; r < want; ++r) m &= m - 1;
                src = __ffs(m) - 1;
            }
            const double u_src = 2.0 * (T - tm_i) / T - 1.0;
            const double w_src = (double)(me.x >> 24);
            const double u_got = __shfl_sync(FULL, u_src, src < 0 ? 0 : src);
            const double w_got = __shfl_sync(FULL, w_src, src < 0 ? 0 : src);
            if (src >= 0) { u_mine = u_got; w_mine = w_got; }
            filled += n_take;
            mask &= ~takers;
            if (filled == 32) flush();
        }
    }
    if (filled > 0) flush();
    double* row = Mt + (size_t)t * Jst + C1;
#pragma unroll
    for (int b = 0; b < MOM_DBLK; ++b) {
        const int d = b * 32 + lane;
        if (d <= Dcap) row[(size_t)d * K + x] = acc[b];
    }
}

// nodal_transform_kernel: one CTA per tree, in place: Chebyshev moments M[d][x] → nodal moments
//     N[k][x] = (2/D) w_k Σ''_{d ≤ D} cos(πkd/D) M[d][x]        (Σ'': d = 0 and d = D halved;  w_0 = w_D = 1/2)
// so that Σ_k f(τ_k) N[k][x] = Σ''_d a_d M[d][x] with a_d the Chebyshev coefficients of the interpolant through f(τ_k).
__global__ void __launch_bounds__(256) nodal_transform_kernel(int K, int D, int C1, long long Jst, double* __restrict__ Mt) {
    extern __shared__ __align__(16) double nsm[];          // [2D] cos table | [(D+1)*K] moments of this tree
    double* ctab = nsm;
    double* Mo = nsm + 2 * D;
    double* row = Mt + (size_t)blockIdx.x * Jst + C1;
    const int NV = (D + 1) * K;
    for (int m = threadIdx.x; m < 2 * D; m += blockDim.x) ctab[m] = cospi((double)m / (double)D);
    for (int i = threadIdx.x; i < NV; i += blockDim.x) Mo[i] = row[i];
    __syncthreads();
    const double scale = 2.0 / (double)D;
    for (int i = threadIdx.x; i < NV; i += blockDim.x) {
        const int k = i / K, x = i - k * K;
        double s0 = 0.0, s1 = 0.0;
        int idx = 0;                                        // (k·d) mod 2D
        for (int d = 0; d <= D; ++d) {
            const double wd = (d == 0 || d == D) ? 0.5 : 1.0;
            const double t = wd * Mo[d * K + x] * ctab[idx];
            if (d & 1) s1 += t; else s0 += t;
            idx += k;
            if (idx >= 2 * D) idx -= 2 * D;
        }
        const double wk = (k == 0 || k == D) ? 0.5 : 1.0;
        row[i] = scale * wk * (s0 + s1);
    }
}

// column sums: colsum[c] = Σ_t Mt[t][c].  A CTA owns 32 columns; its 8 warps take the trees t ≡ w (mod 8) in order,
// then the 8 partial sums are added in warp order: fixed order, coalesced rows.
__global__ void __launch_bounds__(256) colsum_kernel(const double* __restrict__ Mt, int64_t T, long long Jst,
                                                     double* __restrict__ colsum) {
    __shared__ double part[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long c = (long long)blockIdx.x * 32 + lane;
    double s = 0.0;
    if (c < Jst)
        for (int64_t t = w; t < T; t += 8) s += Mt[(size_t)t * Jst + c];
    part[w][lane] = s;
    __syncthreads();
    if (w == 0 && c < Jst) {
        double tot = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += part[i][lane];
        colsum[c] = tot;
    }
}

// degree_probe_kernel (calibration of a forest's grid, once): one CTA per parameter set computes the full Chebyshev
// spectrum of its interpolants from the node values and the smallest degree whose tail Σ_{k>d}|a_k| is below tol for every
// type; the maximum over the batch lands in out[0] (atomicMax; the host zeroes it first), the grid degree it was measured
// on in out[1].  Parameter sets that are flagged already (or failed) do not vote.
__global__ void __launch_bounds__(256) degree_probe_kernel(const double* __restrict__ A, MomentView mv, const int* __restrict__ need_sweep,
                                                           const int* __restrict__ pstatus, double tol, int* __restrict__ out) {
    extern __shared__ __align__(16) double psm[];          // [2D] cos | [(D+1)*K] values | [(D+1)*K] |a_d|
    const int K = mv.K, D = mv.D, NV = (D + 1) * K;
    const int p = blockIdx.x;
    if (need_sweep[p] || (pstatus[p] & ST_SOLVER)) return;   // CTA-uniform
    double* ctab = psm;
    double* V = psm + 2 * D;
    double* Ac = V + NV;
    __shared__ int s_deff;
    if (threadIdx.x == 0) s_deff = 4;
    for (int m = threadIdx.x; m < 2 * D; m += blockDim.x) ctab[m] = cospi((double)m / (double)D);
    const double* row = A + (size_t)p * mv.Jst + mv.C1;
    for (int i = threadIdx.x; i < NV; i += blockDim.x) V[i] = row[i];
    __syncthreads();
    for (int i = threadIdx.x; i < NV; i += blockDim.x) {
        const int d = i / K, x = i - d * K;
        double s0 = 0.0, s1 = 0.0;
        int idx = 0;
        for (int k = 0; k <= D; ++k) {
            const double wk = (k == 0 || k == D) ? 0.5 : 1.0;
            const double t = wk * V[k * K + x] * ctab[idx];
            if (k & 1) s1 += t; else s0 += t;
            idx += d;
            if (idx >= 2 * D) idx -= 2 * D;
        }
        Ac[i] = fabs((2.0 / (double)D) * (s0 + s1)) * ((d == 0 || d == D) ? 0.5 : 1.0);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < K; x += blockDim.x) {
        double tail = 0.0;
        int found = 4;
        for (int d = D; d > 4; --d) {
            tail += Ac[d * K + x];
            if (!(tail <= tol)) { found = d; break; }
        }
        atomicMax(&s_deff, found);
    }
    __syncthreads();
    if (threadIdx.x == 0) { atomicMax(out, s_deff); out[1] = D; }
}

// Shared Γ: per-tree constant Σ_{x→y} count · log Γ_xy (the pset-independent part of Σ log δΓ_xy), one warp per tree.
__global__ void tc_const_kernel(MomentView mv, int64_t T, const double* __restrict__ Gamma /*col-major*/,
                                double* __restrict__ tree_const) {
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    const int K = mv.K;
    const double* cnt = mv.Mt + (size_t)t * mv.Jst + mv.Jtc;
    double acc = 0.0;
    for (int e = lane; e < K * K; e += 32) {
        const double n = cnt[e];
        if (n != 0.0) {
            const int x = e / K, y = e - x * K;
            acc += n * log(Gamma[x + (size_t)y * K]);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) tree_const[t] = acc;
}

// ------------------------------------------------------------------------------------------------
// gemm_kernel: out[p][t] = Σ_j A[p][j]·Mt[t][j] (+ tree_const[t]) with DMMA over up to two column segments.
// CTA tile 32(p) × 32(t), 4 warps of 16×16, k-chunks of 32 staged through shared memory with cp.async
// (2 stages), row stride 36 doubles (conflict-free for the m8n8k4 fragment pattern: row = lane/4, k = lane%4).
// ------------------------------------------------------------------------------------------------
constexpr int GT_TILE = 32, GT_KC = 32, GT_LD = 36, GT_STAGES = 4;
constexpr int GT_SMEM_BYTES = GT_STAGES * 2 * GT_TILE * GT_LD * 8;

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// blockIdx.z = split of the column range (split-K): each split writes its own partial[z][p][t]; finish_kernel adds them.
__global__ void __launch_bounds__(128) gemm_kernel(const double* __restrict__ A, MomentView mv, int use_tc_block,
                                                   int P, int64_t T, double* __restrict__ partial) {
    extern __shared__ __align__(16) double gsm[];          // GT_STAGES × (A tile | B tile)
    auto As = [&](int st) { return gsm + (size_t)st * 2 * GT_TILE * GT_LD; };
    auto Bs = [&](int st) { return gsm + (size_t)st * 2 * GT_TILE * GT_LD + GT_TILE * GT_LD; };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned n_pt = (unsigned)((P + GT_TILE - 1) / GT_TILE);             // pset tiles vary fastest (see the launch site)
    const int p0 = (int)(blockIdx.x % n_pt) * GT_TILE;
    const int64_t t0 = (int64_t)(blockIdx.x / n_pt) * GT_TILE;
    const long long Jst = mv.Jst;
    const double* Mt = mv.Mt;
    // segment 1: count columns + nodal columns; segment 2: K² block
    const long long J1 = mv.J1;
    const int n1 = (int)((J1 + GT_KC - 1) / GT_KC);
    const long long J2 = use_tc_block ? ((long long)mv.K * mv.K + 3) & ~3LL : 0;
    const int n2 = (int)((J2 + GT_KC - 1) / GT_KC);
    const int nall = n1 + n2;
    const int c_lo = (int)((long long)nall * blockIdx.z / gridDim.z), c_hi = (int)((long long)nall * (blockIdx.z + 1) / gridDim.z);
    const int nchunks = c_hi - c_lo;
    // this thread's cp.async assignments: 32 rows × 16 segments of 16 B per operand = 512 per tile, 4 per thread;
    // which: 1 = the A tile (psets), 2 = the B tile (trees), 3 = both
    auto issue = [&](int rel, int buf, int which) {
        const int chunk = c_lo + rel;
        long long j0, jend;
        if (chunk < n1) { j0 = (long long)chunk * GT_KC; jend = J1; }
        else { j0 = mv.Jtc + (long long)(chunk - n1) * GT_KC; jend = mv.Jtc + J2; if (jend > Jst) jend = Jst; }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int seg = tid + r * 128;
            const int row = seg >> 4, col = (seg & 15) * 2;          // col in doubles
            const long long j = j0 + col;
            const bool inj = j < jend;
            if (which & 1) {
                const bool va = inj && (p0 + row) < P;
                cp_async16(As(buf) + row * GT_LD + col, A + (size_t)(va ? p0 + row : 0) * Jst + (va ? j : 0), va);
            }
            if (which & 2) {
                const bool vb = inj && (t0 + row) < T;
                cp_async16(Bs(buf) + row * GT_LD + col, Mt + (size_t)(vb ? t0 + row : 0) * Jst + (vb ? j : 0), vb);
            }
        }
    };
    // the forest's moments do not depend on the trajectory kernel: request their first stages while it is still running
#pragma unroll
    for (int st = 0; st < GT_STAGES - 1; ++st)
        if (st < nchunks) issue(st, st, 2);
    pdl_wait();                                            // A comes from the trajectory kernel
    pdl_launch_dependents();
    double acc[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    const int wr = (warp >> 1) * 16, wc = (warp & 1) * 16;     // warp's 16×16 sub-tile (p rows, t cols)
    const int fr = lane >> 2, fk = lane & 3;
    // multi-stage cp.async pipeline: GT_STAGES−1 chunks in flight, one barrier per chunk
#pragma unroll
    for (int st = 0; st < GT_STAGES - 1; ++st) {
        if (st < nchunks) issue(st, st, 1);
        cp_async_commit();
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<GT_STAGES - 2>();
        __syncthreads();                      // chunk ch has landed for everyone; buffer (ch−1) % STAGES is free
        const int nxt = ch + GT_STAGES - 1;
        if (nxt < nchunks) issue(nxt, nxt % GT_STAGES, 3);
        cp_async_commit();
        const double* as = As(ch % GT_STAGES);
        const double* bs = Bs(ch % GT_STAGES);
#pragma unroll
        for (int kk = 0; kk < GT_KC / 4; ++kk) {
            const double a0 = as[(wr + fr) * GT_LD + kk * 4 + fk];
            const double a1 = as[(wr + 8 + fr) * GT_LD + kk * 4 + fk];
            const double b0 = bs[(wc + fr) * GT_LD + kk * 4 + fk];
            const double b1 = bs[(wc + 8 + fr) * GT_LD + kk * 4 + fk];
            dmma_m8n8k4(acc[0][0][0], acc[0][0][1], a0, b0);
            dmma_m8n8k4(acc[0][1][0], acc[0][1][1], a0, b1);
            dmma_m8n8k4(acc[1][0][0], acc[1][0][1], a1, b0);
            dmma_m8n8k4(acc[1][1][0], acc[1][1][1], a1, b1);
        }
    }
    // epilogue: C fragment (row = lane/4, cols = 2*(lane%4) + {0,1}) → this split's partial sums
    double* dst = partial + (size_t)blockIdx.z * (size_t)P * (size_t)T;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int p = p0 + wr + 8 * a + fr;
        if (p >= P) continue;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t t = t0 + wc + 8 * b + 2 * fk + e;
                if (t < T) dst[(size_t)p * T + t] = acc[a][b][e];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// finish_kernel: one CTA per pset closes the evaluation.  Unflagged psets: add the GEMM's split partials and the
// per-tree constant, apply the -Inf / status rules.  Flagged psets (series unresolved, non-finite node-term logs):
// exact item sums for every tree straight from the pset's Taylor table (the fallback).  Then the fixed-order sum
// over trees → out_pset[p] (mbp.jl:258), in the same order reduce_kernel uses.
// ------------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(256) finish_kernel(ForestView fv, ParamsView pv, PsetPack pk, const double* __restrict__ tab,
                                                     long long tab_stride, int S_cap, int grid_doubles, double Tpres,
                                                     const double* __restrict__ partial, int splits,
                                                     const double* __restrict__ tree_const,
                                                     const int* __restrict__ need_sweep, int P, int D, int* __restrict__ nflag, int* __restrict__ feedback, int totals_only,
                                                     double* __restrict__ out_tree, int32_t* __restrict__ status,
                                                     double* __restrict__ out_pset, FusedAllReduce fc, int mh_on, MhView mh) {
    __shared__ double part[8];
    __shared__ double s_total;
    const int p = blockIdx.x;
    const int K = fv.K;
    const int64_t T = fv.n_trees;
    pdl_wait();                                            // partial sums come from the contraction
    const int pst = pk.pstatus[p];
    const bool redo = need_sweep[p] && !(pst & ST_SOLVER);
    const bool sharded = fc.pv.world > 1;
    // fused Metropolis iteration: accept/reject of the chains this kernel closes, and the iteration counter — bumped by
    // whichever CTA finishes last (every CTA has read the counter before it counts itself)
    const unsigned long long mh_it = mh_on ? *mh.it_dev + 1 : 0;
    auto mh_tick = [&]() {
        if (mh_on && threadIdx.x == 0 && atomicAdd(mh.tick_count, 1) == P - 1) { *mh.tick_count = 0; *mh.it_dev = mh_it; }
    };
    if (totals_only && !redo) {                            // traj_kernel already formed this pset's sum (and decided its chain)
        if (!sharded) { mh_tick(); return; }
        if (threadIdx.x == 0) s_total = fc.totals_in[p];
    } else {
    // psets whose interpolant was not resolved on this grid are summed exactly below.  When they are a quarter of the
    // batch the grid no longer fits the parameters: tell the host (pinned word), the next call widens it.
    if (threadIdx.x == 0 && feedback && (need_sweep[p] & 1) && !(pst & ST_SOLVER)) {
        const int before = atomicAdd(nflag, 1);
        if (4 * before < P && 4 * (before + 1) >= P) *feedback = D;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    double* orow = out_tree + (size_t)p * T;
    double acc = 0.0;
    bool have_acc = false;
    if (need_sweep[p] && !(pst & ST_SOLVER)) {
        if (pv.gamma_stride == 0) {                          // shared Γ: traj_kernel left the K² type-change logs to us
            prep_tc_logs(pv, pk, p, threadIdx.x, blockDim.x);
            __syncthreads();
        }
        const double* tp = tab + (size_t)p * tab_stride;
        const double* nt = pk.nt + (size_t)p * pk.ntw;
        const double lmusig = nt[K + K * K], lrho = nt[K + K * K + 1];
        const int S = pk.S[p];
        for (int64_t t = warp; t < T; t += nwarps) {
            const double isum = tree_items_sum<M, false>(fv, tp, nt, K, S, S_cap, grid_doubles, Tpres, fv.item_off[t],
                                                         fv.item_off[t + 1], lane);
            if (lane == 0) {
                const int rt = fv.root_type[t];
                double r = isum + count_times_log(fv.n_death[t], lmusig) + count_times_log(fv.n_surv[t], lrho)
                           - (rt >= 0 ? pk.logcond[(size_t)p * K + rt] : CUDART_NAN);
                int st = pst;
                if (fv.self_loop[t]) st |= ST_SELF_LOOP;
                if (st & ST_SELF_LOOP) r = -CUDART_INF;
                else if (r != r) st |= ST_DOMAIN;
                orow[t] = r;
                if (status) status[(size_t)p * T + t] = st;
            }
        }
    } else {
        // four trees per thread in flight (the partials come from L2); the per-pset sum accumulates the values just
        // computed, in the same thread/tree assignment — and therefore the same order — as the re-read below
        for (int64_t t0 = threadIdx.x; t0 < T; t0 += 4 * (int64_t)blockDim.x) {
            double r[4];
            int st[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t t = t0 + (int64_t)u * blockDim.x;
                r[u] = 0.0;
                st[u] = pst;
                if (t < T) {
                    for (int z = 0; z < splits; ++z) r[u] += __ldcg(partial + ((size_t)z * P + p) * (size_t)T + t);
                    if (tree_const) r[u] += __ldg(tree_const + t);
                    if (fv.self_loop[t]) st[u] |= ST_SELF_LOOP;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t t = t0 + (int64_t)u * blockDim.x;
                if (t < T) {
                    if (st[u] & (ST_SELF_LOOP | ST_SOLVER)) r[u] = -CUDART_INF;
                    else if (r[u] != r[u]) st[u] |= ST_DOMAIN;
                    orow[t] = r[u];
                    if (status) status[(size_t)p * T + t] = st[u];
                    acc += r[u];
                }
            }
        }
        have_acc = true;
    }
    if (!have_acc) {                                       // CTA-uniform: the fallback wrote orow warp by warp
        __syncthreads();
        for (int64_t t = threadIdx.x; t < T; t += blockDim.x) acc += orow[t];
    }
    acc = warp_sum(acc);
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += part[w];
        s_total = s;
    }
    }
    if (!sharded) {
        if (threadIdx.x == 0) {
            out_pset[p] = s_total;
            if (mh_on) mh_accept_chain(mh, mh_it, p);
        }
        mh_tick();
        return;
    }
    // ---- trees sharded over ranks: the all-reduce of the [P] sums is part of this kernel.  Every CTA stores its pset's
    //      sum as a self-validating entry straight into slot `rank` of EVERY rank's window (P2P stores over NVLink, no
    //      fence); the CTA that arrives last polls the entries of all ranks (bounded) and adds them in rank order, so every
    //      rank ends with bit-identical sums and no collective is launched.
    __shared__ int s_last, s_timeout;
    __shared__ unsigned long long s_seq;
    const PeerView& pr = fc.pv;
    if (threadIdx.x == 0) {
        const unsigned long long seq = *pr.seq_dev + 1;     // read before this CTA counts itself: the last CTA bumps it
        s_seq = seq;
        s_timeout = 0;
        const size_t slot = ((size_t)(seq & 1ULL) * pr.world + pr.rank) * pr.cap + p;
        for (int r = 0; r < pr.world; ++r) ll_store(pr.win[r].ll + slot, s_total, (unsigned)seq);
        s_last = atomicAdd(fc.done, 1) == P - 1;
    }
    __syncthreads();
    if (!s_last) return;
#ifdef GCDYN_COMM_TRACE
    unsigned long long tr0 = global_ns();
#endif
    const unsigned long long seq = s_seq;
    if (threadIdx.x == 0) *fc.done = 0;
    const uint4* local = pr.win[pr.rank].ll + (size_t)(seq & 1ULL) * pr.world * pr.cap;
    const unsigned long long t0 = global_ns();
    for (int q = threadIdx.x; q < P; q += blockDim.x) {
        // the entries of (up to) eight ranks are requested together (independent loads), re-polled until every one carries
        // this call's sequence number, and added in rank order
        double s = 0.0;
        for (int r0 = 0; r0 < pr.world; r0 += 8) {
            double v[8];
            const int nr = min(8, pr.world - r0);
            const unsigned all = (1u << nr) - 1u;
            unsigned got = 0u;
            for (;;) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (r < nr && !((got >> r) & 1u)) {
                        double t;
                        if (ll_load(local + (size_t)(r0 + r) * pr.cap + q, (unsigned)seq, t)) { v[r] = t; got |= 1u << r; }
                    }
                }
                if (got == all) break;
                if (global_ns() - t0 > fc.timeout_ns) {
                    s_timeout = 1;
#pragma unroll
                    for (int r = 0; r < 8; ++r)
                        if (r < nr && !((got >> r) & 1u)) v[r] = __longlong_as_double(0x7ff8000000000000LL);
                    break;
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < nr) s += v[r];
        }
        out_pset[q] = s;
        if (mh_on && !s_timeout) mh_accept_chain(mh, mh_it, q);
    }
    __syncthreads();
#ifdef GCDYN_COMM_TRACE
    if (threadIdx.x == 0) printf("rank %d seq %llu: last CTA arrived, then %llu ns until every entry was in and summed\n", pr.rank, seq, global_ns() - tr0);
#endif
    if (threadIdx.x == 0) {
        if (s_timeout && fc.err) *fc.err = 1;
        *pr.seq_dev = seq;
        if (mh_on) *mh.it_dev = mh_it;
    }
}

}  // namespace gcdyn
