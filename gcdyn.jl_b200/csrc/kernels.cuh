// Device kernels of libgcdyn_b200 (sm_100a, FP64).  See DESIGN.md §3-§5.
//
// Pipeline of one batched ODE likelihood evaluation (P parameter sets × T trees):
//   prep_kernel   per pset: λ_i (table / constant / sigmoid, mbp.jl:27-43), Λ_i, node-term logs
//   traj_kernel   per pset (one warp): Taylor-series integration of the extinction ODE
//                 (dp_dt!, mbp.jl:266-278) on a uniform grid of S cells; emits, per cell, the
//                 polynomial of F_x(τ) = ∫(−Λ_x + 2 λ_x p_x)  (the branch density of dp_logq_dt!,
//                 mbp.jl:285-300) and log(1 − p_x(T)) (mbp.jl:244-245)
//   sweep_kernel  per (tree, pset): Σ over lookup items of w·F_type(τ) + node terms
//                 (mbp.jl:167-191), warp-shuffle reduction → out_tree[p*T + t]
//   reduce_kernel per pset: fixed-order sum over trees → out_pset[p]   (mbp.jl:258)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

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
    const int32_t* it_tw;        // [n_items] type | (int8 weight << 24)
    const int32_t* it_aux;       // [n_items] index into nt row, −1 = none
    const int32_t* n_death;      // [T]
    const int32_t* n_surv;       // [T]
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ double count_times_log(int n, double lg) {
    return n == 0 ? 0.0 : (double)n * lg;      // 0 · (−Inf) must stay 0: the reference never forms that term
}

// ------------------------------------------------------------------------------------------------
// prep_kernel: one CTA per pset.
// ------------------------------------------------------------------------------------------------
__global__ void prep_kernel(ParamsView pv, PsetPack pk, int method) {
    const int p = blockIdx.x;
    const int K = pv.K;
    const double mu = pv.mu[p], delta = pv.delta[p], rho = pv.rho[p], sigma = pv.sigma[p];
    const double* G = pv.Gamma + (size_t)p * pv.gamma_stride;   // column-major: Γ[i,j] = G[i + j*K]
    double* nt = pk.nt + (size_t)p * pk.ntw;
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        double lam;
        if (pv.response == 0)      lam = pv.lambda[(size_t)p * K + i];
        else if (pv.response == 1) lam = pv.lambda[p];
        else {   // sigmoid(x, xscale, xshift, yscale, yshift), mbp.jl:4-7
            const double e = exp(-(pv.xscale[p] * (pv.type_space[i] - pv.xshift[p])));
            lam = pv.yscale[p] * (1.0 / (1.0 + e)) + pv.yshift[p];
        }
        const double gam = delta * -G[i + (size_t)i * K];
        const double Lam = lam + mu + gam;
        pk.lam[(size_t)p * K + i] = lam;
        pk.Lam[(size_t)p * K + i] = Lam;
        nt[i] = log(lam);
        if (method == 2 || method == 3) {   // Stadler per-type constants, extra_likelihoods.jl:50-57
            const double c = sqrt(Lam * Lam - 4.0 * mu * (1.0 - sigma) * lam);
            const double x = (-Lam - c) / 2.0, y = (-Lam + c) / 2.0;
            const double A = y + lam * (1.0 - rho), B = x + lam * (1.0 - rho);
            pk.st_c[(size_t)p * K + i] = c;
            pk.st_A[(size_t)p * K + i] = A;
            pk.st_B[(size_t)p * K + i] = B;
            // conditioning probability for a root of this type, extra_likelihoods.jl:82-99 (root_time = present_time − 0 is applied by the caller through e)
        }
    }
    for (int e = threadIdx.x; e < K * K; e += blockDim.x) {
        const int x = e / K, y = e - x * K;
        nt[K + e] = (x == y) ? -CUDART_INF : log(delta * G[x + (size_t)y * K]);   // γ(model, from, to), mbp.jl:84-97
    }
    if (threadIdx.x == 0) {
        nt[K + K * K] = log(mu) + log(sigma);   // mbp.jl:171
        nt[K + K * K + 1] = log(rho);           // mbp.jl:168
        pk.pstatus[p] = 0;
        pk.perr[p] = 0.0;
    }
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

// ------------------------------------------------------------------------------------------------
// traj_kernel: one warp per pset; lane l owns types l, l+32, … (KPL per lane).
//
// Taylor recurrence for the h-scaled coefficients ĉ_n of p(τ_s + θh) = Σ ĉ_n θ^n:
//   ĉ_{n+1} = h/(n+1) · [ −(λ+μ) ĉ_n + [n=0] μ(1−σ) + λ Σ_{m≤n} ĉ_m ĉ_{n−m} + δ Γ ĉ_n ]
// which is dp_dt! (mbp.jl:266-278) expanded order by order.  Table row for cell s:
//   Fc[s][0] = F(τ_s),  Fc[s][1] = h(2λĉ_0 − Λ),  Fc[s][n+1] = 2hλ ĉ_n/(n+1).
// Table layout: tab[((p*S + s)*(M+2) + n)*K + x].
// ------------------------------------------------------------------------------------------------
template <int M, int KPL>
__global__ void __launch_bounds__(32) traj_kernel(ParamsView pv, PsetPack pk, double present_time, int S,
                                                  double* __restrict__ tab) {
    extern __shared__ double smem[];
    const int p = blockIdx.x;
    const int K = pv.K;
    const int lane = threadIdx.x;
    double* GT = smem;                 // GT[j*K + i] = δ Γ[i,j]   (lane-contiguous in i)
    double* cb = smem + (size_t)K * K; // 2 × K broadcast buffer for ĉ_n

    const double mu = pv.mu[p], delta = pv.delta[p], rho = pv.rho[p], sigma = pv.sigma[p];
    const double* G = pv.Gamma + (size_t)p * pv.gamma_stride;
    for (int e = lane; e < K * K; e += 32) GT[e] = delta * G[e];   // column-major source: G[i + j*K] → GT[j*K + i]
    const double h = present_time / (double)S;
    const double b = mu * (1.0 - sigma);

    double lam[KPL], a[KPL], Lam[KPL], pcur[KPL], Fcur[KPL];
    bool own[KPL];
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int i = lane + 32 * q;
        own[q] = i < K;
        lam[q] = own[q] ? pk.lam[(size_t)p * K + i] : 0.0;
        Lam[q] = own[q] ? pk.Lam[(size_t)p * K + i] : 0.0;
        a[q] = lam[q] + mu;
        pcur[q] = own[q] ? 1.0 - rho : 0.0;      // p(0) = 1 − ρ, mbp.jl:130,141,226
        Fcur[q] = 0.0;
    }
    __syncwarp();

    double errmax = 0.0;
    bool bad = false;
    double* trow = tab + (size_t)p * S * (M + 2) * K;
    for (int s = 0; s < S; ++s) {
        double c[M + 1][KPL];
#pragma unroll
        for (int q = 0; q < KPL; ++q) c[0][q] = pcur[q];
#pragma unroll
        for (int n = 0; n < M; ++n) {
            double* buf = cb + (n & 1) * K;
#pragma unroll
            for (int q = 0; q < KPL; ++q)
                if (own[q]) buf[lane + 32 * q] = c[n][q];
            __syncwarp();
            double dot[KPL][2];
#pragma unroll
            for (int q = 0; q < KPL; ++q) dot[q][0] = dot[q][1] = 0.0;
            int j = 0;
            for (; j + 1 < K; j += 2) {
                const double v0 = buf[j], v1 = buf[j + 1];
#pragma unroll
                for (int q = 0; q < KPL; ++q) {
                    const int i = own[q] ? lane + 32 * q : 0;
                    dot[q][0] = fma(GT[(size_t)j * K + i], v0, dot[q][0]);
                    dot[q][1] = fma(GT[(size_t)(j + 1) * K + i], v1, dot[q][1]);
                }
            }
            if (j < K) {
                const double v0 = buf[j];
#pragma unroll
                for (int q = 0; q < KPL; ++q) {
                    const int i = own[q] ? lane + 32 * q : 0;
                    dot[q][0] = fma(GT[(size_t)j * K + i], v0, dot[q][0]);
                }
            }
            const double hn = h / (double)(n + 1);
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                double conv = 0.0;
#pragma unroll
                for (int m = 0; m <= n; ++m) conv = fma(c[m][q], c[n - m][q], conv);
                double rhs = fma(lam[q], conv, dot[q][0] + dot[q][1]);
                rhs = fma(-a[q], c[n][q], rhs);
                if (n == 0) rhs += b;
                c[n + 1][q] = rhs * hn;
            }
        }
        // emit the F polynomial of this cell, advance p and F to the next grid point
        double* cell = trow + (size_t)s * (M + 2) * K;
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const int i = lane + 32 * q;
            double f1 = h * (2.0 * lam[q] * c[0][q] - Lam[q]);
            double psum = 0.0, fsum = 0.0;
            if (own[q]) { cell[i] = Fcur[q]; cell[K + i] = f1; }
#pragma unroll
            for (int n = M; n >= 1; --n) {
                const double fc = (2.0 * h * lam[q]) * c[n][q] / (double)(n + 1);
                if (own[q]) cell[(size_t)(n + 1) * K + i] = fc;
                fsum += fc;
                psum += c[n][q];
            }
            Fcur[q] += fsum + f1;
            pcur[q] = psum + c[0][q];
            const double ind = fabs(c[M][q]) + fabs(c[M - 1][q]);
            if (own[q]) {
                if (!(ind <= errmax)) errmax = ind;                 // NaN-propagating max
                if (!(pcur[q] >= -1e-9 && pcur[q] <= 1.0 + 1e-9)) bad = true;   // isoutofdomain, mbp.jl:146,204,231
            }
        }
    }
    // conditioning term and diagnostics
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int i = lane + 32 * q;
        if (own[q]) pk.logcond[(size_t)p * K + i] = log1p(-pcur[q]);   // log(1 − p_i(T)), mbp.jl:244-245
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(FULL, errmax, o);
        if (!(other <= errmax)) errmax = other;
    }
    const bool anybad = __any_sync(FULL, bad) || !(errmax == errmax) || isinf(errmax);
    if (lane == 0) {
        pk.perr[p] = errmax;
        if (anybad) pk.pstatus[p] |= ST_SOLVER;
    }
}

// ------------------------------------------------------------------------------------------------
// sweep_kernel: grid (chunk, pset); a warp per tree inside the chunk, lanes stride over items.
//   LL(t,p) = Σ_items [ w·F_type(τ) + nt[aux] ] + n_death·(log μ + log σ) + n_surv·log ρ − logcond[root_type]
// F by Horner on the cell polynomial; θ = τ·S/T − s.
// ------------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(256) sweep_kernel(ForestView fv, PsetPack pk, const double* __restrict__ tab,
                                                    const int64_t* __restrict__ chunk_off, double present_time, int S,
                                                    double* __restrict__ out_tree, int32_t* __restrict__ status) {
    const int p = blockIdx.y;
    const int K = fv.K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const double* tp = tab + (size_t)p * S * (M + 2) * K;
    const double* nt = pk.nt + (size_t)p * pk.ntw;
    const double lmusig = nt[K + K * K], lrho = nt[K + K * K + 1];
    const int pst = pk.pstatus[p];
    const double scale = (double)S / present_time;
    const int64_t t_lo = chunk_off[blockIdx.x], t_hi = chunk_off[blockIdx.x + 1];
    for (int64_t t = t_lo + warp; t < t_hi; t += nwarps) {
        const int64_t i0 = fv.item_off[t], i1 = fv.item_off[t + 1];
        double acc = 0.0;
        for (int64_t i = i0 + lane; i < i1; i += 32) {
            const double tau = present_time - fv.it_time[i];
            const int tw = fv.it_tw[i];
            const int aux = fv.it_aux[i];
            const int x = tw & 0xffffff;
            const double w = (double)(tw >> 24);
            const double u = tau * scale;
            int s = (int)u;
            s = s > S - 1 ? S - 1 : (s < 0 ? 0 : s);
            const double th = u - (double)s;
            const double* cell = tp + ((size_t)s * (M + 2)) * K + x;
            double f = __ldg(cell + (size_t)(M + 1) * K);
#pragma unroll
            for (int n = M; n >= 0; --n) f = fma(f, th, __ldg(cell + (size_t)n * K));
            acc = fma(w, f, acc);
            if (aux >= 0) acc += __ldg(nt + aux);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            double r = acc + count_times_log(fv.n_death[t], lmusig) + count_times_log(fv.n_surv[t], lrho)
                       - pk.logcond[(size_t)p * K + fv.root_type[t]];
            int st = pst;
            if (fv.self_loop[t]) { st |= ST_SELF_LOOP; }
            if (st & (ST_SELF_LOOP | ST_SOLVER)) r = -CUDART_INF;
            else if (r != r) st |= ST_DOMAIN;
            out_tree[(size_t)p * fv.n_trees + t] = r;
            if (status) status[(size_t)p * fv.n_trees + t] = st;
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
            if (method == 2) r -= pk.logcond[(size_t)p * K + fv.root_type[t]];
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

}  // namespace gcdyn
