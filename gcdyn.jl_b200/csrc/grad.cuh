// Exact gradients of the per-pset log-likelihood by FORWARD SENSITIVITIES of the trajectory.
//
// The reference is parametric in `T<:Real` so that ForwardDiff can push dual numbers through `loglikelihood`
// (types.jl:92, mbp.jl:118-122); behind a Float64 C ABI the same derivatives come from integrating, next to
// p(τ) (dp_dt!, mbp.jl:266-278), the sensitivities s^θ = ∂p/∂θ of the autodiff-able fields — the birth-response
// parameters, μ and δ:
//     ds/dτ = G s − (λ+μ) s + 2λ p s + φ_θ(p),      s(0) = 0          (p(0) = 1 − ρ does not depend on θ)
//     φ_θ = w (p² − p)   for a response parameter with ∂λ_i/∂θ = w_i,      (1−σ) − p   for μ,      Γ p = G p / δ   for δ.
// The system is polynomial like the one for p, so its derivatives at the start of a cell follow the same kind of
// recurrence (one K-term dot product with the row of δΓ per stage, plus convolutions with p's coefficients), the cell
// width is chosen a posteriori from the last two coefficients of ALL vectors, and on each cell
//     ∂F_x/∂θ (τ) = ∫ ( −∂Λ_x/∂θ + 2 (∂λ_x/∂θ) p_x + 2 λ_x s_x )
// is a polynomial exactly like F_x (dp_logq_dt!, mbp.jl:285-300).  The per-pset sum over the forest needs no per-tree
// work at all: Σ_trees Σ_items w·F_x(τ) = Σ_k F_x(τ_k)·colsum[k][x] with the forest's nodal Chebyshev moments summed over
// the trees (cheb_gemm.cuh), and the same contraction with ∂F_x/∂θ(τ_k) gives the gradient; the node-term logs and the
// conditioning −log(1 − p_root(T)) (mbp.jl:244-245) are differentiated in closed form.
//
// One CTA of 8 warps per pset: warp 0 integrates p, warps 1..NS one sensitivity each (lane = type, K ≤ 32), in lockstep
// (one named barrier per Taylor stage: every vector's d_n is broadcast through shared memory); after each cell all
// threads turn the cell's coefficients into the polynomials of F and ∂F/∂θ, evaluate them at the Chebyshev nodes inside
// the cell and accumulate value × column sum.  Nothing but [P] log-likelihoods and [P × NS] derivatives leaves the kernel.
#pragma once
#include "kernels.cuh"

namespace gcdyn {

constexpr int GRAD_THREADS = 256;
constexpr int GRAD_MAX_DIRS = 6;       // sigmoid response: xscale, xshift, yscale, yshift, μ, δ;  constant: λ, μ, δ

struct GradCfg {
    double T, tol;
    int D;                 // degree of the forest's nodal grid
    int NS;                // number of directions (3 or 6)
    int compact_tc;        // shared Γ: the type changes enter through colsum[2K+2]·log δ + tc_total
    int any_self_loop;
    MomentView mv;
    const double* tc_total;   // [1] Σ_t Σ count·log Γ_xy (shared Γ) or nullptr
    double* out_ll;        // [P]
    double* out_grad;      // [P][NS]
    int* out_status;       // [P] ST_SOLVER when the trajectory failed
};

enum { GK_P = 0, GK_RESP = 1, GK_MU = 2, GK_DELTA = 3 };

template <int M, int KP>
__global__ void __launch_bounds__(GRAD_THREADS, 1) traj_grad_kernel(ParamsView pv, GradCfg gc) {
    extern __shared__ __align__(16) double gsm_[];
    constexpr int L = M + 2;
    constexpr int NV = 1 + GRAD_MAX_DIRS;               // vectors: p and up to six sensitivities
    const int K = pv.K, D = gc.D, NS = gc.NS, NI = 1 + NS;
    const int KS = KP;                                  // KP is K rounded up to a multiple of 4
    const int p = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // shared memory: broadcast buffers [2][NV][KS] | cell coefficients [NV][M+1][KS] | τ_k [D+2] | accumulators [NV][KS]
    double* bc = gsm_;
    double* Acell = bc + 2 * NV * KS;
    double* tauk = Acell + NV * (M + 1) * KS;
    double* accb = tauk + ((D + 3) & ~1);
    __shared__ unsigned s_hw[NV][2];
    __shared__ double s_h, s_t0, s_pT[32];
    __shared__ int s_last, s_failed;
    const MomentView& mv = gc.mv;
    const int C1 = mv.C1;

    // ---- parameters (every load before the first use)
    const double* G = pv.Gamma + (size_t)p * pv.gamma_stride;
    const double mu = pv.mu[p], delta = pv.delta[p], rho = pv.rho[p], sigma = pv.sigma[p];
    const bool own = lane < K;
    const int li = own ? lane : 0;
    double rs0 = 0.0, rs1 = 0.0, rs2 = 0.0, rs3 = 0.0;
    if (pv.response == 2) { rs0 = pv.xscale[p]; rs1 = pv.xshift[p]; rs2 = pv.yscale[p]; rs3 = pv.yshift[p]; }
    const double in0 = pv.response == 0 ? pv.lambda[(size_t)p * K + li] : (pv.response == 1 ? pv.lambda[p] : pv.type_space[li]);
    const double gdiag = G[li + (size_t)li * K];
    double g[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) g[j] = G[li + (size_t)(j < K ? j : 0) * K];
    for (int m = tid; m <= D; m += GRAD_THREADS) tauk[m] = gc.T * (1.0 + cospi((double)m / (double)D)) * 0.5;
    // λ_i and its derivatives with respect to the response parameters (sigmoid, mbp.jl:4-7,41-43)
    double lam = in0, wresp[4] = {1.0, 0.0, 0.0, 0.0};
    if (pv.response == 2) {
        const double z = rs0 * (in0 - rs1);
        const double sg = 1.0 / (1.0 + exp(-z));
        lam = rs2 * sg + rs3;
        const double dsg = sg * (1.0 - sg);
        wresp[0] = rs2 * dsg * (in0 - rs1);             // ∂λ/∂xscale
        wresp[1] = -rs2 * dsg * rs0;                    // ∂λ/∂xshift
        wresp[2] = sg;                                  // ∂λ/∂yscale
        wresp[3] = 1.0;                                 // ∂λ/∂yshift
    }
    if (!own) lam = 0.0;
    const int NR = pv.response == 2 ? 4 : 1;            // response directions; then μ, then δ
#pragma unroll
    for (int j = 0; j < KP; ++j) g[j] = (own && j < K) ? delta * g[j] : 0.0;
    const double a = lam + mu, b = mu * (1.0 - sigma);
    const double gam = own ? delta * -gdiag : 0.0;
    double rate = own ? fabs(lam) + fabs(mu) + 2.0 * fabs(gam) : 0.0;
    rate = warp_max_nan(rate);
    // this warp's vector
    const int v = warp;                                 // 0 = p, 1..NS = sensitivities, others help only after each cell
    const bool integ = v < NI;
    int kind = GK_P;
    double w = 0.0;
    if (v >= 1 && v <= NR) { kind = GK_RESP; w = wresp[v - 1]; }
    else if (v == NR + 1) kind = GK_MU;
    else if (v == NR + 2) kind = GK_DELTA;
    if (!own) w = 0.0;
    for (int e = tid; e < 2 * NV * KS; e += GRAD_THREADS) bc[e] = 0.0;
    if (tid == 0) { s_failed = 0; s_last = 0; }
    __syncthreads();

    const unsigned tol_hi = (unsigned)__double2hiint(gc.tol);
    (void)tol_hi;
    const float l2aim = (float)((int)(unsigned)__double2hiint(0.25 * gc.tol) - 0x3ff00000) * (1.0f / 1048576.0f);
    double hcap = hrate_cap(M) / rate;
    if (!(hcap <= gc.T)) hcap = gc.T;
    double cur = (v == 0 && own) ? 1.0 - rho : 0.0;     // p(0) = 1 − ρ;  s(0) = 0
    double tau = 0.0;
    // per-thread state of the contraction: thread t < NI·K owns (vector, type) = (t / K, t % K)
    const int cv = tid / K, cx = tid - cv * K;
    const bool citem = tid < NI * K;
    double c_run = 0.0, c_acc = 0.0;                    // running ∫ at the cell start; Σ value × column sum
    int kn = D;
    // constants of the contraction item (they belong to type cx, not to this thread's lane)
    double c_lam = 0.0, c_Lam = 0.0, c_w = 0.0, c_dLam = 0.0;
    if (citem) {
        double l = pv.response == 0 ? pv.lambda[(size_t)p * K + cx] : (pv.response == 1 ? pv.lambda[p] : pv.type_space[cx]);
        double wr[4] = {1.0, 0.0, 0.0, 0.0};
        if (pv.response == 2) {
            const double t_i = l, z = rs0 * (t_i - rs1), sg = 1.0 / (1.0 + exp(-z)), dsg = sg * (1.0 - sg);
            l = rs2 * sg + rs3;
            wr[0] = rs2 * dsg * (t_i - rs1); wr[1] = -rs2 * dsg * rs0; wr[2] = sg; wr[3] = 1.0;
        }
        const double gx = delta * -G[cx + (size_t)cx * K];
        c_lam = l;
        c_Lam = l + mu + gx;
        if (cv >= 1 && cv <= NR) { c_w = wr[cv - 1]; c_dLam = c_w; }         // ∂Λ_x/∂θ = ∂λ_x/∂θ
        else if (cv == NR + 1) c_dLam = 1.0;                                  // ∂Λ_x/∂μ
        else if (cv == NR + 2) c_dLam = gx / delta;                           // ∂Λ_x/∂δ = γ_x/δ
    }

    for (int cell = 0;; ++cell) {
        double A[M + 1], Ap[M + 1];                     // this vector's coefficients; p's (own type), for the sensitivities
        if (integ) {
            double dn = cur;
            A[0] = cur;
            double kappa = 0.0, seed = 0.0;
            if (kind == GK_P) {
                kappa = fma(2.0 * lam, dn, -a);
                seed = fma(dn, fma(lam, dn, -a), b);
            }
#pragma unroll
            for (int n = 0; n < M; ++n) {
                double* buf = bc + ((n & 1) * NV + v) * KS;
                if (own) buf[lane] = dn;
                asm volatile("bar.sync 1, %0;" ::"r"(NI * 32) : "memory");
                const double* bp = bc + ((n & 1) * NV) * KS;        // p's d_n
                double extra = 0.0;                                  // everything but G·d_n
                if (kind != GK_P) {
                    const double dpn = bp[li];
                    Ap[n] = dpn * (1.0 / factd(n));
                    double c0 = 0.0, c1 = 0.0;                       // Σ_{m ≤ n} A^p_m A^s_{n−m}
#pragma unroll
                    for (int m = 0; m <= n; ++m) { if (m & 1) c1 = fma(Ap[m], A[n - m], c1); else c0 = fma(Ap[m], A[n - m], c0); }
                    extra = fma(2.0 * lam * factd(n), c0 + c1, -a * dn);
                    if (kind == GK_RESP) {
                        double q0 = 0.0, q1 = 0.0;                   // (p²)_n = Σ_{m ≤ n} A^p_m A^p_{n−m}
#pragma unroll
                        for (int m = 0; 2 * m < n; ++m) { if (m & 1) q1 = fma(Ap[m], Ap[n - m], q1); else q0 = fma(Ap[m], Ap[n - m], q0); }
                        double q = 2.0 * (q0 + q1);
                        if ((n & 1) == 0) q = fma(Ap[n / 2], Ap[n / 2], q);
                        extra = fma(w, fma(factd(n), q, -dpn), extra);
                    } else if (kind == GK_MU) {
                        extra -= dpn;
                        if (n == 0) extra += 1.0 - sigma;
                    }
                }
                double d0 = kind == GK_P ? seed : extra, d1 = 0.0, d2 = 0.0, d3 = 0.0;
                const double* bs = bc + ((n & 1) * NV + v) * KS;
#pragma unroll
                for (int j = 0; j < KP; j += 4) {
                    const double2 v01 = *reinterpret_cast<const double2*>(bs + j);
                    const double2 v23 = *reinterpret_cast<const double2*>(bs + j + 2);
                    d0 = fma(g[j], v01.x, d0);
                    d1 = fma(g[j + 1], v01.y, d1);
                    d2 = fma(g[j + 2], v23.x, d2);
                    d3 = fma(g[j + 3], v23.y, d3);
                }
                if (kind == GK_DELTA) {                               // φ_δ = Γ p = (δΓ p)/δ
                    double e0 = 0.0, e1 = 0.0;
#pragma unroll
                    for (int j = 0; j < KP; j += 2) {
                        const double2 u = *reinterpret_cast<const double2*>(bp + j);
                        e0 = fma(g[j], u.x, e0);
                        e1 = fma(g[j + 1], u.y, e1);
                    }
                    d1 = fma(e0 + e1, 1.0 / delta, d1);
                }
                const double dnew = (d0 + d1) + (d2 + d3);
                if (kind == GK_P) {
                    double add = 0.0;
                    if (n >= 1) {
                        const int k = n + 1;
                        double r0 = 0.0, r1 = 0.0;
#pragma unroll
                        for (int m = 1; 2 * m < k; ++m) { if (m & 1) r1 = fma(A[m], A[k - m], r1); else r0 = fma(A[m], A[k - m], r0); }
                        double R = 2.0 * (r0 + r1);
                        if ((k & 1) == 0) R = fma(A[k / 2], A[k / 2], R);
                        add = (lam * factd(k)) * R;
                    }
                    seed = fma(kappa, dnew, add);
                }
                dn = dnew;
                A[n + 1] = dnew * (1.0 / factd(n + 1));
            }
            // ---- width: the last two coefficients of every vector
            unsigned hiM = own ? (unsigned)__double2hiint(fabs(A[M])) : 0u, hiM1 = own ? (unsigned)__double2hiint(fabs(A[M - 1])) : 0u;
            hiM = __reduce_max_sync(FULL, hiM);
            hiM1 = __reduce_max_sync(FULL, hiM1);
            if (lane == 0) { s_hw[v][0] = hiM; s_hw[v][1] = hiM1; }
            asm volatile("bar.sync 1, %0;" ::"r"(NI * 32) : "memory");
            if (v == 0) {
                unsigned m0 = 0u, m1 = 0u;
                for (int u = 0; u < NI; ++u) { m0 = max(m0, s_hw[u][0]); m1 = max(m1, s_hw[u][1]); }
                const float lM = (float)((int)m0 - 0x3ff00000) * (1.0f / 1048576.0f);
                const float lM1 = (float)((int)m1 - 0x3ff00000) * (1.0f / 1048576.0f);
                float lh = fminf((l2aim - lM) * (1.0f / (float)M), (l2aim - lM1) * (1.0f / (float)(M - 1)));
                lh = fminf(lh, 100.0f);
                double h = (double)exp2f(lh);
                bool failed = !(h > gc.T * 1e-10), last = false;
                if (!failed) {
                    if (!(h <= hcap)) h = hcap;
                    const double rem = gc.T - tau;
                    if (h >= rem * (1.0 - 1e-3)) { h = rem; last = true; }
                    else if (h > 0.5 * rem) h = 0.5 * rem;
                    for (int shrinks = 0;; ++shrinks) {            // isoutofdomain: shrink and re-evaluate (mbp.jl:146,204,231)
                        const double pn = estrin<M>(A, h);
                        const bool bad = own && !(pn >= -1e-9 && pn <= 1.0 + 1e-9);
                        if (!__any_sync(FULL, bad)) break;
                        h *= 0.5;
                        last = false;
                        if (shrinks >= 60 || !(h > gc.T * 1e-10)) { failed = true; break; }
                    }
                }
                if (lane == 0) { s_h = h; s_t0 = tau; s_last = last ? 1 : 0; if (failed) s_failed = 1; }
            }
            asm volatile("bar.sync 1, %0;" ::"r"(NI * 32) : "memory");
            const double h = s_h;
            if (!s_failed) cur = estrin<M>(A, h);
            if (own) {
#pragma unroll
                for (int n = 0; n <= M; ++n) Acell[(v * (M + 1) + n) * KS + lane] = A[n];
            }
        }
        __syncthreads();
        if (s_failed) break;
        const double h = s_h, t0 = s_t0, ih = 1.0 / h;
        const bool last = s_last != 0;
        // ---- contraction of this cell: polynomial of F (v = 0) or ∂F/∂θ (v ≥ 1) for (vector, type), evaluated at the
        //      Chebyshev nodes inside the cell, times the forest's column sums
        int klo = kn;
        if (last) klo = -1;
        else while (klo >= 0 && tauk[klo] < t0 + h) --klo;
        if (citem) {
            double co[L];
            const double* Av = Acell + (size_t)cv * (M + 1) * KS + cx;
            const double* Apx = Acell + cx;
            co[0] = c_run;
            double hp = h;
            // integrand coefficients: v = 0: −Λ + 2λ p;  v ≥ 1: −∂Λ/∂θ + 2 (∂λ/∂θ) p + 2 λ s
            {
                const double i0 = cv == 0 ? fma(2.0 * c_lam, Apx[0], -c_Lam)
                                          : fma(2.0 * c_lam, Av[0], fma(2.0 * c_w, Apx[0], -c_dLam));
                co[1] = h * i0;
            }
#pragma unroll
            for (int n = 1; n <= M; ++n) {
                hp *= h;
                const double in = cv == 0 ? 2.0 * c_lam * Apx[(size_t)n * KS]
                                          : fma(2.0 * c_lam, Av[(size_t)n * KS], 2.0 * c_w * Apx[(size_t)n * KS]);
                co[n + 1] = in * hp * (1.0 / (double)(n + 1));
            }
            for (int k = kn; k > klo; --k) {
                const double th = (tauk[k] - t0) * ih;
                const double t2 = th * th;
                double ev = co[L - 2], od = co[L - 1];
#pragma unroll
                for (int n = L - 4; n >= 0; n -= 2) { ev = fma(ev, t2, co[n]); od = fma(od, t2, co[n + 1]); }
                c_acc = fma(fma(od, th, ev), mv.colsum[C1 + k * K + cx], c_acc);
            }
            double fs[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int n = M + 1; n >= 1; --n) fs[n & 3] += co[n];
            c_run += (fs[0] + fs[1]) + (fs[2] + fs[3]);
        }
        kn = klo;
        __syncthreads();                                // Acell is rewritten by the next cell
        if (last) break;
        tau = t0 + h;
    }
    // ---- closing terms and the reduction over types
    const bool failed = s_failed != 0;
    if (v == 0 && own) s_pT[lane] = cur;                // p_x(T)
    if (citem) accb[cv * KS + cx] = c_acc;
    __syncthreads();
    if (integ) {
        double part = 0.0;
        if (own) {
            part = accb[v * KS + lane];
            const double pT = s_pT[lane];
            const double nb = mv.colsum[lane], nroot = mv.colsum[K + 2 + lane];
            if (kind == GK_P) {
                if (nb != 0.0) part += nb * log(lam);                                   // births on type-x branches × log λ_x
                if (nroot != 0.0) part -= nroot * log1p(-pT);                           // −log(1 − p_root(T)), mbp.jl:244-245
            } else {
                if (kind == GK_RESP && nb != 0.0) part += nb * w / lam;
                if (nroot != 0.0) part += nroot * cur / (1.0 - pT);                     // cur = s_x(T)
            }
        }
        part = warp_sum(part);
        if (lane == 0) {
            if (kind == GK_P) {
                const double nd = mv.colsum[K], nsv = mv.colsum[K + 1];
                if (nd != 0.0) part += nd * (log(mu) + log(sigma));
                if (nsv != 0.0) part += nsv * log(rho);
                if (gc.compact_tc) {
                    const double ntc = mv.colsum[2 * K + 2];
                    if (ntc != 0.0) part += ntc * log(delta);
                    if (gc.tc_total) part += *gc.tc_total;
                } else {
                    for (int e = 0; e < K * K; ++e) {
                        const double n = mv.colsum[mv.Jtc + e];
                        if (n != 0.0) { const int x = e / K, y = e - x * K; part += n * log(delta * G[x + (size_t)y * K]); }
                    }
                }
                double ll = part;
                if (failed || gc.any_self_loop) ll = -CUDART_INF;
                gc.out_ll[p] = ll;
                gc.out_status[p] = failed ? ST_SOLVER : 0;
            } else {
                if (kind == GK_MU) { const double nd = mv.colsum[K]; if (nd != 0.0) part += nd / mu; }
                if (kind == GK_DELTA) {
                    double ntc = 0.0;
                    if (gc.compact_tc) ntc = mv.colsum[2 * K + 2];
                    else for (int e = 0; e < K * K; ++e) ntc += mv.colsum[mv.Jtc + e];
                    if (ntc != 0.0) part += ntc / delta;
                }
                gc.out_grad[(size_t)p * NS + (v - 1)] = (failed || gc.any_self_loop) ? CUDART_NAN : part;
            }
        }
    }
}

}  // namespace gcdyn
