// Micro-benchmark of the round-2 Taylor stage loop (derivative variables, folded diagonal, a-posteriori cell width),
// against the round-1 loop.  Not part of the library.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_trace/traj_micro2 tools/traj_micro2.cu
#include <cstdio>
#include <cuda_runtime.h>

__host__ __device__ constexpr double factd(int n) { double f = 1.0; for (int i = 2; i <= n; ++i) f *= (double)i; return f; }

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

// V = 0: round-1 loop (scaled coefficients, convolution and -a c_n on the critical path, width given)
// V = 1: derivative variables d_n = p^(n), folded diagonal (kappa), seeded dot, A_n = d_n/n! published progressively,
//        width chosen AFTER the derivatives from |A_M|, |A_{M-1}| (REDUX on the high words), Estrin evaluation of p(tau+h)
// V = 2: V = 1 with the width given (no controller) - isolates the stage cost
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_hint(unsigned long long* bar, unsigned parity, unsigned ns) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
template <int M, int KP, int V, int W = 0>
__global__ void cell_loop(double* out, long long* cyc, int cells, double hfix, double lamv, double av, double bv, double gv) {
    __shared__ __align__(16) double cb[2 * KP];
    __shared__ __align__(16) double slot[32 * (M + 2)];
    __shared__ __align__(8) unsigned long long full;
    if (W > 0) {
        if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full))); }
        __syncthreads();
        if (threadIdx.x >= 32) {                     // consumer warps: wait for every cell, do nothing else
            for (int s = 0; s < cells; ++s) {
                const unsigned par = (unsigned)s & 1u;
                if (W == 1) { while (!mbar_try(&full, par)) { } }
                else if (W == 2) { if ((threadIdx.x & 31) == 0) { while (!mbar_try(&full, par)) { } } __syncwarp(); }
                else if (W == 3) { while (!mbar_try_hint(&full, par, 2000u)) { } }
                else if (W == 4) { while (!mbar_try(&full, par)) { __nanosleep(200); } }
            }
            return;
        }
    }
    const int lane = threadIdx.x & 31;
    const bool own = lane < KP;
    double g[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) g[j] = own ? (j == lane ? -gv * (KP - 1) : gv * (1.0 + 0.01 * j + 0.001 * lane)) : 0.0;
    const double lam = lamv * (1.0 + 0.01 * lane), a = av, b = bv;
    for (int e = lane; e < 2 * KP; e += 32) cb[e] = 0.0;
    __syncwarp();
    double pcur = own ? 0.5 : 0.0;
    double hsum = 0.0;
    const long long t0 = clock64();
    for (int s = 0; s < cells; ++s) {
        if constexpr (V == 0) {
            const double hc = hfix;
            double c[M + 1];
            c[0] = pcur;
#pragma unroll
            for (int n = 0; n < M; ++n) {
                double* buf = cb + (n & 1) * KP;
                if (own) buf[lane] = c[n];
                __syncwarp();
                double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                for (int j = 0; j < KP; j += 4) {
                    const double2 v01 = *reinterpret_cast<const double2*>(buf + j);
                    const double2 v23 = *reinterpret_cast<const double2*>(buf + j + 2);
                    d0 = fma(g[j], v01.x, d0); d1 = fma(g[j + 1], v01.y, d1);
                    d2 = fma(g[j + 2], v23.x, d2); d3 = fma(g[j + 3], v23.y, d3);
                }
                const double dot = (d0 + d1) + (d2 + d3);
                const double hn = hc * (1.0 / (double)(n + 1));
                double cv0 = 0.0, cv1 = 0.0;
#pragma unroll
                for (int m = 0; 2 * m < n; ++m) { if (m & 1) cv1 = fma(c[m], c[n - m], cv1); else cv0 = fma(c[m], c[n - m], cv0); }
                double conv = 2.0 * (cv0 + cv1);
                if ((n & 1) == 0) conv = fma(c[n / 2], c[n / 2], conv);
                double rhs = fma(lam, conv, dot);
                rhs = fma(-a, c[n], rhs);
                if (n == 0) rhs += b;
                c[n + 1] = rhs * hn;
            }
            double ps[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int n = M; n >= 1; --n) ps[n & 3] += c[n];
            pcur = ((ps[0] + ps[1]) + (ps[2] + ps[3])) + c[0];
            hsum += hc;
        } else {
            double A[M + 1];
            A[0] = pcur;
            double dn = pcur;
            const double kappa = fma(2.0 * lam, dn, -a);
            double seed = fma(dn, fma(lam, dn, -a), b);      // stage 0: lam d0^2 - a d0 + b
#pragma unroll
            for (int n = 0; n < M; ++n) {
                double* buf = cb + (n & 1) * KP;
                if (own) { buf[lane] = dn; slot[lane * (M + 2) + n] = A[n]; }
                __syncwarp();
                // off the critical path: lam n! R_{n+1}, R_k = sum_{m=1}^{k-1} A_m A_{k-m} (needs A_1..A_n only)
                double nextseed_add = 0.0;
                if (n >= 1) {
                    constexpr int dummy = 0; (void)dummy;
                    const int k = n + 1;
                    double r0 = 0.0, r1 = 0.0;
#pragma unroll
                    for (int m = 1; 2 * m < k; ++m) { if (m & 1) r1 = fma(A[m], A[k - m], r1); else r0 = fma(A[m], A[k - m], r0); }
                    double R = 2.0 * (r0 + r1);
                    if ((k & 1) == 0) R = fma(A[k / 2], A[k / 2], R);
                    nextseed_add = (lam * factd(k)) * R;
                }
                double d0 = seed, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                for (int j = 0; j < KP; j += 4) {
                    const double2 v01 = *reinterpret_cast<const double2*>(buf + j);
                    const double2 v23 = *reinterpret_cast<const double2*>(buf + j + 2);
                    d0 = fma(g[j], v01.x, d0); d1 = fma(g[j + 1], v01.y, d1);
                    d2 = fma(g[j + 2], v23.x, d2); d3 = fma(g[j + 3], v23.y, d3);
                }
                dn = (d0 + d1) + (d2 + d3);                  // d_{n+1}
                A[n + 1] = dn * (1.0 / factd(n + 1));
                seed = fma(kappa, dn, nextseed_add);         // seed of stage n+1 (in the shadow of its broadcast)
            }
            if (own) slot[lane * (M + 2) + M] = A[M];
            double h = hfix;
            if constexpr (V == 1) {
                const unsigned hiM = __reduce_max_sync(0xffffffffu, own ? (unsigned)__double2hiint(fabs(A[M])) : 0u);
                const unsigned hiM1 = __reduce_max_sync(0xffffffffu, own ? (unsigned)__double2hiint(fabs(A[M - 1])) : 0u);
                const float lM = (float)((int)hiM - 0x3ff00000) * (1.0f / 1048576.0f);
                const float lM1 = (float)((int)hiM1 - 0x3ff00000) * (1.0f / 1048576.0f);
                const float la = (-38.5f - lM) * (1.0f / (float)M), lb = (-38.5f - lM1) * (1.0f / (float)(M - 1));
                const double hctl = (double)exp2f(fminf(la, lb));
                h = fmin(fmax(hctl, hfix), hfix * 1.0000001);
            }
            pcur = estrin<M>(A, h);
            hsum += h;
        }
        if (W > 0) { __syncwarp(); if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full)) : "memory"); }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = pcur + 1e-30 * (hsum + slot[lane]);
}

template <int M, int V, int W = 0, int KPX = 20>
void run(const char* name, double* out, long long* cyc, int nblocks, int nthreads) {
    const int cells = 2000;
    long long h;
    double res[32];
    cell_loop<M, KPX, V, W><<<nblocks, nthreads>>>(out, cyc, cells, 0.05, 1.3, 1.8, 0.4, 0.02);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(res, out, 32 * 8, cudaMemcpyDeviceToHost);
    printf("%-52s M=%d: %.0f cycles per cell, %.1f per order   p[0]=%.15f p[19]=%.15f\n", name, M, (double)h / cells,
           (double)h / cells / M, res[0], res[19]);
}

int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
    for (int rep = 0; rep < 2; ++rep) {
        run<16, 0>("V0 round-1 loop", out, cyc, 1, 32);
        run<16, 2>("V2 derivative variables, width given", out, cyc, 1, 32);
        run<16, 1>("V1 derivative variables + a-posteriori width", out, cyc, 1, 32);
        run<20, 0>("V0 round-1 loop", out, cyc, 1, 32);
        run<20, 1>("V1 derivative variables + a-posteriori width", out, cyc, 1, 32);
        run<24, 0>("V0 round-1 loop", out, cyc, 1, 32);
        run<24, 1>("V1 derivative variables + a-posteriori width", out, cyc, 1, 32);
        run<16, 1>("V1, 256 CTAs", out, cyc, 256, 32);
        run<16, 1, 1>("V1 + 3 warps try_wait spin (all lanes)", out, cyc, 1, 128);
        run<16, 1, 2>("V1 + 3 warps, lane 0 polls", out, cyc, 1, 128);
        run<16, 1, 3>("V1 + 3 warps try_wait with suspend hint", out, cyc, 1, 128);
        run<16, 1, 4>("V1 + 3 warps try_wait + nanosleep(200)", out, cyc, 1, 128);
        run<16, 0, 1>("V0 + 3 warps try_wait spin (all lanes)", out, cyc, 1, 128);
        run<16, 1, 1>("V1 + spin, 256 CTAs", out, cyc, 256, 128);
        run<16, 1, 4>("V1 + nanosleep, 256 CTAs", out, cyc, 256, 128);
    }
    run<16, 2, 0, 4>("V2 K=4", out, cyc, 1, 32);
    run<16, 2, 0, 8>("V2 K=8", out, cyc, 1, 32);
    run<16, 2, 0, 12>("V2 K=12", out, cyc, 1, 32);
    run<16, 2, 0, 16>("V2 K=16", out, cyc, 1, 32);
    run<16, 2, 0, 20>("V2 K=20", out, cyc, 1, 32);
    run<16, 2, 0, 24>("V2 K=24", out, cyc, 1, 32);
    run<16, 2, 0, 32>("V2 K=32", out, cyc, 1, 32);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
