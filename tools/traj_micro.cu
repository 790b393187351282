// Micro-benchmark of the Taylor order loop of traj_kernel (one warp, K = 20, M = 16), not part of the library.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_trace/traj_micro tools/traj_micro.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int M, int KP, int V>
__global__ void cell_loop(double* out, long long* cyc, int cells, double hc, double lamv, double av, double bv, double gv) {
    __shared__ __align__(16) double cb[2 * KP];
    const int lane = threadIdx.x & 31;
    const bool own = lane < KP;
    double g[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) g[j] = own ? gv * (1.0 + 0.01 * j + 0.001 * lane) : 0.0;
    const double lam = lamv * (1.0 + 0.01 * lane), a = av, b = bv;
    for (int e = lane; e < 2 * KP; e += 32) cb[e] = 0.0;
    __syncwarp();
    double pcur = own ? 0.5 : 0.0;
    if (blockDim.x == 160 && (threadIdx.x >> 5) != 0 && (threadIdx.x >> 5) != 4) return;   // warps 0 and 4 only: same sub-partition
    const long long t0 = clock64();
    for (int s = 0; s < cells; ++s) {
        double c[M + 1];
        c[0] = pcur;
#pragma unroll
        for (int n = 0; n < M; ++n) {
            double dot = 0.0;
            if constexpr (V == 0) {
                double* buf = cb + (n & 1) * KP;
                if (own) buf[lane] = c[n];
                __syncwarp();
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
                dot = (d0 + d1) + (d2 + d3);
            } else if constexpr (V == 1) {
                double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                for (int j = 0; j < KP; j += 4) {
                    d0 = fma(g[j], __shfl_sync(0xffffffffu, c[n], j), d0);
                    d1 = fma(g[j + 1], __shfl_sync(0xffffffffu, c[n], j + 1), d1);
                    d2 = fma(g[j + 2], __shfl_sync(0xffffffffu, c[n], j + 2), d2);
                    d3 = fma(g[j + 3], __shfl_sync(0xffffffffu, c[n], j + 3), d3);
                }
                dot = (d0 + d1) + (d2 + d3);
            } else if constexpr (V == 2) {
                dot = c[n] * g[0];
            } else if constexpr (V >= 3) {
                double* buf = cb + (n & 1) * KP;
                if (own) buf[lane] = c[n];
                __syncwarp();
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
                dot = (d0 + d1) + (d2 + d3);
            }
            const double hn = hc * (1.0 / (double)(n + 1));
            double conv = 0.0;
            if constexpr (V != 3) {
                double cv0 = 0.0, cv1 = 0.0;
#pragma unroll
                for (int m = 0; 2 * m < n; ++m) {
                    if (m & 1) cv1 = fma(c[m], c[n - m], cv1);
                    else cv0 = fma(c[m], c[n - m], cv0);
                }
                conv = 2.0 * (cv0 + cv1);
                if ((n & 1) == 0) conv = fma(c[n / 2], c[n / 2], conv);
            }
            if constexpr (V == 6) {
                double t = fma(-a, c[n], lam * conv);        // off the matvec's critical path
                if (n == 0) t += b;
                c[n + 1] = fma(dot, hn, t * hn);
            } else {
                double rhs = fma(lam, conv, dot);
                rhs = fma(-a, c[n], rhs);
                if (n == 0) rhs += b;
                c[n + 1] = rhs * hn;
            }
        }
        if constexpr (V == 6) {
            double p0 = c[M], p1 = c[M - 1], p2 = c[M - 2], p3 = c[M - 3];
#pragma unroll
            for (int n = M - 4; n >= 1; n -= 4) { p0 += c[n]; if (n - 1 >= 1) p1 += c[n - 1]; if (n - 2 >= 1) p2 += c[n - 2]; if (n - 3 >= 1) p3 += c[n - 3]; }
            pcur = ((p0 + p1) + (p2 + p3)) + c[0];
        } else {
            double psum = 0.0;
#pragma unroll
            for (int n = M; n >= 1; --n) psum += c[n];
            pcur = psum + c[0];
        }
        if constexpr (V == 4) {   // the product's controller: warp max of the indicator, float pow
            const double ind = fabs(c[M]) + fabs(c[M - 1]);
            const float ind_f = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fabsf((float)ind))));
            const float ratio = exp2f(__log2f(1e-11f / ind_f) * (1.0f / (float)M));
            hc = hc * (double)fminf(fmaxf(0.9f * ratio, 0.999f), 1.001f);
        }
        if constexpr (V == 5 || V == 6) {   // controller on the high word of the indicator: no F2F, no division, no LG2
            const double ind = fabs(c[M]) + fabs(c[M - 1]);
            const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(ind));
            // log2(ind) ≈ (hi − 0x3ff00000) / 2^20 (exponent + linear mantissa, error < 0.09)
            const float l2 = (float)((int)hi - 0x3ff00000) * (1.0f / 1048576.0f);
            const float ratio = exp2f((-36.54f - l2) * (1.0f / (float)M));
            hc = hc * (double)fminf(fmaxf(0.9f * ratio, 0.999f), 1.001f);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = pcur + hc;
}

template <int V>
void run(const char* name, double* out, long long* cyc, int nblocks, int nthreads) {
    const int cells = 2000;
    long long h;
    cell_loop<16, 20, V><<<nblocks, nthreads>>>(out, cyc, cells, 0.05, 1.3, 1.8, 0.4, 0.02);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s blocks=%d threads=%d: %.0f cycles per cell, %.1f per order\n", name, nblocks, nthreads, (double)h / cells, (double)h / cells / 16);
}

int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
    for (int rep = 0; rep < 2; ++rep) {
        run<0>("V0 product (STS/LDS.128 broadcast)", out, cyc, 1, 32);
        run<1>("V1 shuffle matvec", out, cyc, 1, 32);
        run<2>("V2 no matvec", out, cyc, 1, 32);
        run<3>("V3 matvec only (no conv)", out, cyc, 1, 32);
        run<4>("V4 product + controller", out, cyc, 1, 32);
        run<5>("V5 high-word controller", out, cyc, 1, 32);
        run<6>("V6 V5 + off-path combine + tree psum", out, cyc, 1, 32);
        run<0>("V0 two warps same CTA", out, cyc, 1, 64);
        run<0>("V0 256 CTAs", out, cyc, 256, 32);
        run<0>("V0 warps 0 and 4 (one sub-partition)", out, cyc, 1, 160);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
