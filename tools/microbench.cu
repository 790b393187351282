// Micro-benchmarks used to size the trajectory kernel's critical path on B200 (not part of the library).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_chain(double* out, long long* cyc, int iters, double m, double c, int nchains) {
    double a0 = 1.0, a1 = 2.0, a2 = 3.0, a3 = 4.0;
    long long t0 = clock64();
    if (nchains == 1) for (int i = 0; i < iters; ++i) a0 = fma(a0, m, c);
    else if (nchains == 2) for (int i = 0; i < iters; ++i) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); }
    else for (int i = 0; i < iters; ++i) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c); }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
__global__ void dadd_chain(double* out, long long* cyc, int iters, double c) {
    double a0 = 1.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a0 = a0 + c;
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0;
}
__global__ void sts_lds_roundtrip(double* out, long long* cyc, int iters) {
    __shared__ double buf[64];
    double a = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        buf[threadIdx.x] = a;
        __syncwarp();
        a = buf[(threadIdx.x + 1) & 31] + 1.0;
        __syncwarp();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void shfl_roundtrip(double* out, long long* cyc, int iters) {
    double a = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31) + 1.0;
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    long long h; const int it = 100000;
    for (int nwarps : {1, 2, 4, 8}) for (int nch : {1, 2, 4}) {
        dfma_chain<<<1, 32 * nwarps>>>(out, cyc, it, 1.0000001, 1e-9, nch); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA warps/SM=%d chains=%d : %.2f cycles per loop iter (%.2f per DFMA issued by the warp)\n", nwarps, nch, (double)h / it, (double)h / it / nch);
    }
    dadd_chain<<<1, 32>>>(out, cyc, it, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DADD dependent chain: %.2f cycles\n", (double)h / it);
    sts_lds_roundtrip<<<1, 32>>>(out, cyc, it); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("STS -> syncwarp -> LDS -> DADD -> syncwarp round trip: %.2f cycles\n", (double)h / it);
    shfl_roundtrip<<<1, 32>>>(out, cyc, it); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("SHFL(double) -> DADD round trip: %.2f cycles\n", (double)h / it);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
