// tools/latency_probe.cu -- dependent-chain latencies of the FP64 ops the cycle kernel is made of (B200, sm_100a).
// build: nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -o /tmp/latency_probe tools/latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void probe(double* out, long long* cyc, double seed) {
    __shared__ double sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = seed + i * 1e-3;
    __syncthreads();
    double a = seed, b = seed * 0.5;
    long long t0, t1;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a += b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fma(a, 0.999999, b);
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a * 1.0000001;
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // sqrt chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) a = sqrt(a + 2.0);
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // div chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) a = b / (a + 1.5);
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // LDS dependent chain (pointer chase through indices)
    int idx = threadIdx.x & 7;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = ((int)sh[idx]) & 1023;
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // seq-sum style: LDS (independent addresses) + DADD chain
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) a += sh[i];
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // compare-select chain (argmin update): DSETP + 2 SEL
    double bd = 1e300; int bj = 0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { double d = sh[i] * b; if (d < bd) { bd = d; bj = i; } }
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // shuffle of a double (2 SHFL) dependent
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0;
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
    out[threadIdx.x] = a + idx + bd + bj;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
    const char* names[9] = {"DADD", "DFMA", "DMUL", "sqrt(x+2)", "b/(a+1.5)", "LDS chase(+cvt)", "LDS+DADD sum", "argmin update", "shfl.f64+DADD"};
    for (int warps = 1; warps <= 16; warps *= 4) {
        probe<<<1, 32 * warps>>>(out, cyc, 1.5);
        probe<<<1, 32 * warps>>>(out, cyc, 1.5);
        cudaDeviceSynchronize();
        long long h[16];
        cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
        printf("warps/SM=%d:", warps);
        for (int i = 0; i < 9; ++i) printf("  %s %.1f", names[i], (double)h[i] / N);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
