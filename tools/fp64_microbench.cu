// FP64 pipe microbenchmarks for B200 (sm_100a): what does a DFMA really cost?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_microbench tools/fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
// independent chains, multiplier/addend are kernel-constant (constant-bank / uniform operands)
__global__ void k_const(int iters, double m, double c, double *sink) {
    double a[CHAINS];
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], m, c);
    double s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
}
// independent chains, all three sources are distinct per-thread registers
__global__ void k_3reg(int iters, const double *in, double *sink) {
    double a[CHAINS], b[CHAINS], c[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; c[i] = in[threadIdx.x + 16 + i]; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], b[i], c[i]);
    double s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
}
// two distinct register sources: a = fma(a, b, a)
__global__ void k_2reg(int iters, const double *in, double *sink) {
    double a[CHAINS], b[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], b[i], a[i]);
    double s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
}
// DMUL / DADD with two distinct register sources
__global__ void k_mul(int iters, const double *in, double *sink) {
    double a[CHAINS], b[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) a[i] = a[i] * b[i];
    double s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
}
__global__ void k_add(int iters, const double *in, double *sink) {
    double a[CHAINS], b[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) a[i] = a[i] + b[i];
    double s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
}
// dependent chain latency: one warp, cycles per dependent DFMA
__global__ void k_latency(int iters, const double *in, double *sink, long long *cyc) {
    double a = in[threadIdx.x], b = in[threadIdx.x + 8], c = in[threadIdx.x + 16];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c);
        a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (a == 12345.678) sink[0] = a;
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double *in, *sink; long long *cyc;
    cudaMalloc(&in, 4096 * 8); cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    double h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 20000;
    printf("SMs %d, clock %d kHz\n", sms, clk);
    for (int wps = 1; wps <= 16; wps *= 2) {   // warps per SMSP
        int block = 128 * (wps >= 4 ? 4 : wps), grid = sms * (wps >= 4 ? wps / 4 : 1);
        if (wps < 4) { block = 128 * wps; grid = sms; }
        double ops = (double)CHAINS * iters * block * grid;
        float t;
        t = timeit([&] { k_const<<<grid, block>>>(iters, 1.0000001, 1e-7, sink); }); double r0 = ops / t / 1e9;
        t = timeit([&] { k_3reg<<<grid, block>>>(iters, in, sink); }); double r1 = ops / t / 1e9;
        t = timeit([&] { k_2reg<<<grid, block>>>(iters, in, sink); }); double r2 = ops / t / 1e9;
        t = timeit([&] { k_mul<<<grid, block>>>(iters, in, sink); }); double r3 = ops / t / 1e9;
        t = timeit([&] { k_add<<<grid, block>>>(iters, in, sink); }); double r4 = ops / t / 1e9;
        printf("warps/SMSP %2d (block %d grid %d): Tinst/s  const %.2f  3reg %.2f  2reg %.2f  dmul %.2f  dadd %.2f\n", wps, block, grid, r0, r1, r2, r3, r4);
    }
    k_latency<<<1, 32>>>(1000, in, sink, cyc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles\n", (double)c / 8000.0);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
