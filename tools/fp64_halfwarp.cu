// Does the B200 FP64 pipe charge a warp-instruction by ACTIVE lanes?  (16 FP64 lanes per SM sub-partition: a full
// warp takes two passes.)  Same independent-DFMA kernel with 32 / 16 / 8 / 1 active lanes per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_halfwarp tools/fp64_halfwarp.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
__global__ void k_2reg(int iters, int active, int lane_lo, const double *in, double *sink) {
    const int lane = threadIdx.x & 31;
    if (lane < lane_lo || lane >= lane_lo + active) return;
    double a[CHAINS], b[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], b[i], a[i]);
    double s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *in, *sink; cudaMalloc(&in, 4096 * 8); cudaMalloc(&sink, 8);
    double h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps = 1; wps <= 8; wps *= 2) {
        int block = wps >= 4 ? 512 : 128 * wps, grid = sms * (wps >= 4 ? wps / 4 : 1);
        const int cfg[][2] = {{32, 0}, {16, 0}, {16, 16}, {16, 8}, {8, 0}, {1, 0}};
        printf("warps/SMSP %d:", wps);
        for (auto &c : cfg) {
            k_2reg<<<grid, block>>>(iters, c[0], c[1], in, sink); cudaDeviceSynchronize();
            float best = 1e30f;
            for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k_2reg<<<grid, block>>>(iters, c[0], c[1], in, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
            printf("  lanes[%d..%d) %.3f ms", c[1], c[1] + c[0], best);
        }
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
