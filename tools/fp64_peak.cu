// tools/fp64_peak.cu — explores the achievable DFMA issue rate of one B200 (the denominator of the RK kernels'
// roofline).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_peak.cu -o fp64_peak && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int MODE>
__global__ void k(double* sink, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = 1.0 + 1e-3 * (threadIdx.x + c);
    double yv[CHAINS], zv[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { yv[c] = a + 1e-9 * (threadIdx.x + c); zv[c] = b * (1 + c) + 1e-12 * threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (MODE == 0) x[c] = fma(x[c], a, b);          // 3 register sources (a, b shared)
                else if (MODE == 1) x[c] = fma(x[c], a, x[c]);  // 2 distinct register sources
                else if (MODE == 2) x[c] = fma(x[c], 0.999999, b);  // one immediate/constant source
                else if (MODE == 3) x[c] = fma(x[c], yv[c], zv[c]);  // 3 distinct register sources, nothing shared
                else if (MODE == 4) x[c] = fma(yv[c], zv[c], x[c]);  // same, accumulator form
                else if (MODE == 5) x[c] = x[c] * yv[c];              // DMUL, 2 distinct registers
                else if (MODE == 6) x[c] = x[c] + yv[c];              // DADD
                else x[c] = fma(x[c], a, zv[c]);                      // 3 register sources, one shared
            }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 123.456) sink[0] = s;
}

template <int CHAINS, int MODE> void run(int block, int blocks_per_sm, int sm, double* sink) {
    const int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sm * blocks_per_sm;
    k<CHAINS, MODE><<<grid, block>>>(sink, 16, 0.999999, 1e-9);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<CHAINS, MODE><<<grid, block>>>(sink, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * CHAINS * 8.0 * iters * (double)grid * block;
    printf("chains %2d mode %d block %4d blocks/SM %d warps/SMSP %4.1f : %7.3f TFLOP/s (%.3f ms)\n", CHAINS, MODE, block,
           blocks_per_sm, block / 32.0 * blocks_per_sm / 4.0, flops / (best * 1e-3) / 1e12, best);
}

int main() {
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    double* sink;
    cudaMalloc(&sink, 8);
    printf("SMs %d; nominal 64 DFMA/clk/SM * 2 * 1.965 GHz * %d = %.2f TFLOP/s\n", sm, sm, 64 * 2 * 1.965e9 * sm / 1e12);
    run<8, 0>(256, 4, sm, sink);
    run<16, 0>(256, 4, sm, sink);
    run<16, 0>(128, 4, sm, sink);
    run<16, 0>(512, 2, sm, sink);
    run<16, 0>(1024, 1, sm, sink);
    run<16, 0>(256, 2, sm, sink);
    run<16, 0>(256, 1, sm, sink);
    run<32, 0>(256, 2, sm, sink);
    run<16, 1>(256, 4, sm, sink);
    run<16, 2>(256, 4, sm, sink);
    run<8, 1>(512, 2, sm, sink);
    run<4, 1>(1024, 2, sm, sink);
    run<8, 3>(256, 4, sm, sink);
    run<8, 4>(256, 4, sm, sink);
    run<8, 5>(256, 4, sm, sink);
    run<8, 6>(256, 4, sm, sink);
    run<8, 7>(256, 4, sm, sink);
    run<8, 3>(128, 6, sm, sink);
    return 0;
}
