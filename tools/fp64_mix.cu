// tools/fp64_mix.cu — does non-FP64 work issue "for free" next to a saturated FP64 pipe on B200?
// Per loop iteration: 16 full-rate DFMAs (two register sources) + M independent integer/FP32 instructions.
// Reports the DFMA rate as a fraction of 64 DFMA/clk/SM for M = 0..32 at 6 and 8 warps per SMSP.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_mix.cu -o fp64_mix && ./fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int M, int KIND>
__global__ void k(double* sink, int iters, double a, unsigned s0) {
    double x[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) x[c] = 1.0 + 1e-3 * (threadIdx.x + c);
    unsigned v[8];
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { v[c] = s0 + threadIdx.x * (c + 1); f[c] = 1.0f + c; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            x[c] = fma(x[c], a, x[c]);
#pragma unroll
            for (int m = 0; m < M / 16 + ((c < M % 16) ? 1 : 0); ++m) {
                const int j = (c + m) & 7;
                if (KIND == 0) v[j] = (v[j] ^ (v[j] >> 3)) + s0;   // LOP3/SHF/IADD on the integer pipe (counted as 2-3 instr)
                else if (KIND == 1) f[j] = fmaf(f[j], 0.999f, 1e-3f);  // FFMA
                else asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));  // MUFU
            }
        }
    }
    double s = 0;
    unsigned w = 0;
    float g = 0;
#pragma unroll
    for (int c = 0; c < 16; ++c) s += x[c];
#pragma unroll
    for (int c = 0; c < 8; ++c) { w ^= v[c]; g += f[c]; }
    if (s == 123.456 || w == 0x12345u || g == 77.125f) sink[0] = s + w + g;
}

template <int M, int KIND> void run(int block, int bps, int sm, double* sink) {
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sm * bps;
    k<M, KIND><<<grid, block>>>(sink, 16, 0.999999, 3u);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<M, KIND><<<grid, block>>>(sink, iters, 0.999999, 3u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double dfma = 16.0 * iters * (double)grid * block;
    const double peak = 64.0 * sm * 1.965e9;
    printf("kind %d  M %2d per 16 DFMA  warps/SMSP %4.1f : DFMA rate %5.1f %% of 64/clk/SM (%.3f ms)\n", KIND, M,
           block / 32.0 * bps / 4.0, 100.0 * dfma / (best * 1e-3) / peak, best);
}

int main() {
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    double* sink;
    cudaMalloc(&sink, 8);
    for (int bps : {6, 8}) {
        run<0, 0>(128, bps, sm, sink);
        run<4, 0>(128, bps, sm, sink);
        run<8, 0>(128, bps, sm, sink);
        run<16, 0>(128, bps, sm, sink);
        run<32, 0>(128, bps, sm, sink);
        run<4, 1>(128, bps, sm, sink);
        run<8, 1>(128, bps, sm, sink);
        run<16, 1>(128, bps, sm, sink);
        run<32, 1>(128, bps, sm, sink);
        run<2, 2>(128, bps, sm, sink);
        run<4, 2>(128, bps, sm, sink);
    }
    return 0;
}
