// tools/lds_pattern_probe.cu — what does a shared-memory load cost the L1/shared data pipe when the lanes of a warp
// read 1, 2 or 4 distinct addresses (broadcast / multicast)?  rk_warp_linear32_kernel is bound by that pipe
// (ncu: l1tex__data_pipe_lsu_wavefronts 91 % busy, profiles/r04d_cfg4_source.md): a full broadcast LDS.128 takes two
// wavefronts.  Every variant issues the same number of loads from 16 resident warps per SM (4 CTAs of 128) and reports
// SM cycles per warp-level load, i.e. wavefronts if the pipe takes one per cycle.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/lds_pattern_probe.cu -o tools/lds_pattern_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int WIDTH, int MODE>  // WIDTH: bytes per lane (4, 8, 16); MODE: how the lanes' addresses differ
__global__ void __launch_bounds__(128, 4) probe(double* out, int iters, long long* cycles) {
    __shared__ __align__(16) double s[4][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < 64; i += 32) s[warp][i] = 1.0 + i;
    __syncwarp();
    int off;  // in doubles
    if (MODE == 0) off = 0;                        // one address: full broadcast
    else if (MODE == 1) off = (lane >> 4) * 16;    // two addresses, split by half-warp
    else if (MODE == 2) off = (lane & 1) * 16;     // two addresses, interleaved by lane parity
    else if (MODE == 3) off = (lane >> 3) * 8;     // four addresses, split by quarter-warp
    else if (MODE == 4) off = (lane & 3) * 8;      // four addresses, interleaved
    else off = lane * (WIDTH / 8 > 0 ? WIDTH / 8 : 1) % 32;  // every lane its own (no broadcast): the plain cost
    const unsigned base = (unsigned)__cvta_generic_to_shared(&s[warp][off]);
    double acc0 = 0, acc1 = 0;
    float f0 = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if constexpr (WIDTH == 16) {
                double a, b;
                asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(base + 16u * (u & 3)));
                if (u == 7) { acc0 += a; acc1 += b; }
            } else if constexpr (WIDTH == 8) {
                double a;
                asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(a) : "r"(base + 8u * (u & 7)));
                if (u == 7) acc0 += a;
            } else {
                float a;
                asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(a) : "r"(base + 4u * (u & 7)));
                if (u == 7) f0 += a;
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    out[(size_t)blockIdx.x * 128 + threadIdx.x] = acc0 + acc1 + f0;
}

template <int WIDTH, int MODE> void run(const char* what, double* out, long long* dcyc, int sms) {
    const int iters = 4096;
    probe<WIDTH, MODE><<<sms * 4, 128>>>(out, 64, dcyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<WIDTH, MODE><<<sms * 4, 128>>>(out, iters, dcyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double loads_per_sm = 16.0 * iters * 8;  // warp-level loads per SM (16 resident warps)
    const double cyc = ms * 1e-3 * khz * 1e3 / loads_per_sm;  // (event time at the boost clock; clock64 of one warp under-reports)
    printf("%-40s %2d B/lane: %5.2f SM cycles per warp-level load (%4.2f per 8 B per lane), %.3f ms\n", what, WIDTH, cyc, cyc * 8.0 / WIDTH, ms);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; long long* dcyc;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 4 * 128 * 8); cudaMalloc(&dcyc, 8);
    printf("%s, %d SMs (volatile loads; only one in eight feeds an add)\n", p.name, p.multiProcessorCount);
#define ALL(W) \
    run<W, 0>("one address (full broadcast)", out, dcyc, p.multiProcessorCount); \
    run<W, 1>("two addresses, by half-warp", out, dcyc, p.multiProcessorCount); \
    run<W, 2>("two addresses, by lane parity", out, dcyc, p.multiProcessorCount); \
    run<W, 3>("four addresses, by quarter-warp", out, dcyc, p.multiProcessorCount); \
    run<W, 4>("four addresses, interleaved", out, dcyc, p.multiProcessorCount); \
    run<W, 5>("every lane its own address", out, dcyc, p.multiProcessorCount);
    ALL(16) ALL(8) ALL(4)
    return 0;
}
