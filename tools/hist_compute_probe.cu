// tools/hist_compute_probe.cu — a synthetic twin of the dense-output kernel: every lane runs ~128 FP64 instructions
// per "step" (like one RKF45 attempt of Lorenz) and then records 32 bytes into its own stream.  Compares: no store,
// one 256-bit store per step (K = 1), four records kept in registers and written as one 128-byte line every 4th step
// (K = 4), and the same with the line written by 4 cooperating lanes (each store instruction covers 8 whole lines).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/hist_compute_probe tools/hist_compute_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// SCATTER: the lanes of a warp write streams that lie far apart (a different 2 MB page each), as in the real kernel once
// lanes have refilled with whatever trajectory index came next; otherwise lane l of a warp writes stream 32 w + l.
template <int MODE, bool SCATTER = false>  // 0 none, 1 K=1, 2 K=4 per lane, 3 K=4 by 4 cooperating lanes
__global__ void __launch_bounds__(128, 6) twin(double* base, int steps, size_t stride, double* sink) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, total = (size_t)gridDim.x * blockDim.x;
    double* mine = base + (SCATTER ? (t * 48271u) % total : t) * stride;
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = 1e-3 * (double)(t + i);
    const double b = 0.999999, c = 1e-9;
    double r0[4], r1[4], r2[4];
    const unsigned lane = threadIdx.x & 31;
    for (int s = 0; s < steps; ++s) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
        if (MODE == 1) {
            st256(mine + (size_t)s * 4, a[0], a[1], a[2], a[3]);
        } else if (MODE == 2 || MODE == 3) {
            const int ph = s & 3;
            if (ph == 0) { r0[0] = a[0]; r0[1] = a[1]; r0[2] = a[2]; r0[3] = a[3]; }
            else if (ph == 1) { r1[0] = a[0]; r1[1] = a[1]; r1[2] = a[2]; r1[3] = a[3]; }
            else if (ph == 2) { r2[0] = a[0]; r2[1] = a[1]; r2[2] = a[2]; r2[3] = a[3]; }
            else {
                double* dst = mine + (size_t)(s - 3) * 4;
                if (MODE == 2) {
                    st256(dst, r0[0], r0[1], r0[2], r0[3]);
                    st256(dst + 4, r1[0], r1[1], r1[2], r1[3]);
                    st256(dst + 8, r2[0], r2[1], r2[2], r2[3]);
                    st256(dst + 12, a[0], a[1], a[2], a[3]);
                } else {
                    // lane l writes record (l & 3) of the lines of lanes (l & ~3) + j, j = 0..3
                    const unsigned q = lane & 3;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned owner = (lane & ~3u) + j;
                        double v[4];
#pragma unroll
                        for (int d = 0; d < 4; ++d) {
                            const double x0 = __shfl_sync(0xffffffffu, r0[d], owner), x1 = __shfl_sync(0xffffffffu, r1[d], owner);
                            const double x2 = __shfl_sync(0xffffffffu, r2[d], owner), x3 = __shfl_sync(0xffffffffu, a[d], owner);
                            v[d] = q == 0 ? x0 : q == 1 ? x1 : q == 2 ? x2 : x3;
                        }
                        const unsigned long long p = __shfl_sync(0xffffffffu, (unsigned long long)dst, owner);
                        st256((double*)p + q * 4, v[0], v[1], v[2], v[3]);
                    }
                }
            }
        }
    }
    double sum = 0;
    for (int i = 0; i < 8; ++i) sum += a[i];
    if (sum == 123.456) *sink = sum;
}

int main(int argc, char** argv) {
    const int grid = 888, block = 128, steps = argc > 1 ? atoi(argv[1]) : 2048;
    const size_t stride = (size_t)steps * 4, total = (size_t)grid * block * stride;
    double *buf, *sink;
    if (cudaMalloc(&buf, total * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double gb = total * 8.0 / 1e9;
    printf("%d lanes x %d steps, 128 DFMA + one 32-byte record per step, %.1f GB\n", grid * block, steps, gb);
    auto run = [&](auto k, const char* what) {
        k<<<grid, block>>>(buf, steps, stride, sink);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k<<<grid, block>>>(buf, steps, stride, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-34s %8.2f ms  %7.1f GB/s  %6.2f TFLOP/s\n", what, ms, gb / ms * 1e3, 2.0 * 128 * grid * block * steps / ms / 1e9);
    };
    run(twin<0>, "no store");
    run(twin<1>, "K = 1 (a sector per step)");
    run(twin<2>, "K = 4 per lane (a line per 4 steps)");
    run(twin<3>, "K = 4, 4 lanes per line");
    run(twin<1, true>, "K = 1, lanes on scattered pages");
    run(twin<2, true>, "K = 4, lanes on scattered pages");
    if (cudaGetLastError() != cudaSuccess) printf("CUDA error\n");
    return 0;
}
