// tools/hist_write_probe.cu — what write bandwidth does the dense-output access pattern reach on B200, and how does it
// depend on how many 32-byte records a lane writes back to back into one trajectory's history before moving on?
// Every thread owns S streams (trajectories) of L records; it writes K records to stream 0, K to stream 1, ... and
// comes back to stream 0 S*K stores later — by then ~S*K*113664*32 bytes of other traffic went through L2, like the
// ~1 us between two accepted steps of a lane in the real kernel.  K = 1 is the kernel's present pattern.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/hist_write_probe tools/hist_write_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int WIDTH>  // bytes per store instruction: 32 (st.v4.f64), 16, 8
__global__ void __launch_bounds__(128) write_kernel(double* base, int S, int L, int K, size_t stream_stride) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double* mine = base + t * (size_t)S * stream_stride;
    const double v = (double)t;
    for (int r0 = 0; r0 < L; r0 += K)
        for (int s = 0; s < S; ++s) {
            double* p = mine + (size_t)s * stream_stride + (size_t)r0 * 4;
            for (int k = 0; k < K; ++k, p += 4) {
                if (WIDTH == 32)
                    asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p), "d"(v) : "memory");
                else if (WIDTH == 16) {
                    asm volatile("st.global.v2.f64 [%0], {%1, %1};" ::"l"(p), "d"(v) : "memory");
                    asm volatile("st.global.v2.f64 [%0], {%1, %1};" ::"l"(p + 2), "d"(v) : "memory");
                } else {
                    for (int j = 0; j < 4; ++j) asm volatile("st.global.f64 [%0], %1;" ::"l"(p + j), "d"(v) : "memory");
                }
            }
        }
}

// the same bytes written fully coalesced (a warp writes 1 KB contiguous per instruction): the ceiling
__global__ void __launch_bounds__(128) stream_kernel(double* base, size_t n4) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, T = (size_t)gridDim.x * blockDim.x;
    const double v = (double)t;
    for (size_t i = t; i < n4; i += T) asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(base + i * 4), "d"(v) : "memory");
}

int main(int argc, char** argv) {
    const int grid = 888, block = 128;
    const int S = argc > 1 ? atoi(argv[1]) : 8, L = argc > 2 ? atoi(argv[2]) : 256;
    const size_t stream_stride = (size_t)L * 4;  // doubles: streams are back to back, L*32 bytes each
    const size_t total = (size_t)grid * block * S * stream_stride;
    double* buf;
    if (cudaMalloc(&buf, total * sizeof(double)) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto launch) {
        launch();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        return ms;
    };
    const double gb = total * 8.0 / 1e9;
    printf("%.1f GB per pass, %d lanes, %d streams per lane of %d records\n", gb, grid * block, S, L);
    float ms = time([&] { stream_kernel<<<grid, block>>>(buf, total / 4); });
    printf("coalesced stream           : %7.2f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    for (int K : {1, 2, 4, 8, 16, 64, 256}) {
        ms = time([&] { write_kernel<32><<<grid, block>>>(buf, S, L, K, stream_stride); });
        printf("K = %3d records, 256-bit st: %7.2f ms  %7.1f GB/s\n", K, ms, gb / ms * 1e3);
    }
    for (int K : {1, 8}) {
        ms = time([&] { write_kernel<16><<<grid, block>>>(buf, S, L, K, stream_stride); });
        printf("K = %3d records, 128-bit st: %7.2f ms  %7.1f GB/s\n", K, ms, gb / ms * 1e3);
        ms = time([&] { write_kernel<8><<<grid, block>>>(buf, S, L, K, stream_stride); });
        printf("K = %3d records,  64-bit st: %7.2f ms  %7.1f GB/s\n", K, ms, gb / ms * 1e3);
    }
    if (cudaGetLastError() != cudaSuccess) printf("CUDA error\n");
    return 0;
}
