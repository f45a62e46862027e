// tools/smsp_probe.cu — which sub-partition (SMSP) of an SM does a warp run on?
// (1) records %smid / %warpid of every warp of a 6-CTA-per-SM, 128-thread launch (the ensemble kernels' shape);
// (2) times FP64 work placed on chosen warps of a CTA: two busy warps on the SAME sub-partition take twice as long as
//     two on different ones (the FP64 pipe is per sub-partition), which tells whether SMSP = warp-in-CTA % 4,
//     %warpid % 4, or neither.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/smsp_probe tools/smsp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void __launch_bounds__(128, 6) where_kernel(unsigned* smid, unsigned* warpid, double* sink, int spin) {
    unsigned s, w;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
    double a = threadIdx.x * 1e-9, b = 1.0000001;
    for (int i = 0; i < spin; ++i) a = fma(a, b, 1e-9);  // keep every CTA resident while the others start
    if ((threadIdx.x & 31) == 0) {
        smid[blockIdx.x * 4 + (threadIdx.x >> 5)] = s;
        warpid[blockIdx.x * 4 + (threadIdx.x >> 5)] = w;
    }
    if (a == 123.0) *sink = a;
}

// one CTA per SM, `nwarps` warps; warp w works iff bit w of mask is set
__global__ void busy_kernel(unsigned mask, int iters, double* sink, unsigned* warpid_out) {
    const int w = threadIdx.x >> 5;
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
        unsigned hw;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
        warpid_out[w] = hw;
    }
    if (!((mask >> w) & 1u)) return;
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.0) *sink = s;
}

static float time_busy(unsigned mask, int nwarps, int sms, double* sink, unsigned* wid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    busy_kernel<<<sms, nwarps * 32>>>(mask, 1000, sink, wid);
    cudaEventRecord(e0);
    busy_kernel<<<sms, nwarps * 32>>>(mask, 200000, sink, wid);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, grid = 6 * sms;
    unsigned *smid, *warpid, *wid;
    double* sink;
    cudaMalloc(&smid, grid * 4 * sizeof(unsigned)); cudaMalloc(&warpid, grid * 4 * sizeof(unsigned));
    cudaMalloc(&wid, 64 * sizeof(unsigned)); cudaMalloc(&sink, 8);
    where_kernel<<<grid, 128>>>(smid, warpid, sink, 2000000);
    std::vector<unsigned> hs(grid * 4), hw(grid * 4);
    cudaMemcpy(hs.data(), smid, hs.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hw.data(), warpid, hw.size() * 4, cudaMemcpyDeviceToHost);
    printf("SMs %d, grid %d\n", sms, grid);
    int same_mod4 = 0, consecutive = 0;
    for (int b = 0; b < grid; ++b) {
        bool m4 = true, cons = true;
        for (int w = 0; w < 4; ++w) {
            if ((hw[b * 4 + w] & 3u) != (unsigned)w) m4 = false;
            if (hw[b * 4 + w] != hw[b * 4] + w) cons = false;
        }
        same_mod4 += m4; consecutive += cons;
    }
    printf("CTAs whose warp w has %%warpid %% 4 == w: %d of %d; with 4 consecutive %%warpid: %d\n", same_mod4, grid, consecutive);
    for (int b : {0, 1, 2, 147, 148, 149, 296, 444, 592, 740, 887}) {
        if (b >= grid) continue;
        printf("  block %4d: smid %3u  warpid %2u %2u %2u %2u   block/%d %% 4 = %d\n", b, hs[b * 4], hw[b * 4], hw[b * 4 + 1],
               hw[b * 4 + 2], hw[b * 4 + 3], sms, (b / sms) & 3);
    }
    // how many CTAs per SM, and which (warpid >> 2) slots they hold
    std::vector<int> per(sms, 0);
    for (int b = 0; b < grid; ++b) per[hs[b * 4]]++;
    int mn = 1 << 30, mx = 0;
    for (int s = 0; s < sms; ++s) { mn = per[s] < mn ? per[s] : mn; mx = per[s] > mx ? per[s] : mx; }
    printf("CTAs per SM: min %d max %d\n", mn, mx);
    printf("blocks on SM of block 0 (smid %u):", hs[0]);
    for (int b = 0; b < grid; ++b) if (hs[b * 4] == hs[0]) printf(" %d(w%u)", b, hw[b * 4]);
    printf("\n");

    // timing: 8 warps per CTA, one CTA per SM
    unsigned hwid[8];
    struct { const char* what; unsigned mask; } cases[] = {
        {"warp 0 only", 0x01}, {"warps 0,1", 0x03}, {"warps 0,4", 0x11}, {"warps 0,2", 0x05}, {"warps 0,1,2,3", 0x0f},
        {"warps 0,4,1,5", 0x33}, {"all 8", 0xff}};
    for (auto& c : cases) {
        const float ms = time_busy(c.mask, 8, sms, sink, wid);
        cudaMemcpy(hwid, wid, sizeof(hwid), cudaMemcpyDeviceToHost);
        printf("%-16s %8.3f ms   (%%warpid of warps 0..7 in block 0: %u %u %u %u %u %u %u %u)\n", c.what, ms, hwid[0], hwid[1],
               hwid[2], hwid[3], hwid[4], hwid[5], hwid[6], hwid[7]);
    }
    return 0;
}
