// tools/matvec_probe.cu — how fast can one warp apply its trajectory's 32 x 32 matrix (config 4, rk_warp_linear.cuh)?
// The matrix lives in registers (32 doubles per lane); the layouts differ in WHICH 32 entries a lane owns:
//   ROWS x COLS block per lane, (32/ROWS) x (32/COLS) lanes: lane (r, c) owns rows ROWS*r.., columns COLS*c..
//   1 x 32  = the round-1 kernel: a lane owns one row, every DFMA has three unrelated register sources
//   R x C   = R accumulators share each Y value (operand reuse), partial sums reduce-scattered over the C-lanes
// Each kernel iterates y <- y + h * A y (dependent matvecs, like the RK stages) and reports DFMA-rate.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/matvec_probe.cu -o tools/matvec_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 32;

// GROUPED: the LC lanes that share a row are 32/LC lanes apart (column group c = lane / LR: half-warps for 2 x 16,
// quarter-warps for 4 x 8), so that each LDS.128 of the Y values reads one address per half/quarter warp — two
// wavefronts like a full broadcast (tools/lds_pattern_probe.cu); the round-2 layouts interleaved them (c = lane % LC),
// which costs 4 and 8 wavefronts and hid what the blockings save.
template <int ROWS, int COLS, int CHAINS = 4, int MINB = 4, bool GROUPED = false>
__global__ void __launch_bounds__(128, MINB) probe(const double* __restrict__ Aglob, double* out, int iters, double h) {
    constexpr int LR = N / ROWS;   // lanes along rows
    constexpr int LC = N / COLS;   // lanes along columns (these lanes reduce)
    static_assert(LR * LC == 32, "one warp");
    __shared__ __align__(16) double s_y[4][N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = GROUPED ? lane % LR : lane / LC, c = GROUPED ? lane / LR : lane % LC;
    constexpr int XS = GROUPED ? LR : 1;  // lane distance of neighbouring column groups
    double* sy = s_y[warp];
    const size_t traj = (size_t)blockIdx.x * 4 + warp;
    double A[ROWS][COLS];
#pragma unroll
    for (int a = 0; a < ROWS; ++a)
#pragma unroll
        for (int b = 0; b < COLS; ++b) A[a][b] = Aglob[(traj * N + (GROUPED ? r + LR * a : ROWS * r + a)) * N + COLS * c + b];
    // the lane's own component of y: component lane (1 x 32), or ROWS * r + (c % ROWS) ... kept simple: component `lane`
    // is owned by lane `lane` in every layout (lane (r, c) -> row index ROWS * r + c when LC == ROWS)
    double y = 1.0 + 1e-3 * lane;
    for (int it = 0; it < iters; ++it) {
        __syncwarp();
        sy[lane] = y;
        __syncwarp();
        double acc[ROWS];
#pragma unroll
        for (int a = 0; a < ROWS; ++a) acc[a] = 0.0;
        const double2* v = reinterpret_cast<const double2*>(sy + COLS * c);
        if constexpr (ROWS == 1 && CHAINS == 8) {  // eight independent chains of four
            double c[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) c[q] = 0.0;
#pragma unroll
            for (int j8 = 0; j8 < COLS / 8; ++j8) {
#pragma unroll
                for (int q2 = 0; q2 < 4; ++q2) {
                    const double2 pp = v[4 * j8 + q2];
                    c[2 * q2] = fma(A[0][8 * j8 + 2 * q2], pp.x, c[2 * q2]);
                    c[2 * q2 + 1] = fma(A[0][8 * j8 + 2 * q2 + 1], pp.y, c[2 * q2 + 1]);
                }
            }
            acc[0] = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
        } else if constexpr (ROWS == 2 && CHAINS == 4) {  // two rows, two chains each
            double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int b2 = 0; b2 < COLS / 2; ++b2) {
                const double2 pp = v[b2];
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    c[a][0] = fma(A[a][2 * b2], pp.x, c[a][0]);
                    c[a][1] = fma(A[a][2 * b2 + 1], pp.y, c[a][1]);
                }
            }
            acc[0] = c[0][0] + c[0][1];
            acc[1] = c[1][0] + c[1][1];
        } else if constexpr (ROWS == 1) {
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
            for (int j4 = 0; j4 < COLS / 4; ++j4) {
                const double2 p0 = v[2 * j4], p1 = v[2 * j4 + 1];
                s0 = fma(A[0][4 * j4 + 0], p0.x, s0);
                s1 = fma(A[0][4 * j4 + 1], p0.y, s1);
                s2 = fma(A[0][4 * j4 + 2], p1.x, s2);
                s3 = fma(A[0][4 * j4 + 3], p1.y, s3);
            }
            acc[0] = (s0 + s1) + (s2 + s3);
        } else {
#pragma unroll
            for (int b2 = 0; b2 < COLS / 2; ++b2) {
                const double2 p = v[b2];
#pragma unroll
                for (int a = 0; a < ROWS; ++a) acc[a] = fma(A[a][2 * b2], p.x, acc[a]);
#pragma unroll
                for (int a = 0; a < ROWS; ++a) acc[a] = fma(A[a][2 * b2 + 1], p.y, acc[a]);
            }
        }
        // reduce-scatter over the LC lanes that share the rows: lane c ends with row ROWS * r + (its share)
        double dy;
        if constexpr (LC == 1) {
            dy = acc[0];
        } else {
            // butterfly: at each step a lane keeps half of its values and receives the partner's partials for them
            double val[ROWS];
#pragma unroll
            for (int a = 0; a < ROWS; ++a) val[a] = acc[a];
            int keep = ROWS;
#pragma unroll
            for (int m = LC / 2; m >= 1; m >>= 1) {
                const bool upper = (c & m) != 0;
                if (keep > 1) {
                    keep >>= 1;
#pragma unroll
                    for (int a = 0; a < ROWS / 2; ++a) {
                        if (a < keep) {
                            const double send = upper ? val[a] : val[a + keep];
                            const double mine = upper ? val[a + keep] : val[a];
                            val[a] = mine + __shfl_xor_sync(0xffffffffu, send, m * XS);
                        }
                    }
                } else {  // more lanes than rows: plain all-reduce of the one value
                    val[0] += __shfl_xor_sync(0xffffffffu, val[0], m * XS);
                }
            }
            dy = val[0];
        }
        y = fma(h, dy, y);
    }
    out[traj * N + lane] = y;
}

template <int ROWS, int COLS, int CHAINS = 4, int MINB = 4, bool GROUPED = false> void run(const double* A, double* out, int sm, int blocks_per_sm) {
    const int iters = 4096, grid = sm * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<ROWS, COLS, CHAINS, MINB, GROUPED><<<grid, 128>>>(A, out, 64, 1e-9);
    float best = 1e30f;
    for (int k = 0; k < 3; ++k) {
        cudaEventRecord(e0);
        probe<ROWS, COLS, CHAINS, MINB, GROUPED><<<grid, 128>>>(A, out, iters, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double flops = 2.0 * N * N * (double)iters * grid * 4;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, probe<ROWS, COLS, CHAINS, MINB, GROUPED>);
    printf("%s block %2d x %2d per lane, %d chains, %d CTAs/SM (%d warps/SMSP), %3d regs: %7.3f TFLOP/s = %5.1f %% of 37.2 (%.3f ms)\n", GROUPED ? "grouped    " : "interleaved", ROWS, COLS, CHAINS,
           blocks_per_sm, blocks_per_sm, fa.numRegs, flops / (best * 1e-3) / 1e12, 100 * flops / (best * 1e-3) / 37.2e12, best);
}

int main() {
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    double *A, *out;
    const size_t traj = (size_t)sm * 5 * 4;
    cudaMalloc(&A, traj * N * N * 8);
    cudaMalloc(&out, traj * N * 8);
    cudaMemset(A, 0, traj * N * N * 8);
    for (int bps : {4, 3}) {
        run<2, 16, 2, 4, true>(A, out, sm, bps);
        run<2, 16, 4, 4, true>(A, out, sm, bps);
        run<4, 8, 4, 4, true>(A, out, sm, bps);
        run<8, 4, 4, 4, true>(A, out, sm, bps);
    }
    run<2, 16, 4, 5, true>(A, out, sm, 5);
    run<2, 16, 2, 5, true>(A, out, sm, 5);
    run<4, 8, 4, 5, true>(A, out, sm, 5);
    for (int bps : {4}) {
        run<1, 32>(A, out, sm, bps);
        run<1, 32, 8>(A, out, sm, bps);
        run<2, 16, 2>(A, out, sm, bps);
        run<2, 16, 4>(A, out, sm, bps);
        run<4, 8>(A, out, sm, bps);
        run<8, 4>(A, out, sm, bps);
    }
    run<2, 16, 4, 5>(A, out, sm, 5);
    run<2, 16, 2, 5>(A, out, sm, 5);
    run<4, 8, 4, 5>(A, out, sm, 5);
    return 0;
}
