// Micro-benchmark: do FP64 arithmetic and shared-memory traffic overlap on a B200 SM?
// Half of the warps of every CTA run a DFMA/DADD loop, the other half a 128-bit LDS loop (different warps, no
// dependencies between them).  Times: FP64 warps alone, LDS warps alone, both together.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) k(double* out, int iters_f, int iters_s, int mode) {
    __shared__ double2 sm[2048];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 2048; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    const bool do_f = (mode & 1) && (warp & 1) == 0;     // even warps: FP64
    const bool do_s = (mode & 2) && (warp & 1) == 1;     // odd warps: shared-memory loads
    double a0 = tid, a1 = tid + 1, a2 = tid + 2, a3 = tid + 3, a4 = tid + 4, a5 = tid + 5, a6 = tid + 6, a7 = tid + 7;
    if (do_f) {
        const double m = 0.999999, c = 1e-9;
        for (int i = 0; i < iters_f; ++i) {
            a0 = a0 * m + c; a1 = a1 + a0; a2 = a2 * m + c; a3 = a3 + a2;
            a4 = a4 * m + c; a5 = a5 + a4; a6 = a6 * m + c; a7 = a7 + a6;
        }
    }
    double2 s0 = make_double2(0, 0), s1 = s0, s2 = s0, s3 = s0;
    if (do_s) {
        int idx = tid & 511;
        for (int i = 0; i < iters_s; ++i) {
            double2 v0 = sm[idx], v1 = sm[idx + 512], v2 = sm[idx + 1024], v3 = sm[idx + 1536];
            s0.x += v0.y == 1e300 ? 1.0 : 0.0; s1.x += v1.y == 1e300 ? 1.0 : 0.0;
            s2.x += v2.y == 1e300 ? 1.0 : 0.0; s3.x += v3.y == 1e300 ? 1.0 : 0.0;
            idx = (idx + 32) & 511;
        }
    }
    double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + s0.x + s1.x + s2.x + s3.x;
    if (r == 123.456) out[0] = r;
}

int main() {
    double* out;
    cudaMalloc(&out, 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int itf = 20000, its = 20000;
    for (int mode = 1; mode <= 3; ++mode) {
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            k<<<148, 512>>>(out, itf, its, mode);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        // 8 warps per role and SM: FP64 warp-instructions = 8 warps * 8 * iters ; LDS.128 wavefronts = 8 warps * 4 loads * 4 * iters
        printf("mode %d (%s): %.3f ms | FP64 %.2f warp-instr/clk/SM | LDS %.2f wavefronts/clk/SM\n", mode,
               mode == 1 ? "fp64 only" : mode == 2 ? "lds only" : "both", best,
               (mode & 1) ? 8.0 * 8 * itf / (best * 1e-3 * 1.965e9) : 0.0,
               (mode & 2) ? 8.0 * 4 * 4 * its / (best * 1e-3 * 1.965e9) : 0.0);
    }
    return 0;
}
