// Micro-benchmark: how fast can 24 warps/SM stream the row pass's sources with different buffer layouts?
//   mode 0: (patch k = i*n + j, u, PH)  -- the layout the column pass writes today
//   mode 1: (j, u, i, PH)              -- all column blocks of one (row-patch, row) contiguous
//   mode 2: contiguous streaming of the same number of bytes
// Each warp handles one pair of column blocks of one image row: NE dependent batches of 2*CNT 128-bit loads per lane.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int CNT>
__global__ void __launch_bounds__(128, 6) k(const double2* __restrict__ buf, double2* __restrict__ out, int mode, int n, int N,
                                           int ROWS, int PH, int dx, int shift, long total_elems) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double2 acc = make_double2(0.0, 0.0);
    for (long r = blockIdx.x; r < N; r += gridDim.x) {
        // covering row-patches: j with u = r + shift - j*dx in [0, ROWS)
        int jl[CNT], ul[CNT], cnt = 0;
        for (int j = 0; j < n && cnt < CNT; ++j) { int u = (int)r + shift - j * dx; if (u >= 0 && u < ROWS) { jl[cnt] = j; ul[cnt] = u; ++cnt; } }
        for (int c = cnt; c < CNT; ++c) { jl[c] = jl[0]; ul[c] = ul[0]; }
        const int ia = 2 * wid, ib = ia + 1;
        for (int e = 0; e < 7; ++e) {
            int k = lane + 32 * e;
            if (k >= PH) break;
            double2 v[2 * CNT];
#pragma unroll
            for (int q = 0; q < CNT; ++q) {
                long oa, ob;
                if (mode == 0) { oa = (((long)ia * n + jl[q]) * ROWS + ul[q]) * PH + k; ob = (((long)ib * n + jl[q]) * ROWS + ul[q]) * PH + k; }
                else if (mode == 1) { oa = (((long)jl[q] * ROWS + ul[q]) * n + ia) * PH + k; ob = oa + PH; }
                else { oa = ((r * 4 + wid) * 2 * CNT + 2 * q) * (long)PH + k; ob = oa + PH; oa %= total_elems; ob %= total_elems; }
                asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v[2 * q].x), "=d"(v[2 * q].y) : "l"(buf + oa));
                asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v[2 * q + 1].x), "=d"(v[2 * q + 1].y) : "l"(buf + ob));
            }
#pragma unroll
            for (int q = 0; q < 2 * CNT; ++q) { acc.x += v[q].x; acc.y += v[q].y; }
        }
    }
    if (acc.x == 1.2345) out[blockIdx.x] = acc;
}

int main() {
    const int n = 8, N = 1024, PH = 193, dx = 128;
    for (int adj = 0; adj < 2; ++adj) {
        const int ROWS = adj ? 256 : 384, shift = adj ? 64 : 128, planes = 3;
        const long elems = (long)n * n * ROWS * PH;      // one plane
        double2 *buf, *out, *flush;
        CK(cudaMalloc(&buf, elems * planes * sizeof(double2)));
        CK(cudaMemset(buf, 0, elems * planes * sizeof(double2)));
        CK(cudaMalloc(&out, 1 << 20));
        CK(cudaMalloc(&flush, 256 << 20));
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        for (int mode = 0; mode < 3; ++mode) {
            float best = 1e9;
            for (int rep = 0; rep < 5; ++rep) {
                CK(cudaMemsetAsync(flush, rep, 256 << 20));
                cudaEventRecord(a);
                // the three planes as one launch: rows of plane p offset by p*elems (here: just relaunch per plane on one stream)
                for (int p = 0; p < planes; ++p) {
                    if (adj) k<2><<<148 * 6, 128>>>(buf + p * elems, out, mode, n, N, ROWS, PH, dx, shift, elems);
                    else k<3><<<148 * 6, 128>>>(buf + p * elems, out, mode, n, N, ROWS, PH, dx, shift, elems);
                }
                cudaEventRecord(b);
                CK(cudaEventSynchronize(b));
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (ms < best) best = ms;
            }
            // bytes actually requested: rows * 4 warps * 7 batches... = N * n * CNT * PH * 16 per plane
            double bytes = (double)planes * N * n * (adj ? 2 : 3) * PH * 16.0;
            printf("%s mode %d: %.1f us  %.0f GB/s (%.0f MB)\n", adj ? "adj" : "fwd", mode, best * 1e3, bytes / best / 1e6, bytes / 1e6);
        }
        cudaFree(buf); cudaFree(out); cudaFree(flush);
    }
    return 0;
}
