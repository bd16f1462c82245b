// Micro-benchmark: one "FFT stage" as the column pass runs it -- every thread loads R complex values from shared memory
// (LDS.128), runs NFP FP64 instructions on them (R independent chains), stores them back (STS.128), block barrier --
// timed as (a) shared-memory traffic only, (b) FP64 only, (c) both, for CTA shapes 2 x 256 and 1 x 512 threads per SM.
// Question: how much of max(a, b) vs a + b does this structure reach on a B200 SM (what the column pass can hope for)?
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int NFP, bool SMEM, bool SYNC>
__global__ void k(double* out, int stages, int stride) {
    extern __shared__ double2 sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < R * nt; i += nt) sm[i] = make_double2(i * 1e-3, -i * 1e-3);
    __syncthreads();
    double2 v[R];
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = make_double2(tid + i, tid - i);
    const double m = 0.999999, c = 1e-9;
    for (int s = 0; s < stages; ++s) {
        if (SMEM) {
#pragma unroll
            for (int i = 0; i < R; ++i) { double2 t = sm[tid + i * stride]; v[i].x += t.x; v[i].y += t.y; }
        }
#pragma unroll
        for (int j = 0; j < NFP / (2 * R); ++j) {
#pragma unroll
            for (int i = 0; i < R; ++i) { v[i].x = v[i].x * m + c; v[i].y = v[i].y * m + v[(i + 1) % R].x; }
        }
        if (SMEM) {
#pragma unroll
            for (int i = 0; i < R; ++i) sm[tid + i * stride] = v[i];
        }
        if (SYNC) __syncthreads();
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) r += v[i].x + v[i].y;
    if (r == 123.456) out[0] = r;
}

template <int R, int NFP, bool SMEM, bool SYNC>
float run(int ctas_per_sm, int threads, int stages, double* out) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const size_t smem = (size_t)R * threads * sizeof(double2);
    cudaFuncSetAttribute(k<R, NFP, SMEM, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        k<R, NFP, SMEM, SYNC><<<148 * ctas_per_sm, threads, smem>>>(out, stages, threads);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

template <int R, int NFP>
void suite(const char* tag, int cps, int threads, double* out) {
    const int stages = 4000;
    const double clk = 1.965e6;  // cycles per ms
    float t_s = run<R, 0, true, true>(cps, threads, stages, out);
    float t_f = run<R, NFP, false, true>(cps, threads, stages, out);
    float t_b = run<R, NFP, true, true>(cps, threads, stages, out);
    float t_n = run<R, NFP, true, false>(cps, threads, stages, out);
    // per stage and SM: wavefronts = warps * R * 4 (LDS.128) * 2 (load + store); FP64 warp-instr = warps * NFP
    const double warps = cps * threads / 32.0;
    printf("%s R=%d NFP=%d: smem only %.0f clk/stage (%.0f wavefronts) | fp64 only %.0f clk (%.0f per SP at 2 clk/instr) | both %.0f clk "
           "(sum %.0f, max %.0f) | both, no barrier %.0f clk\n", tag, R, NFP, t_s * clk / stages, warps * R * 8, t_f * clk / stages,
           warps * NFP * 2 / 4, t_b * clk / stages, (t_s + t_f) * clk / stages, (t_s > t_f ? t_s : t_f) * clk / stages,
           t_n * clk / stages);
}

int main() {
    double* out;
    cudaMalloc(&out, 8);
    suite<8, 96>("2x256", 2, 256, out);
    suite<8, 96>("1x512", 1, 512, out);
    suite<8, 96>("1x384", 1, 384, out);
    suite<8, 192>("2x256", 2, 256, out);
    suite<12, 240>("1x224", 1, 224, out);
    suite<4, 48>("4x256", 4, 256, out);
    return 0;
}
