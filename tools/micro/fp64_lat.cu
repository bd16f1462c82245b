// Micro-benchmark: FP64 issue rate per SM sub-partition as a function of the independent DFMA chains in flight
// (ILP per warp x warps per sub-partition).  Tells how much parallelism an FP64 butterfly stage needs on B200.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, int iters) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
    const double m = 0.999999, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = a[i] * m + c;
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += a[i];
    if (r == 123.456) out[0] = r;
}

template <int ILP>
void run(double* out, int warps_per_smsp) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 100000 / ILP * 4;
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        k<ILP><<<148, 128 * warps_per_smsp>>>(out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    double cyc = best * 1e-3 * 1.965e9;
    double per_smsp = (double)warps_per_smsp * ILP * iters / cyc;   // warp-instructions per clock and sub-partition
    printf("ILP %2d x %d warps/SMSP = %3d chains: %.3f DFMA warp-instr/clk/SMSP (%.1f cycles per dependent DFMA)\n", ILP,
           warps_per_smsp, ILP * warps_per_smsp, per_smsp, cyc / iters);
}

int main() {
    double* out;
    cudaMalloc(&out, 8);
    for (int w = 1; w <= 4; w *= 2) {
        run<1>(out, w); run<2>(out, w); run<4>(out, w); run<8>(out, w); run<16>(out, w);
    }
    return 0;
}
