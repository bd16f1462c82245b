// jb_rt.h -- thin runtime layer shared by the CUDA build (nvcc, sm_100a) and the host-thread
// emulation build that tests/emu uses to debug kernel index logic on a box without a GPU.
//
// The PRODUCT is the CUDA build.  The emulation (JB_EMU) exists only under tests/emu, is never
// loaded by the jubik_b200 package and is not a fallback: jubik_b200/_lib.py refuses to run
// without libjubik_b200.so and a CUDA device.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <type_traits>
#include <utility>

#ifndef JB_EMU
// ------------------------------------------------------------------------------- CUDA
#include <cuda_runtime.h>
#define JB_HD __host__ __device__ __forceinline__
#define JB_D __device__ __forceinline__
#define JB_GLOBAL __global__
#define JB_GLOBAL_S static __global__
#define JB_LB(t, b) __launch_bounds__(t, b)
#define JB_SMEM(T, name) extern __shared__ __align__(128) unsigned char jb_smem_raw_[]; \
                         T* name = reinterpret_cast<T*>(jb_smem_raw_)
typedef cudaStream_t jb_stream_t;
typedef double2 cd;

template <class T> JB_D T jb_ldg(const T* p) { return __ldg(p); }
// streaming (evict-first) 128-bit load for data that is read exactly once per launch
JB_D cd jb_ld_stream(const cd* p) {
    cd v;
    asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
// read-only 128-bit load that does not allocate in L1 (data used once per CTA)
JB_D cd jb_ld_nc(const cd* p) {
    cd v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
// L2-only (no L1 allocation) load: data another thread of this CTA wrote a moment ago
JB_D cd jb_ld_cg(const cd* p) {
    cd v;
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
JB_D void jb_st_stream(cd* p, cd v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
#define JB_SYNC() __syncthreads()
#define JB_SYNCWARP() __syncwarp()
// named barrier over warp-aligned group `grp` (of NG groups, hardware barriers 1..NG) of `nthr` threads; `pred` is uniform per warp
// The id is spelled as an IMMEDIATE (one branch per group, warp-uniform): with a register operand ptxas reserves all 16
// hardware barriers for the CTA and a second CTA no longer fits the SM.
template <int NG, int I = 0>
__device__ __forceinline__ void jb_group_sync_imm(int grp, int nthr) {
    if constexpr (I < NG) {
        if (grp == I) asm volatile("bar.sync %0, %1;" ::"n"(I + 1), "r"(nthr) : "memory");
        else jb_group_sync_imm<NG, I + 1>(grp, nthr);
    }
}
#define JB_GROUP_SYNC(pred, NG, grp, nthr) do { if (pred) jb_group_sync_imm<NG>(grp, nthr); } while (0)
JB_D unsigned jb_atomic_inc(unsigned* p) { return atomicAdd(p, 1u); }
JB_D void jb_threadfence() { __threadfence(); }
JB_D double jb_ld_cg_f64(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

template <class K, class... A>
inline cudaError_t jb_launch(K kern, dim3 grid, dim3 block, size_t smem, jb_stream_t st, A... args) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, block, smem, st>>>(args...);
    return cudaGetLastError();
}
#else
// ------------------------------------------------------------------------------- emulation
#include <barrier>
#include <thread>
#include <vector>
#include <memory>
#define JB_HD inline
#define JB_D inline
#define JB_GLOBAL
#define JB_GLOBAL_S static
#define JB_LB(t, b)
struct cd { double x, y; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct jb_emu_ctx {
    dim3 tid_, bid_, bdim_, gdim_;
    unsigned char* smem;
    std::barrier<>* bar;
};
extern thread_local jb_emu_ctx jb_ctx;
#define threadIdx (jb_ctx.tid_)
#define blockIdx (jb_ctx.bid_)
#define blockDim (jb_ctx.bdim_)
#define gridDim (jb_ctx.gdim_)
#define JB_SMEM(T, name) T* name = reinterpret_cast<T*>(jb_ctx.smem)
#define JB_SYNC() jb_ctx.bar->arrive_and_wait()
// emulated threads are not grouped in warps: a warp-level sync becomes a CTA-level one (callers keep the number
// of warp syncs uniform across the CTA)
#define JB_SYNCWARP() jb_ctx.bar->arrive_and_wait()
// ... and so does a named group barrier (every thread of the CTA reaches the call, members or not)
#define JB_GROUP_SYNC(pred, NG, grp, nthr) jb_ctx.bar->arrive_and_wait()
typedef void* jb_stream_t;
typedef int cudaError_t;
#define cudaSuccess 0
template <class T> inline T jb_ldg(const T* p) { return *p; }
inline cd jb_ld_stream(const cd* p) { return *p; }
inline cd jb_ld_cg(const cd* p) { return *p; }
inline unsigned jb_atomic_inc(unsigned* p) { return (*p)++; }   // emulated CTAs run one after the other
inline void jb_threadfence() {}
inline double jb_ld_cg_f64(const double* p) { return *p; }
inline cd jb_ld_nc(const cd* p) { return *p; }
inline void jb_st_stream(cd* p, cd v) { *p = v; }
inline cd make_double2(double x, double y) { return cd{x, y}; }

template <class K, class... A>
inline cudaError_t jb_launch(K kern, dim3 grid, dim3 block, size_t smem, jb_stream_t, A... args) {
    unsigned nt = block.x * block.y * block.z;
    std::barrier<> bar((std::ptrdiff_t)nt);
    std::vector<unsigned char> sm(smem + 64);
    unsigned char* smp = sm.data();
    smp += (16 - (reinterpret_cast<uintptr_t>(smp) & 15)) & 15;
    auto body = [&](unsigned tid) {
        jb_ctx.bdim_ = block;
        jb_ctx.gdim_ = grid;
        jb_ctx.smem = smp;
        jb_ctx.bar = &bar;
        jb_ctx.tid_ = dim3(tid % block.x, (tid / block.x) % block.y, tid / (block.x * block.y));
        for (unsigned bz = 0; bz < grid.z; ++bz)
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    jb_ctx.bid_ = dim3(bx, by, bz);
                    kern(args...);
                    bar.arrive_and_wait();
                }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(body, t);
    body(0);
    for (auto& t : th) t.join();
    return 0;
}
#endif

// ------------------------------------------------------------------------------- complex helpers
JB_HD cd cd_make(double x, double y) { cd r; r.x = x; r.y = y; return r; }
JB_HD cd cd_add(cd a, cd b) { return cd_make(a.x + b.x, a.y + b.y); }
JB_HD cd cd_sub(cd a, cd b) { return cd_make(a.x - b.x, a.y - b.y); }
JB_HD cd cd_mul(cd a, cd b) { return cd_make(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
JB_HD cd cd_mulc(cd a, cd b) { return cd_make(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a*conj(b)
JB_HD cd cd_conj(cd a) { return cd_make(a.x, -a.y); }
JB_HD cd cd_scale(cd a, double s) { return cd_make(a.x * s, a.y * s); }

// compile-time loop with the index available as a constant expression
template <int I, int N, class F>
JB_HD void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}
