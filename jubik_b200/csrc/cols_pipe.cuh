// cols_pipe.cuh -- building blocks of the column pass as a persistent TMA-fed pipeline (sm_100a): mbarrier /
// cp.async.bulk wrappers, the tile source (3-D tensor map) and the per-item geometry.  The kernel is in cols3.cuh.
//
// One CTA per SM loops over (plane, patch, strip) work items.  Thread 0 stages, two items ahead,
//   * the CB-column input tile of the patch's row spectra with cp.async.bulk.tensor.2d (TMA; columns
//     beyond L/2 are zero-filled by the hardware, row chunks outside the image are skipped), and
//   * the matching slice of the resident PSF spectrum with one 1-D cp.async.bulk (it is contiguous),
// into a 2-stage shared-memory ring guarded by mbarriers (full[s] carries the transaction bytes,
// done[s] is released by the warps as soon as their last shared-memory read of the stage is in
// registers; thread 0 then refills the slot with item it+2).  Eight warps run forward column FFT ->
// * K^ (or conj K^) -> inverse column FFT in place on the tile and store the result rows straight from
// registers (fire-and-forget), so HBM latency never sits on the FP64 critical path.
#pragma once
#include "fft_pass.cuh"

#ifndef JB_EMU
#include <cuda.h>
#endif

struct JbTileSrc {
#ifndef JB_EMU
    CUtensorMap map;   // 3-D fp64 view of a row-spectrum buffer (planes*n, N, PH): dim0 = 2*PH doubles (contiguous),
                       // dim1 = N image rows, dim2 = plane*n + i ; box = (2*CB, BR, 1).  Rows outside [0,N) and
                       // columns beyond PH are zero-filled by the TMA unit.
#endif
    const cd* base;    // same buffer, for the emulation build
    int PH;
    int BR;            // rows per TMA box
    int N;
};

#define JB_PIPE_COMPUTE_THREADS 256
#define JB_PIPE_THREADS JB_PIPE_COMPUTE_THREADS

#ifndef JB_EMU
JB_D uint32_t jb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
JB_D void jb_mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(jb_smem_u32(b)), "r"(count) : "memory");
}
JB_D void jb_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(jb_smem_u32(b)), "r"(bytes) : "memory");
}
JB_D void jb_mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(jb_smem_u32(b)) : "memory");
}
template <bool SLEEP>
JB_D void jb_mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done;
    unsigned spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(jb_smem_u32(b)), "r"(parity)
            : "memory");
        if (!done) {
            if (SLEEP) __nanosleep(128);
            if (++spins > (1u << 24)) __trap();  // a lost transaction becomes an error, not a hang
        }
    } while (!done);
}
JB_D void jb_tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(jb_smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(jb_smem_u32(bar))
        : "memory");
}
JB_D void jb_bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(jb_smem_u32(dst)), "l"(src), "r"(bytes), "r"(jb_smem_u32(bar))
                 : "memory");
}
// warm L2 with a contiguous global range (no shared-memory destination, no completion tracking)
JB_D void jb_bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
JB_D void jb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
JB_D void jb_bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(JB_PIPE_COMPUTE_THREADS) : "memory"); }
#else
inline void jb_bar_compute() { JB_SYNC(); }
#endif

// per-item geometry shared by producer and consumers
struct JbColItem {
    int strip, k, plane, te, j, i;
    int in_p0, in_rows, vlo, vhi, out_p0, out_rows;
    int row0;         // image row of tile row in_p0 (may be negative / run past N: zero rows)
    int in_slab;      // in_plane*n + i : which (plane, column block) slab of the row-spectrum buffer
    size_t out_off;   // element offset of this patch in the output buffer
    size_t k_off;     // element offset of this (te, patch, strip) slice of K^
};

template <int L, int CB, bool ADJ>
JB_D JbColItem jb_col_item(const JbGeom& g, long item, int nstrips) {
    JbColItem it;
    const int nn2 = g.n * g.n;
    it.strip = (int)(item % nstrips);
    long r = item / nstrips;
    it.k = (int)(r % nn2);
    int se;
    jb_map_ste(g, (int)(r / nn2), it.plane, it.te, se);
    it.j = it.k % g.n;
    it.i = it.k / g.n;
    int in_plane;
    if (!ADJ) {
        it.in_p0 = g.m; it.in_rows = g.w; it.out_p0 = 0; it.out_rows = g.P;
        in_plane = se;
        it.row0 = g.offs[it.j] - g.hp;
    } else {
        it.in_p0 = 0; it.in_rows = g.P; it.out_p0 = g.m; it.out_rows = g.w;
        in_plane = it.plane;
        it.row0 = g.offs[it.j] - g.hp - g.m;
    }
    // tile rows whose image row lies in [0,N)
    int lo = -it.row0, hi = g.N - it.row0;
    it.vlo = it.in_p0 + (lo > 0 ? lo : 0);
    it.vhi = it.in_p0 + (hi < it.in_rows ? hi : it.in_rows);
    it.in_slab = in_plane * g.n + it.i;
    it.out_off = (((size_t)it.plane * nn2 + it.k) * (size_t)it.out_rows) * g.PH;
    it.k_off = (((size_t)it.te * nn2 + it.k) * nstrips + it.strip) * ((size_t)L * CB);
    return it;
}
