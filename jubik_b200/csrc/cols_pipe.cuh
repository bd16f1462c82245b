// cols_pipe.cuh -- the column pass as a persistent, warp-specialised, TMA-fed pipeline (sm_100a).
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
    CUtensorMap map;   // 2-D fp64 view: dim0 = 2*PH doubles (contiguous), dim1 = all rows of the buffer
#endif
    const cd* base;    // same buffer, for the emulation build
    int PH;
    int BR;            // rows per TMA box
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
JB_D void jb_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(jb_smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(jb_smem_u32(bar))
        : "memory");
}
JB_D void jb_bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(jb_smem_u32(dst)), "l"(src), "r"(bytes), "r"(jb_smem_u32(bar))
                 : "memory");
}
JB_D void jb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
JB_D void jb_bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(JB_PIPE_COMPUTE_THREADS) : "memory"); }
#else
inline void jb_bar_compute() { JB_SYNC(); }
#endif

// per-item geometry shared by producer and consumers
struct JbColItem {
    int strip, k, plane, te, j;
    int in_p0, in_rows, vlo, vhi, out_p0, out_rows;
    size_t in_row0;   // first row of this patch in the input buffer (row units)
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
    it.plane = (int)(r / nn2);
    it.j = it.k % g.n;
    const int TE = g.T * g.E;
    const int s = it.plane / TE;
    it.te = it.plane - s * TE;
    const int e = it.te % g.E;
    size_t in_plane;
    if (!ADJ) {
        it.in_p0 = g.m; it.in_rows = g.w; it.out_p0 = 0; it.out_rows = g.P;
        in_plane = (size_t)s * g.E + e;
        int lo = g.hp - g.offs[it.j], hi = g.N + g.hp - g.offs[it.j];
        it.vlo = g.m + (lo > 0 ? lo : 0);
        it.vhi = g.m + (hi < g.w ? hi : g.w);
    } else {
        it.in_p0 = 0; it.in_rows = g.P; it.out_p0 = g.m; it.out_rows = g.w;
        in_plane = (size_t)it.plane;
        int lo = g.hp + g.m - g.offs[it.j], hi = g.N + g.hp + g.m - g.offs[it.j];
        it.vlo = lo > 0 ? lo : 0;
        it.vhi = hi < g.P ? hi : g.P;
    }
    it.in_row0 = (in_plane * nn2 + it.k) * (size_t)it.in_rows;
    it.out_off = (((size_t)it.plane * nn2 + it.k) * (size_t)it.out_rows) * g.PH;
    it.k_off = (((size_t)it.te * nn2 + it.k) * nstrips + it.strip) * ((size_t)L * CB);
    return it;
}

// grid = min(items, SMs) persistent CTAs, block = JB_PIPE_THREADS, dynamic smem = 2*(2*L*CB*16) + 64.
template <class FP, int CB, bool ADJ>
JB_GLOBAL void
#ifndef JB_EMU
__maxnreg__(224)
#endif
k_cols_pipe(const
#ifndef JB_EMU
            __grid_constant__
#endif
            JbTileSrc src, JbGeom g, cd* __restrict__ out, const cd* __restrict__ khat,
            const cd* __restrict__ tw_nk, int planes) {
    constexpr int L = FP::L, R1 = FP::R1, R2 = FP::R2;
    static_assert(R2 > 2 && R1 > 2, "the pipelined column pass is for two-pass lengths");
    static_assert(R2 * CB <= JB_PIPE_COMPUTE_THREADS && R1 * CB <= JB_PIPE_COMPUTE_THREADS,
                  "one work item per thread and pass (thread-constant twiddle seeds)");
    constexpr int TILE = L * CB;  // cd per tile
    JB_SMEM(unsigned char, smraw);
    cd* tile0 = reinterpret_cast<cd*>(smraw);
    cd* kh0 = tile0 + 2 * TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(kh0 + 2 * TILE);  // full[0], full[1], done[0], done[1]
    const int nstrips = (g.PH + CB - 1) / CB;
    const long nitems = (long)planes * g.n * g.n * nstrips;
    const int tid = threadIdx.x;
    const int PH = g.PH;

#ifndef JB_EMU
    // Thread 0 doubles as the producer: it stages item it+2 into a ring slot as soon as every warp has
    // pulled its last operands of item it out of that slot (a separate producer warp would cost a whole
    // 4-warp register allocation granule and cap the compute warps at 168 registers).
    auto issue_loads = [&](long item_, int s_) {
        JbColItem ci = jb_col_item<L, CB, ADJ>(g, item_, nstrips);
        cd* tile = tile0 + s_ * TILE;
        const int BR = src.BR;
        int nch = (ci.in_rows + BR - 1) / BR, issued = 0;
        for (int c = 0; c < nch; ++c) {  // row chunks of the input tile that intersect the in-image range
            int r0 = c * BR;
            if (r0 + BR > ci.in_rows) r0 = ci.in_rows - BR;
            int p0 = ci.in_p0 + r0;
            if (p0 + BR > ci.vlo && p0 < ci.vhi) ++issued;
        }
        jb_mbar_expect_tx(&bars[s_], (uint32_t)(issued * BR * CB * sizeof(cd) + TILE * sizeof(cd)));
        for (int c = 0; c < nch; ++c) {
            int r0 = c * BR;
            if (r0 + BR > ci.in_rows) r0 = ci.in_rows - BR;
            int p0 = ci.in_p0 + r0;
            if (p0 + BR > ci.vlo && p0 < ci.vhi)
                jb_tma_load_2d(tile + (size_t)p0 * CB, &src.map, 2 * ci.strip * CB, (int)(ci.in_row0 + r0), &bars[s_]);
        }
        jb_bulk_load_1d(kh0 + s_ * TILE, khat + ci.k_off, (uint32_t)(TILE * sizeof(cd)), &bars[s_]);
    };
    if (tid == 0) {
        jb_mbar_init(&bars[0], 1);
        jb_mbar_init(&bars[1], 1);
        jb_mbar_init(&bars[2], JB_PIPE_THREADS / 32);
        jb_mbar_init(&bars[3], JB_PIPE_THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((long)blockIdx.x < nitems) issue_loads(blockIdx.x, 0);
        if ((long)blockIdx.x + gridDim.x < nitems) issue_loads((long)blockIdx.x + gridDim.x, 1);
    }
    __syncthreads();
#endif
    // ---------------------------------------------------------------------- compute warps
    // The thread <-> (n2 | k1) mapping is the same for every item, so the inter-pass twiddles
    // W_L^{n2 k1} are generated from two thread-constant seeds by a two-chain recurrence
    // (odd/even powers): no table loads on the critical path, error <= ~R/2 roundings.
    cd p1w1 = cd_make(1.0, 0.0), p1w2 = p1w1, p2w1 = p1w1, p2w2 = p1w1;
    {
        int n2 = tid / CB, k1 = tid / CB;
        if (n2 < R2) { p1w1 = tw_nk[1 * R2 + n2]; p1w2 = tw_nk[2 * R2 + n2]; }            // W^{n2}, W^{2 n2}
        if (k1 < R1) { p2w1 = cd_conj(tw_nk[k1 * R2 + 1]); p2w2 = cd_conj(tw_nk[k1 * R2 + 2]); }  // W^{-k1}, W^{-2 k1}
    }
    long item = blockIdx.x;
    for (unsigned it = 0; item < nitems; ++it, item += gridDim.x) {
        const int s = it & 1;
        JbColItem ci = jb_col_item<L, CB, ADJ>(g, item, nstrips);
        cd* tile = tile0 + s * TILE;
        const cd* kh = kh0 + s * TILE;
#ifndef JB_EMU
        jb_mbar_wait<false>(&bars[s], (it >> 1) & 1);
#else
        // emulation: synchronous staging of the same tiles (tests/emu only)
        JB_SYNC();
        for (int idx = tid; idx < ci.in_rows * CB; idx += JB_PIPE_COMPUTE_THREADS) {
            int r = idx / CB, b = idx - r * CB, col = ci.strip * CB + b;
            tile[(size_t)(ci.in_p0 + r) * CB + b] =
                col < PH ? src.base[(ci.in_row0 + r) * PH + col] : cd_make(0.0, 0.0);
        }
        for (int idx = tid; idx < TILE; idx += JB_PIPE_COMPUTE_THREADS) const_cast<cd*>(kh)[idx] = khat[ci.k_off + idx];
        JB_SYNC();
#endif
        // pass 1: forward radix-R1 over n1, in place (index set {R2*a + n2} is closed under the exchange)
        for (int w = tid; w < R2 * CB; w += JB_PIPE_COMPUTE_THREADS) {
            int n2 = w / CB, b = w - n2 * CB;
            cd v[R1];
            static_for<0, R1>([&](auto nn) {
                constexpr int n1 = nn;
                int p = R2 * n1 + n2;
                v[n1] = (p >= ci.vlo && p < ci.vhi) ? tile[p * CB + b] : cd_make(0.0, 0.0);
            });
            dft<R1, -1>(v);
            cd po = p1w1, pe = p1w2;  // W^{n2 k1} for odd / even k1
            static_for<0, R1>([&](auto kk) {
                constexpr int k1 = kk;
                cd y = v[k1];
                if constexpr (k1 > 0) {
                    if constexpr (k1 & 1) { y = cd_mul(y, po); if constexpr (k1 + 2 < R1) po = cd_mul(po, p1w2); }
                    else { y = cd_mul(y, pe); if constexpr (k1 + 2 < R1) pe = cd_mul(pe, p1w2); }
                }
                tile[(k1 * R2 + n2) * CB + b] = y;
            });
        }
        jb_bar_compute();
        // pass 2: forward radix-R2, spectrum multiply, inverse radix-R2, in place
        for (int w = tid; w < R1 * CB; w += JB_PIPE_COMPUTE_THREADS) {
            int k1 = w / CB, b = w - k1 * CB;
            cd u[R2];
            cd* col = tile + (size_t)(k1 * R2) * CB + b;
            static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; u[n2] = col[n2 * CB]; });
            dft<R2, -1>(u);
            static_for<0, R2>([&](auto kk) {
                constexpr int k2 = kk;
                cd kv = kh[(k2 * R1 + k1) * CB + b];
                u[k2] = ADJ ? cd_mulc(u[k2], kv) : cd_mul(u[k2], kv);
            });
            dft<R2, +1>(u);
            cd po = p2w1, pe = p2w2;  // W^{-k1 n2} for odd / even n2
            static_for<0, R2>([&](auto nn) {
                constexpr int n2 = nn;
                cd y = u[n2];
                if constexpr (n2 > 0) {
                    if constexpr (n2 & 1) { y = cd_mul(y, po); if constexpr (n2 + 2 < R2) po = cd_mul(po, p2w2); }
                    else { y = cd_mul(y, pe); if constexpr (n2 + 2 < R2) pe = cd_mul(pe, p2w2); }
                }
                col[n2 * CB] = y;
            });
        }
        jb_bar_compute();
        // pass B: inverse radix-R1; the stage is released as soon as its last read is in registers
        constexpr int NB = (R2 * CB + JB_PIPE_COMPUTE_THREADS - 1) / JB_PIPE_COMPUTE_THREADS;
        static_assert(NB == 1, "pass B assumes one item per thread (R2*CB <= compute threads)");
        {
            const int w = tid;
            const bool act = w < R2 * CB;
            int n2 = w / CB, b = w - n2 * CB;
            cd v[R1];
            if (act) static_for<0, R1>([&](auto kk) { constexpr int k1 = kk; v[k1] = tile[(k1 * R2 + n2) * CB + b]; });
#ifndef JB_EMU
            jb_fence_proxy_async();
            __syncwarp();
            if ((tid & 31) == 0) jb_mbar_arrive(&bars[2 + s]);
            if (tid == 0) {
                long nxt = item + 2 * (long)gridDim.x;
                if (nxt < nitems) {
                    jb_mbar_wait<false>(&bars[2 + s], (it >> 1) & 1);
                    issue_loads(nxt, s);
                }
            }
#endif
            if (act) {
                dft<R1, +1>(v);
                const int colg = ci.strip * CB + b;
                cd* ob = out + ci.out_off + colg;
                static_for<0, R1>([&](auto nn) {
                    constexpr int n1 = nn;
                    int p = R2 * n1 + n2;
                    if (colg < PH && p >= ci.out_p0 && p < ci.out_p0 + ci.out_rows)
                        ob[(size_t)(p - ci.out_p0) * PH] = v[n1];
                });
            }
        }
    }
}
