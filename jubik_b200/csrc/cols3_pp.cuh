// cols3_pp.cuh -- the three-stage column pass (cols3.cuh) as ONE 2 x THREADS CTA per SM whose two halves ("groups")
// run independent item streams in enforced ANTI-PHASE.
//
// Why: in k_cols3_pipe the two co-resident CTAs drift into lock-step -- both in their shared-memory phase (loads /
// stores of a stage), then both in their FP64 phase (the butterflies) -- so the kernel time is the SUM of the FP64
// issue time (~1.5 k cycles per item) and the shared-memory wavefront time (~2.5 k), although the two pipes overlap
// perfectly when different warps use them (tools/micro/fp64_smem.cu).  Here a token (two named barriers, bar.arrive /
// bar.sync over all 2 x THREADS threads) makes the FP64 phases of the two groups mutually exclusive: while one group
// runs a stage's butterflies the other one stores / exchanges / loads, and vice versa, so the SM sees both pipes busy.
//
// Same maths, index maps, K^ layout and twiddle tables as Cols3; only for tiles of <= 256 threads per group
// (register file: 2 x 256 threads x 128 registers).  No sweep / direct-K^ / swizzle variants.
#pragma once
#include "cols3.cuh"

#ifndef JB_EMU
JB_D void jb_nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
JB_D void jb_nbar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// an opaque zero: per-stage index arithmetic that depends on it is recomputed where it is used instead of being hoisted
// out of the item loop (hoisted, it pins ~40 registers and the stage operands spill around the barriers)
JB_D int jb_opaque_zero() { int z; asm volatile("mov.u32 %0, 0;" : "=r"(z)); return z; }

template <class F3, int CB, bool ADJ>
JB_GLOBAL void JB_LB((2 * Cols3<F3, CB>::THREADS), 1)
k_cols3_pp(const __grid_constant__ JbTileSrc src, JbGeom g, cd* __restrict__ out, const cd* __restrict__ khat,
           const cd* __restrict__ tw1g, const cd* __restrict__ tw2g, const double* __restrict__ wgt, int planes) {
    using C3 = Cols3<F3, CB>;
    constexpr int L = F3::L, R1 = F3::R1, R2 = F3::R2, R3 = F3::R3, M = F3::M;
    constexpr int TILE = L * CB, GT = C3::THREADS, ALL = 2 * GT;
    constexpr int I1 = C3::I1, I2 = C3::I2, I3 = C3::I3;
    JB_SMEM(unsigned char, smraw);
    const int grp = threadIdx.x / GT, tid = threadIdx.x - grp * GT;
    // shared memory: [group 0: tile x2, K^ x2][group 1: tile x2, K^ x2][tw1][tw2][8 mbarriers][4 item infos][row weights]
    cd* tile0 = reinterpret_cast<cd*>(smraw) + (size_t)grp * 4 * TILE;
    cd* kh0 = tile0 + 2 * TILE;
    cd* tw1 = reinterpret_cast<cd*>(smraw) + 8 * TILE;
    cd* tw2 = tw1 + L;
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(tw2 + M);
    uint64_t* bars = bars_all + grp * 4;
    JbColInfo* info = reinterpret_cast<JbColInfo*>(bars_all + 8) + grp * 2;
    double* rw = reinterpret_cast<double*>(reinterpret_cast<JbColInfo*>(bars_all + 8) + 4);
    const int PH = g.PH;
    const int nstrips = (PH + CB - 1) / CB;
    const int nitems = planes * g.n * g.n * nstrips;     // < 2^31 (checked by the launcher)
    const int vcta = (int)blockIdx.x * 2 + grp, vgrid = (int)gridDim.x * 2;
    auto item_of = [&](int it_) -> int {
        const long long v = (long long)vcta + (long long)it_ * vgrid;
        return v < nitems ? (int)v : -1;
    };
    // both groups run the same number of rounds (the token is handed over five times per round)
    const int first0 = (int)blockIdx.x * 2;
    const int rounds = first0 < nitems ? (nitems - first0 + vgrid - 1) / vgrid : 0;
    const int bar_grp = 4 + grp;            // this group's stage barrier (GT threads)
    const int bar_mine = 2 + grp;           // "group grp may use the FP64 pipe"
    const int bar_other = 2 + (1 - grp);
    for (int i = threadIdx.x; i < L; i += ALL) tw1[i] = tw1g[i];
    for (int i = threadIdx.x; i < M; i += ALL) tw2[i] = tw2g[i];
    if (!ADJ) for (int i = threadIdx.x; i < g.w; i += ALL) rw[i] = wgt[i];
    auto issue_loads = [&](int item_, int s_) {
        JbColItem ci = jb_col_item<L, CB, ADJ>(g, item_, nstrips);
        JbColInfo& f = info[s_];   // ordered before the consumers' reads by the release/acquire of full[s_]
        f.out_off = (long long)ci.out_off + (long long)ci.strip * CB;
        f.vlo = ci.vlo; f.vhi = ci.vhi; f.out_p0 = ci.out_p0; f.out_rows = ci.out_rows;
        int nc = PH - ci.strip * CB;
        f.ncols = nc < CB ? nc : CB;
        cd* tile = tile0 + s_ * TILE;
        const int BR = src.BR;
        int nch = (ci.in_rows + BR - 1) / BR, issued = 0;
        for (int c = 0; c < nch; ++c) {
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
                jb_tma_load_3d(tile + (size_t)p0 * CB, &src.map, 2 * ci.strip * CB, ci.row0 + r0, ci.in_slab, &bars[s_]);
        }
        jb_bulk_load_1d(kh0 + s_ * TILE, khat + ci.k_off, (uint32_t)(TILE * sizeof(cd)), &bars[s_]);
    };
    if (tid == 0) {
        jb_mbar_init(&bars[0], 1);
        jb_mbar_init(&bars[1], 1);
        jb_mbar_init(&bars[2], GT / 32);
        jb_mbar_init(&bars[3], GT / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (item_of(0) >= 0) issue_loads(item_of(0), 0);
        if (item_of(1) >= 0) issue_loads(item_of(1), 1);
    }
    __syncthreads();
    if (grp == 1) jb_nbar_arrive(2, ALL);   // group 0 computes first
#define JB_PP_ACQ() jb_nbar_sync(bar_mine, ALL)
#define JB_PP_REL() jb_nbar_arrive(bar_other, ALL)
#define JB_PP_GSYNC() jb_nbar_sync(bar_grp, GT)
    for (int it = 0; it < rounds; ++it) {
        const int item = item_of(it);
        if (item < 0) {   // this group has run out of items: keep handing the token back
            for (int k = 0; k < 5; ++k) { JB_PP_ACQ(); JB_PP_REL(); }
            continue;
        }
        const int s = (int)(it & 1);
        cd* tile = tile0 + s * TILE;
        const cd* kh = kh0 + s * TILE;
        jb_mbar_wait<false>(&bars[s], (uint32_t)((it >> 1) & 1));
        const JbColInfo* fi = &info[s];   // fields are read where they are used (registers are the scarce resource here)
        // ---------------- stage 1 forward: items (m, b), DFT_R1 over n1, * W_L^{m k1}
        {
            const int t = tid;
            const bool act = t < I1;
            const int m = t / CB, b = t - m * CB;
            const int vlo = fi->vlo, vhi = fi->vhi;
            cd v[R1];
            JB_PP_ACQ();
            if (act) {
                static_for<0, R1>([&](auto nn) {
                    constexpr int n1 = nn;
                    const int p = M * n1 + m;
                    cd x = cd_make(0.0, 0.0);
                    if (p >= vlo && p < vhi) {
                        x = tile[p * CB + b];
                        if (!ADJ) x = cd_scale(x, rw[p - g.m]);
                    }
                    v[n1] = x;
                });
                dft<R1, -1>(v);
                static_for<1, R1>([&](auto kk) { constexpr int k1 = kk; v[k1] = cd_mul(v[k1], tw1[k1 * M + m]); });
            }
            JB_PP_REL();
            if (act) static_for<0, R1>([&](auto kk) { constexpr int k1 = kk; tile[(k1 * M + m) * CB + b] = v[k1]; });
        }
        JB_PP_GSYNC();
        // ---------------- stage 2 forward: items (k1, n3, b), DFT_R2 over n2, * W_M^{n3 k2}
        {
            const int t = tid;
            const bool act = t < I2;
            const int c = t / CB, b = t - c * CB;
            const int k1 = c / R3, n3 = c - k1 * R3;
            cd* base = tile + (size_t)(k1 * M + n3) * CB + b;
            cd v[R2];
            JB_PP_ACQ();
            if (act) {
                static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; v[n2] = base[(R3 * n2) * CB]; });
                dft<R2, -1>(v);
                static_for<1, R2>([&](auto kk) { constexpr int k2 = kk; v[k2] = cd_mul(v[k2], tw2[k2 * R3 + n3]); });
            }
            JB_PP_REL();
            if (act) static_for<0, R2>([&](auto kk) { constexpr int k2 = kk; base[(R3 * k2) * CB] = v[k2]; });
        }
        JB_PP_GSYNC();
        // ---------------- stage 3: DFT_R3 over n3, * K^ (conj for the adjoint), inverse DFT_R3, in registers
        {
            const int t = tid;
            const bool act = t < I3;
            const int c = t / CB, b = t - c * CB;
            cd* base = tile + (size_t)(R3 * c) * CB + b;
            cd v[R3];
            JB_PP_ACQ();
            if (act) {
                static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; v[n3] = base[n3 * CB]; });
                dft<R3, -1>(v);
                static_for<0, R3>([&](auto kk) {
                    constexpr int k3 = kk;
                    const cd kv = kh[k3 * I3 + t];
                    v[k3] = ADJ ? cd_mulc(v[k3], kv) : cd_mul(v[k3], kv);
                });
                dft<R3, +1>(v);
            }
            JB_PP_REL();
            if (act) static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; base[n3 * CB] = v[n3]; });
        }
        JB_PP_GSYNC();
        // ---------------- stage 2 inverse
        {
            const int t = tid;
            const bool act = t < I2;
            const int c = t / CB, b = t - c * CB;
            const int k1 = c / R3, n3 = c - k1 * R3;
            cd* base = tile + (size_t)(k1 * M + n3) * CB + b;
            cd v[R2];
            JB_PP_ACQ();
            if (act) {
                static_for<0, R2>([&](auto kk) { constexpr int k2 = kk; v[k2] = base[(R3 * k2) * CB]; });
                static_for<1, R2>([&](auto kk) { constexpr int k2 = kk; v[k2] = cd_mulc(v[k2], tw2[k2 * R3 + n3]); });
                dft<R2, +1>(v);
            }
            JB_PP_REL();
            if (act) static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; base[(R3 * n2) * CB] = v[n2]; });
        }
        JB_PP_GSYNC();
        // ---------------- stage 1 inverse: operands to registers, slot released, butterflies, rows stored from registers
        {
            const int t = tid;
            const bool act = t < I1;
            const int m = t / CB, b = t - m * CB;
            cd v[R1];
            const long long out_off = fi->out_off;     // read before the slot (and its info) is handed back to the producer
            const int out_p0 = fi->out_p0, out_rows = fi->out_rows, ncols = fi->ncols;
            JB_PP_ACQ();
            if (act) {
                static_for<0, R1>([&](auto kk) { constexpr int k1 = kk; v[k1] = tile[(k1 * M + m) * CB + b]; });
                static_for<1, R1>([&](auto kk) { constexpr int k1 = kk; v[k1] = cd_mulc(v[k1], tw1[k1 * M + m]); });
                dft<R1, +1>(v);
            }
            JB_PP_REL();
            // every shared-memory read of the slot is done: hand it back, thread 0 refills it two items ahead
            jb_fence_proxy_async();
            __syncwarp();
            if ((tid & 31) == 0) jb_mbar_arrive(&bars[2 + s]);
            if (tid == 0) {
                const int nxt = item_of(it + 2);
                if (nxt >= 0) {
                    jb_mbar_wait<false>(&bars[2 + s], (uint32_t)((it >> 1) & 1));
                    issue_loads(nxt, s);
                }
            }
            __syncwarp();
            if (act && b < ncols) {
                const int q0 = m - out_p0;            // output row of n1 = 0 (may be negative)
                const long long stride = (long long)M * PH;
                cd* ob = out + out_off + b + (long long)q0 * PH;
                static_for<0, R1>([&](auto nn) {
                    constexpr int n1 = nn;
                    const int q = q0 + M * n1;
                    if (q >= 0 && q < out_rows) ob[n1 * stride] = v[n1];
                });
            }
        }
        // (no group barrier needed here: the next round works on the other ring slot, and a slot is only refilled
        //  after every warp of the group has arrived on its "empty" barrier)
    }
    if (grp == 0) jb_nbar_sync(2, ALL);     // pairs with group 1's last hand-over: no barrier is left half-arrived
#undef JB_PP_ACQ
#undef JB_PP_REL
#undef JB_PP_GSYNC
}
#endif
