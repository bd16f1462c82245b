// cols3.cuh -- column pass with a THREE-stage transform L = R1*R2*R3 (e.g. 384 = 8*6*8) on a CB-column tile.
//
// Why three light stages instead of two heavy ones: an FP64 butterfly of radix 24 keeps ~100 registers
// live, so only 4-6 warps per SM could work at a time and the FP64 pipe (2 issue cycles per warp
// instruction, long dependent latency) sat at ~25 %.  Radix 6/8 items need ~40 registers, every stage has
// >= 384 independent items, and a 512-thread CTA keeps 4 warps per scheduler busy in every stage.
//
// Index maps (decimation in time), M = R2*R3:
//   n = M*n1 + R3*n2 + n3          k = k1 + R1*(k2 + R2*k3)
//   stage 1: DFT_R1 over n1, * W_L^{(R3 n2 + n3) k1}        rows {M*a + m}      in place
//   stage 2: DFT_R2 over n2, * W_M^{n3 k2}                   rows {k1*M + R3*a + n3} in place
//   stage 3: DFT_R3 over n3  -> X[k] at row k1*M + R3*k2 + k3; * K^ ; inverse DFT_R3 in the same registers
//   then stages 2', 1' mirror 2, 1 with conjugated twiddles applied BEFORE the inverse butterflies.
// K^ is stored per (module/energy, patch, strip) as [k3][k1*R2 + k2][b]: the register distribution of
// stage 3, so the slice is contiguous (one 1-D bulk copy) and read conflict-free.
#pragma once
#include "cols_pipe.cuh"

template <int L_, int R1_, int R2_, int R3_>
struct Fft3 {
    static constexpr int L = L_, R1 = R1_, R2 = R2_, R3 = R3_, M = R2_ * R3_;
    static_assert(R1_ * R2_ * R3_ == L_, "L = R1*R2*R3");
};

// threads per CTA = the widest stage (one work item per thread and stage), rounded up to whole warps; tiles of
// <= 256 threads run two CTAs per SM so that one CTA's barrier/latency stalls overlap the other's butterflies
JB_HD constexpr int jb_c3_max3(int a, int b, int c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }

JB_HD constexpr int jb_c3_gk(int per_k1) { int g = 1; while ((g * per_k1) % 32 != 0) ++g; return g; }   // k1 per warp-aligned group

#ifndef JB_MT_ISSUER
#define JB_MT_ISSUER 1        // ring refills issued by an otherwise idle warp instead of thread 0
#endif

// what the compute threads need to know about a work item; written by thread 0 when it stages the item
struct JbColInfo {
    long long out_off;  // element offset of (row 0, column strip*CB) of this patch in the output buffer
    int vlo, vhi;       // valid (in-image) input rows of the tile
    int out_p0, out_rows;
    int ncols;          // valid columns in this strip (PH - strip*CB, capped at CB)
    int row_base;       // sweep mode: image row of output row 0 of this patch (may be negative)
    long long sw_off;   // sweep mode: element offset of (slab, image row 0, column strip*CB) in the summed buffer
    int rmw_end;        // sweep mode: image rows < rmw_end were already written by the previous patch of the chain
    int pad;
    long long k_off;    // element offset of this item's K^ slice (direct-load mode)
};

// twiddle tables (host-built, uploaded once): tw1[k1*M + m] = W_L^{m k1}, tw2[k2*R3 + n3] = W_M^{n3 k2}
template <class F3, int CB>
struct Cols3 {
    static constexpr int L = F3::L, R1 = F3::R1, R2 = F3::R2, R3 = F3::R3, M = F3::M;
    static constexpr int I1 = M * CB, I2 = R1 * R3 * CB, I3 = R1 * R2 * CB;
    // WL ("warp-local middle"): the R2*R3-point sub-transform of one k1 (stage 2, stage 3 * K^ * stage 3', stage 2')
    // for all CB columns fits one warp, so warp k1 runs those three stages with warp-level syncs only and the
    // item needs two block barriers instead of five
    // Measured on B200 (r1): no gain at L = 384 (158 -> 158 us) and a loss at L = 336 (idle lanes), i.e. block
    // barriers are not what limits this kernel -- kept behind JB_COLS_WL for the record.
#ifdef JB_COLS_WL
    static constexpr bool WL = (R3 * CB <= 32) && (R2 * CB <= 32);
#else
    static constexpr bool WL = false;
#endif
    static constexpr int THREADS = WL ? (jb_c3_max3(I1, 32 * R1, 0) + 31) / 32 * 32 : (jb_c3_max3(I1, I2, I3) + 31) / 32 * 32;
    // CTAs per SM the pipeline is sized for: by threads (<= 512 resident) and by the ring's shared memory (2 tiles + 2 K^
    // slices + twiddles + slack)
    static constexpr int SMEM_EST = 4 * L * CB * 16 + (L + M) * 16 + 4096;
    static constexpr int CPS_T = THREADS <= 128 ? 4 : (THREADS <= 256 ? 2 : 1);
    static constexpr int CPS_S = (227 * 1024) / SMEM_EST < 1 ? 1 : (227 * 1024) / SMEM_EST;
    static constexpr int CPS = CPS_T < CPS_S ? CPS_T : CPS_S;
    static constexpr int CPS_KD = THREADS <= 256 ? 3 : 1; // ... when K^ is read straight from L2 (no staging ring)
    static_assert(THREADS <= 1024, "one item per thread and stage");
    // Stage 2 <-> stage 3 only exchange data within one k1: the threads of GK consecutive k1 (GT = GK*R2*CB threads, whole
    // warps) form a GROUP that owns these sub-transforms in both stages, so that a named barrier over GT threads replaces
    // the block barrier between them (module-sharing kernels, JB_MT_GROUP)
    static constexpr int GK = jb_c3_gk(R2 * CB), GT = GK * R2 * CB, NG = R1 / GK, GI2 = GK * R3 * CB;   // GI2: stage-2 items of a group
    static constexpr bool GROUPS = (R1 % GK == 0) && NG > 1 && NG <= 15;
    // SWZ: with 4-column tiles two stage-3 items of one quarter-warp (rows 8c + n3 and 8c + 8 + n3) hit the same
    // banks.  Between the first forward stage and the last inverse stage the tile is therefore kept row-swizzled,
    // physical row = row ^ ((row >> 3) & 1): neighbouring 8-row groups swap their even/odd rows, every stage stays
    // on the same pair of rows per lane pair, and stage 3 becomes conflict-free (ncu r1h: 19 % of the column pass's
    // shared-memory wavefronts were replays).  Row parities follow from M % 16 == 0 and R3 == 8.
    // Measured on B200: conflict-free, yet SLOWER (158 -> 165 us): the extra address arithmetic costs more than the
    // replays -- the kernel is bound by per-warp instruction latency, not by shared-memory bandwidth.  Opt-in.
#ifdef JB_COLS_SWZ
    static constexpr bool SWZ = !WL && (R3 == 8) && (CB == 4) && (M % 16 == 0);
#else
    static constexpr bool SWZ = false;
#endif
    static JB_HD int swz_off(int par, int odd) { return SWZ ? par * (1 - 2 * odd) * CB : 0; }   // element offset of the swap

    // forward stage 1 with zero rows outside [vlo, vhi); rw != nullptr: multiply tile row p by rw[p - p0]
    // (the window row weight of the forward direction, convolve.py:88-90)
    static JB_D void s1_fwd(cd* tile, const cd* tw1, int w, int vlo, int vhi, const double* rw, int p0) {
        const bool act = w < I1;
        const int m = w / CB, b = w - m * CB;
        cd v[R1];
        if (act) {
            static_for<0, R1>([&](auto nn) {
                constexpr int n1 = nn;
                int p = M * n1 + m;
                cd t = cd_make(0.0, 0.0);
                if (p >= vlo && p < vhi) {
                    t = tile[p * CB + b];
                    if (rw) t = cd_scale(t, rw[p - p0]);
                }
                v[n1] = t;
            });
            dft<R1, -1>(v);
        }
        if (SWZ) JB_SYNCWARP();   // the swapped destination row is an input row of a neighbouring lane of this warp
        if (act) {
            const int so = swz_off((m >> 3) & 1, m & 1);
            static_for<0, R1>([&](auto kk) {
                constexpr int k1 = kk;
                cd y = v[k1];
                if constexpr (k1 > 0) y = cd_mul(y, tw1[k1 * M + m]);
                tile[(k1 * M + m) * CB + b + so] = y;
            });
        }
    }
    static JB_D void s2_fwd(cd* tile, const cd* tw2, int w) {
        if (w >= I2) return;
        const int c = w / CB, b = w - c * CB;   // c = k1*R3 + n3
        const int k1 = c / R3, n3 = c - k1 * R3;
        cd* base = tile + (size_t)(k1 * M + n3) * CB + b;
        const int p0 = ((k1 * M + n3) >> 3) & 1;
        const int so[2] = {swz_off(p0, n3 & 1), swz_off(1 - p0, n3 & 1)};   // even / odd n2 (R3 == 8 rows per step)
        cd v[R2];
        static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; v[n2] = base[(R3 * n2) * CB + so[n2 & 1]]; });
        dft<R2, -1>(v);
        static_for<0, R2>([&](auto kk) {
            constexpr int k2 = kk;
            cd y = v[k2];
            if constexpr (k2 > 0) y = cd_mul(y, tw2[k2 * R3 + n3]);
            base[(R3 * k2) * CB + so[k2 & 1]] = y;
        });
    }
    // stage 3 forward only (plan time): returns the spectrum values of this thread
    static JB_D void s3_fwd(const cd* tile, int w, cd (&v)[R3]) {
        const int c = w / CB, b = w - c * CB;   // c = k1*R2 + k2 ; rows R3*c + n3
        const cd* base = tile + (size_t)(R3 * c) * CB + b;
        const int par = c & 1;
        static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; v[n3] = base[n3 * CB + swz_off(par, n3 & 1)]; });
        dft<R3, -1>(v);
    }
    // stage 3: forward, * K^ (conj for the adjoint), inverse -- the spectrum never leaves the registers
    // KD: kh is the slice in GLOBAL memory (L2-warm thanks to the producer's prefetch): the loads are issued first
    // and land while the forward butterfly runs
    template <bool ADJ, bool KD = false>
    static JB_D void s3_mid(cd* tile, const cd* kh, int w) {
        if (w >= I3) return;
        const int c = w / CB, b = w - c * CB;
        cd* base = tile + (size_t)(R3 * c) * CB + b;
        const int par = c & 1;
        cd v[R3], kq[KD ? R3 : 1];
        if (KD) static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; kq[KD ? k3 : 0] = jb_ld_nc(kh + k3 * I3 + w); });
        static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; v[n3] = base[n3 * CB + swz_off(par, n3 & 1)]; });
        dft<R3, -1>(v);
        static_for<0, R3>([&](auto kk) {
            constexpr int k3 = kk;
            cd kv = KD ? kq[KD ? k3 : 0] : kh[k3 * I3 + w];
            v[k3] = ADJ ? cd_mulc(v[k3], kv) : cd_mul(v[k3], kv);
        });
        dft<R3, +1>(v);
        static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; base[n3 * CB + swz_off(par, n3 & 1)] = v[n3]; });
    }
    static JB_D void s2_inv(cd* tile, const cd* tw2, int w) {
        if (w >= I2) return;
        const int c = w / CB, b = w - c * CB;
        const int k1 = c / R3, n3 = c - k1 * R3;
        cd* base = tile + (size_t)(k1 * M + n3) * CB + b;
        const int p0 = ((k1 * M + n3) >> 3) & 1;
        const int so[2] = {swz_off(p0, n3 & 1), swz_off(1 - p0, n3 & 1)};
        cd v[R2];
        static_for<0, R2>([&](auto kk) {
            constexpr int k2 = kk;
            cd y = base[(R3 * k2) * CB + so[k2 & 1]];
            if constexpr (k2 > 0) y = cd_mulc(y, tw2[k2 * R3 + n3]);
            v[k2] = y;
        });
        dft<R2, +1>(v);
        static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; base[(R3 * n2) * CB + so[n2 & 1]] = v[n2]; });
    }
    // ---- warp-local variants: warp k1, lane = (n3 or k2)*CB + b
    static JB_D void s2_fwd_w(cd* tile, const cd* tw2, int k1, int lane) {
        if (lane >= R3 * CB) return;
        const int n3 = lane / CB, b = lane - n3 * CB;
        cd* base = tile + (size_t)(k1 * M + n3) * CB + b;
        cd v[R2];
        static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; v[n2] = base[(R3 * n2) * CB]; });
        dft<R2, -1>(v);
        static_for<0, R2>([&](auto kk) {
            constexpr int k2 = kk;
            cd y = v[k2];
            if constexpr (k2 > 0) y = cd_mul(y, tw2[k2 * R3 + n3]);
            base[(R3 * k2) * CB] = y;
        });
    }
    template <bool ADJ, bool KD = false>
    static JB_D void s3_mid_w(cd* tile, const cd* kh, int k1, int lane) {
        if (lane >= R2 * CB) return;
        const int w = k1 * (R2 * CB) + lane;          // = (k1*R2 + k2)*CB + b : the K^ slice index of s3_mid
        const int c = w / CB, b = w - c * CB;
        cd* base = tile + (size_t)(R3 * c) * CB + b;
        cd v[R3], kq[KD ? R3 : 1];
        if (KD) static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; kq[KD ? k3 : 0] = jb_ld_nc(kh + k3 * I3 + w); });
        static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; v[n3] = base[n3 * CB]; });
        dft<R3, -1>(v);
        static_for<0, R3>([&](auto kk) {
            constexpr int k3 = kk;
            cd kv = KD ? kq[KD ? k3 : 0] : kh[k3 * I3 + w];
            v[k3] = ADJ ? cd_mulc(v[k3], kv) : cd_mul(v[k3], kv);
        });
        dft<R3, +1>(v);
        static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; base[n3 * CB] = v[n3]; });
    }
    static JB_D void s2_inv_w(cd* tile, const cd* tw2, int k1, int lane) {
        if (lane >= R3 * CB) return;
        const int n3 = lane / CB, b = lane - n3 * CB;
        cd* base = tile + (size_t)(k1 * M + n3) * CB + b;
        cd v[R2];
        static_for<0, R2>([&](auto kk) {
            constexpr int k2 = kk;
            cd y = base[(R3 * k2) * CB];
            if constexpr (k2 > 0) y = cd_mulc(y, tw2[k2 * R3 + n3]);
            v[k2] = y;
        });
        dft<R2, +1>(v);
        static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; base[(R3 * n2) * CB] = v[n2]; });
    }
    // inverse stage 1: load + conj twiddle (registers only); the caller may release the tile afterwards
    static JB_D void s1_inv_load(const cd* tile, const cd* tw1, int w, cd (&v)[R1]) {
        const int m = w / CB, b = w - m * CB;
        const int so = swz_off((m >> 3) & 1, m & 1);
        static_for<0, R1>([&](auto kk) {
            constexpr int k1 = kk;
            cd y = tile[(k1 * M + m) * CB + b + so];
            if constexpr (k1 > 0) y = cd_mulc(y, tw1[k1 * M + m]);
            v[k1] = y;
        });
    }
};

// --------------------------------------------------------------------------------------------------------
// Persistent TMA pipeline, 512 threads, 2-stage ring of (tile, K^ slice); thread 0 refills a slot as soon as
// every warp has pulled its stage-1' operands out of it.  dynamic smem = 4*L*CB*16 + (L + M)*16 + 64.
// SWEEP: a CTA walks whole column chains (plane, column block i, strip) through the row-patches j = 0..n-1 and
// ACCUMULATES the patch outputs that land on the same image row into one summed buffer (planes*n, N, PH) -- plain
// store on first touch, read-add-write (L2-resident, written by this CTA one item earlier) otherwise; fixed order
// j ascending, so the sums are reproducible.  The adjoint applies the window row weight before accumulating.  The
// row pass then reads ONE spectrum row per (image row, i) instead of one per covering patch.
// KD: no K^ staging ring -- thread 0 prefetches the slice into L2 two items ahead and the stage-3 threads load it
// straight into registers; the shared memory saved lets a third CTA (24 warps) share the SM.
template <class F3, int CB, bool ADJ>
JB_GLOBAL void JB_LB((Cols3<F3, CB>::THREADS), (Cols3<F3, CB>::CPS))
k_cols3_pipe(const
#ifndef JB_EMU
             __grid_constant__
#endif
             JbTileSrc src, JbGeom g, cd* __restrict__ out, const cd* __restrict__ khat,
             const cd* __restrict__ tw1g, const cd* __restrict__ tw2g, const double* __restrict__ wgt, int planes) {
    using C3 = Cols3<F3, CB>;
    constexpr int L = F3::L, R1 = F3::R1, M = F3::M;
    constexpr int TILE = L * CB;
    constexpr int JB_C3_THREADS = C3::THREADS;
    constexpr bool SWEEP = false, KD = false;   // round-1 experiments (column-chain accumulation, K^ straight from L2): see DESIGN.md
    JB_SMEM(unsigned char, smraw);
    // The TMA destination (tile row in_p0 + chunk start) must be 128-byte aligned.  Chunk starts are even, in_p0 is the
    // margin in the forward direction (0 in the adjoint): with CB*16 = 64-byte rows an ODD margin (e.g. ceil(0.062*2048)
    // = 127) would put every box on a 64-byte boundary, so the whole tile region is shifted by the missing rows
    // (compute-sanitizer memcheck, r2c: "misaligned shared address" at m = 17).
    constexpr int ROWS_PER_128 = (128 / (CB * (int)sizeof(cd))) > 0 ? (128 / (CB * (int)sizeof(cd))) : 1;
    const int tshift = ADJ ? 0 : (ROWS_PER_128 - g.m % ROWS_PER_128) % ROWS_PER_128;
    cd* tile0 = reinterpret_cast<cd*>(smraw) + tshift * CB;
    cd* kh0 = tile0 + 2 * TILE;
    cd* tw1 = kh0 + (KD ? 0 : 2 * TILE);
    cd* tw2 = tw1 + L;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tw2 + M);
    JbColInfo* info = reinterpret_cast<JbColInfo*>(bars + 4);
    double* rw = reinterpret_cast<double*>(info + 2);  // window row weights (forward direction only)
    const int nstrips = (g.PH + CB - 1) / CB;
    const long nitems = (long)planes * g.n * g.n * nstrips;
    const long nchains = (long)planes * g.n * nstrips;
    const int tid = threadIdx.x;
    const int PH = g.PH;
    // local iteration -> global item ((plane*n*n + i*n + j)*nstrips + strip), -1 past the end
    auto item_of = [&](long it_) -> long {
        if (!SWEEP) {
            long v = (long)blockIdx.x + it_ * (long)gridDim.x;
            return v < nitems ? v : -1;
        }
        const long c = (long)blockIdx.x + (it_ / g.n) * (long)gridDim.x;
        if (c >= nchains) return -1;
        const int j = (int)(it_ % g.n), strip = (int)(c % nstrips);
        const long r = c / nstrips;
        const int i = (int)(r % g.n);
        const long pl = r / g.n;
        return ((pl * g.n + i) * g.n + j) * nstrips + strip;
    };
    auto fill_info = [&](const JbColItem& ci, int s_) {
        JbColInfo& f = info[s_];
        f.out_off = (long long)ci.out_off + (long long)ci.strip * CB;
        f.vlo = ci.vlo; f.vhi = ci.vhi; f.out_p0 = ci.out_p0; f.out_rows = ci.out_rows;
        int nc = PH - ci.strip * CB;
        f.ncols = nc < CB ? nc : CB;
        f.k_off = (long long)ci.k_off;
        if (SWEEP) {
            const int back = ADJ ? g.hp : g.hp + g.m;    // image row of output row 0 = offs[j] - back
            f.row_base = g.offs[ci.j] - back;
            f.sw_off = ((long long)(ci.plane * g.n + ci.i) * g.N) * PH + (long long)ci.strip * CB;
            f.rmw_end = ci.j > 0 ? g.offs[ci.j - 1] - back + ci.out_rows : -(1 << 30);
        }
    };
    for (int i = tid; i < L; i += JB_C3_THREADS) tw1[i] = tw1g[i];
    for (int i = tid; i < M; i += JB_C3_THREADS) tw2[i] = tw2g[i];
    if (!ADJ || SWEEP) for (int i = tid; i < g.w; i += JB_C3_THREADS) rw[i] = wgt[i];
#ifndef JB_EMU
    auto issue_loads = [&](long item_, int s_) {
        JbColItem ci = jb_col_item<L, CB, ADJ>(g, item_, nstrips);
        fill_info(ci, s_);  // ordered before the consumers' reads by the release/acquire of full[s_]
        cd* tile = tile0 + s_ * TILE;
        const int BR = src.BR;
        int nch = (ci.in_rows + BR - 1) / BR, issued = 0;
        for (int c = 0; c < nch; ++c) {
            int r0 = c * BR;
            if (r0 + BR > ci.in_rows) r0 = ci.in_rows - BR;
            int p0 = ci.in_p0 + r0;
            if (p0 + BR > ci.vlo && p0 < ci.vhi) ++issued;
        }
#ifdef JB_ABL_NOTILE
        issued = 0;
#endif
#ifdef JB_ABL_NOKHAT
        jb_mbar_expect_tx(&bars[s_], (uint32_t)(issued * BR * CB * sizeof(cd)));
#else
        jb_mbar_expect_tx(&bars[s_], (uint32_t)(issued * BR * CB * sizeof(cd) + (KD ? 0 : TILE * sizeof(cd))));
#endif
        for (int c = 0; c < nch; ++c) {
            int r0 = c * BR;
            if (r0 + BR > ci.in_rows) r0 = ci.in_rows - BR;
            int p0 = ci.in_p0 + r0;
#ifndef JB_ABL_NOTILE
            if (p0 + BR > ci.vlo && p0 < ci.vhi)
                jb_tma_load_3d(tile + (size_t)p0 * CB, &src.map, 2 * ci.strip * CB, ci.row0 + r0, ci.in_slab, &bars[s_]);
#endif
        }
#ifndef JB_ABL_NOKHAT
        if (KD) jb_bulk_prefetch_l2(khat + ci.k_off, (uint32_t)(TILE * sizeof(cd)));
        else jb_bulk_load_1d(kh0 + s_ * TILE, khat + ci.k_off, (uint32_t)(TILE * sizeof(cd)), &bars[s_]);
#endif
    };
    if (tid == 0) {
        jb_mbar_init(&bars[0], 1);
        jb_mbar_init(&bars[1], 1);
        jb_mbar_init(&bars[2], JB_C3_THREADS / 32);
        jb_mbar_init(&bars[3], JB_C3_THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (item_of(0) >= 0) issue_loads(item_of(0), 0);
        if (item_of(1) >= 0) issue_loads(item_of(1), 1);
    }
#endif
    JB_SYNC();
    for (unsigned it = 0;; ++it) {
        const long item = item_of(it);
        if (item < 0) break;
        const int s = it & 1;
        cd* tile = tile0 + s * TILE;
        const cd* kh = KD ? nullptr : kh0 + s * TILE;
#ifndef JB_EMU
        jb_mbar_wait<false>(&bars[s], (it >> 1) & 1);
#else
        JbColItem ci = jb_col_item<L, CB, ADJ>(g, item, nstrips);
        JB_SYNC();
        if (tid == 0) fill_info(ci, s);
        for (int idx = tid; idx < ci.in_rows * CB; idx += JB_C3_THREADS) {
            int r = idx / CB, b = idx - r * CB, col = ci.strip * CB + b, ir = ci.row0 + r;
            tile[(size_t)(ci.in_p0 + r) * CB + b] = (col < PH && ir >= 0 && ir < src.N)
                ? src.base[((size_t)ci.in_slab * src.N + ir) * PH + col] : cd_make(0.0, 0.0);
        }
        if (!KD) for (int idx = tid; idx < TILE; idx += JB_C3_THREADS) kh0[s * TILE + idx] = khat[ci.k_off + idx];
        JB_SYNC();
#endif
        const JbColInfo f = info[s];
        if (KD) kh = khat + f.k_off;
#ifndef JB_ABL_NOCOMPUTE
        C3::s1_fwd(tile, tw1, tid, f.vlo, f.vhi, ADJ ? nullptr : rw, g.m);
        JB_SYNC();
        if constexpr (C3::WL) {
            const int k1 = tid >> 5, lane = tid & 31;
            if (k1 < F3::R1) C3::s2_fwd_w(tile, tw2, k1, lane);
            JB_SYNCWARP();
            if (k1 < F3::R1) C3::template s3_mid_w<ADJ, KD>(tile, kh, k1, lane);
            JB_SYNCWARP();
            if (k1 < F3::R1) C3::s2_inv_w(tile, tw2, k1, lane);
            JB_SYNC();
        } else {
            C3::s2_fwd(tile, tw2, tid);
            JB_SYNC();
            C3::template s3_mid<ADJ, KD>(tile, kh, tid);
            JB_SYNC();
            C3::s2_inv(tile, tw2, tid);
            JB_SYNC();
        }
#endif
        {
            const bool act = tid < C3::I1;
            cd v[R1];
            if (act) C3::s1_inv_load(tile, tw1, tid, v);
#ifndef JB_EMU
            jb_fence_proxy_async();
            __syncwarp();
            if ((tid & 31) == 0) jb_mbar_arrive(&bars[2 + s]);
            // the refill is issued by the first lane of a warp WITHOUT a stage-1' item where there is one (I1 < THREADS, e.g.
            // L = 384: warps 6 and 7): waiting for the "consumed" barrier and feeding the copy engine then delays no butterfly
            constexpr int PISSUER = JB_MT_ISSUER && ((C3::I1 + 31) / 32 * 32 < JB_C3_THREADS) ? (C3::I1 + 31) / 32 * 32 : 0;
            if (tid == PISSUER) {
                // (refilling behind the next item's first block barrier instead -- what k_cols3_fwd_mt does -- costs this
                // kernel its prefetch lead: 300 -> 326 us at C2; refilling after thread 0's own stores: 271 -> 281 us, r2u)
                long nxt = item_of((long)it + 2);
                if (nxt >= 0) {
                    jb_mbar_wait<false>(&bars[2 + s], (it >> 1) & 1);
                    issue_loads(nxt, s);
                }
            }
#endif
            if (act) {
                const int m = tid / CB, b = tid - m * CB;
                const int q0 = m - f.out_p0;            // output row of n1 = 0 (may be negative)
                if (SWEEP) {
                    cd* ob = out + f.sw_off + b;
                    cd old[R1];
                    // previous partial sums first (L2 hits), so that their latency hides behind the last butterfly
                    static_for<0, R1>([&](auto nn) {
                        constexpr int n1 = nn;
                        const int q = q0 + M * n1, r = f.row_base + q;
                        old[n1] = cd_make(0.0, 0.0);
                        if (b < f.ncols && q >= 0 && q < f.out_rows && r >= 0 && r < g.N && r < f.rmw_end)
                            old[n1] = jb_ld_cg(ob + (long long)r * PH);
                    });
                    dft<R1, +1>(v);
                    static_for<0, R1>([&](auto nn) {
                        constexpr int n1 = nn;
                        const int q = q0 + M * n1, r = f.row_base + q;
                        if (b < f.ncols && q >= 0 && q < f.out_rows && r >= 0 && r < g.N) {
                            cd y = v[n1];
                            if (ADJ) y = cd_scale(y, rw[q]);
                            ob[(long long)r * PH] = cd_add(old[n1], y);
                        }
                    });
                } else {
#ifndef JB_ABL_NOCOMPUTE
                    dft<R1, +1>(v);
#endif
                    const long long stride = (long long)M * PH;
                    cd* ob = out + f.out_off + b + (long long)q0 * PH;
#ifdef JB_ABL_NOSTORE
                    if (b < f.ncols && v[0].x == 1.2345e300) {
#else
                    if (b < f.ncols) {
#endif
                        static_for<0, R1>([&](auto nn) {
                            constexpr int n1 = nn;
                            int q = q0 + M * n1;
                            if (q >= 0 && q < f.out_rows) ob[n1 * stride] = v[n1];
                        });
                    }
                }
            }
        }
        // (the stage barriers of the next item order this item's stores before the next read-add-write)
    }
}

// --------------------------------------------------------------------------------------------------------
// Forward column pass for T > 1 modules: the forward column transform of a windowed sky patch does not depend on the
// telescope module -- only K^ does.  One work item = (sample, energy bin, patch, strip); the tile is loaded and
// transformed ONCE (stages 1, 2 and the forward half of stage 3); the spectrum stays in the stage-3 threads' REGISTERS
// (R3 <= JB_ADJ_ACC_REG; otherwise it is written back to tile A), then every module t multiplies it by its K^ slice and
// runs the inverse half (stage 3', 2', 1') on tile A / tile B alternately (tile B only, when the spectrum lives in tile A).
// Per module this is 4/7 of the FP64 work and 6.9 (7.9) instead of 11.3 shared-memory passes of the one-module kernel
// (T = 7: BASELINE configs[2] and [4]).  Shared memory: tile A, tile B, a two-slot ring of K^ slices -- the same
// 4 * L * CB * 16 bytes.  Refills are issued by a warp that is idle in the following stage (ISSUER).
#ifndef JB_ADJ_ACC_REG
#define JB_ADJ_ACC_REG 7   // largest last-stage radix R3 whose per-thread spectrum / module accumulator (R3 complex values per stage-3
                           // thread) is kept in registers by the *_mt kernels; above it the forward kernel keeps the spectrum in tile A
                           // and the adjoint's accumulator lives in a per-CTA global scratch tile that stays in L2.
                           // r2w, C5 chunk: R3 = 12 in registers spills (cols_adj 12.15 -> 12.88 ms), R3 = 7 does not (9.79 ms)
#endif
#ifndef JB_FWD_SPEC_REG
#define JB_FWD_SPEC_REG 1
#endif
#ifndef JB_ADJ_DEFER
#define JB_ADJ_DEFER 1
#endif
// Measured and removed (r2z): the sample index as the fastest item index (neighbouring CTAs share K^ slices through L2:
// cols_fwd 7.32 -> 7.27 ms, cols_adj 8.94 -> 10.20 ms) and an L2 prefetch of the K^ slices 2 / 4 ring steps ahead of their bulk copy
// (7.30 -> 8.29 ms, 8.94 -> 10.10 ms).
#ifndef JB_MT_GROUP_FWD_MINCPS
#define JB_MT_GROUP_FWD_MINCPS 2   // forward kernel: group barriers only from this many CTAs per SM on (see the kernel)
#endif
#ifndef JB_MT_GROUP
#define JB_MT_GROUP 3      // stage 2 <-> stage 3 of the module loops synchronise per k1 group (named barriers) instead of per CTA:
                           // bit 0 = adjoint kernel, bit 1 = forward kernel
#endif
template <class F3, int CB>
JB_GLOBAL void JB_LB((Cols3<F3, CB>::THREADS), (Cols3<F3, CB>::CPS))
k_cols3_fwd_mt(const
#ifndef JB_EMU
               __grid_constant__
#endif
               JbTileSrc src, JbGeom g, cd* __restrict__ out, const cd* __restrict__ khat,
               const cd* __restrict__ tw1g, const cd* __restrict__ tw2g, const double* __restrict__ wgt, int planes_se) {
    using C3 = Cols3<F3, CB>;
    constexpr int L = F3::L, R1 = F3::R1, R3 = F3::R3, M = F3::M;
    constexpr int TILE = L * CB;
    constexpr int NT = C3::THREADS;
    JB_SMEM(unsigned char, smraw);
    constexpr int ROWS_PER_128 = (128 / (CB * (int)sizeof(cd))) > 0 ? (128 / (CB * (int)sizeof(cd))) : 1;
    const int tshift = (ROWS_PER_128 - g.m % ROWS_PER_128) % ROWS_PER_128;   // odd margins: see k_cols3_pipe
    cd* tileA = reinterpret_cast<cd*>(smraw) + tshift * CB;
    cd* tileB = tileA + TILE;
    cd* kh0 = tileB + TILE;
    cd* tw1 = kh0 + 2 * TILE;
    cd* tw2 = tw1 + L;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tw2 + M);     // [0] tile A full, [1..2] K^ slot full (the slots are handed back
                                                               // behind a block barrier: no "consumed" barriers needed)
    double* rw = reinterpret_cast<double*>(bars + 6);
    const int tid = threadIdx.x;
    const int n = g.n, N = g.N, PH = g.PH, nn2 = n * n, T = g.T;
    const int nstrips = (PH + CB - 1) / CB;
    const long nitems = (long)planes_se * nn2 * nstrips;
    for (int i = tid; i < L; i += NT) tw1[i] = tw1g[i];
    for (int i = tid; i < M; i += NT) tw2[i] = tw2g[i];
    for (int i = tid; i < g.w; i += NT) rw[i] = wgt[i];
    struct Item { int strip, k, s, e, j, i, vlo, vhi, row0, ncols; };
    auto decode = [&](long item) -> Item {
        Item q;
        q.strip = (int)(item % nstrips);
        long r = item / nstrips;
        q.k = (int)(r % nn2);
        const int lp = (int)(r / nn2);
        q.s = lp / g.Es; q.e = g.e0 + (lp - q.s * g.Es);
        q.j = q.k % n; q.i = q.k / n;
        q.row0 = g.offs[q.j] - g.hp;                       // image row of tile row m
        const int lo = -q.row0, hi = N - q.row0;
        q.vlo = g.m + (lo > 0 ? lo : 0);
        q.vhi = g.m + (hi < g.w ? hi : g.w);
        const int nc = PH - q.strip * CB;
        q.ncols = nc < CB ? nc : CB;
        return q;
    };
    auto item_of = [&](long it_) -> long {
        long v = (long)blockIdx.x + it_ * (long)gridDim.x;
        return v < nitems ? v : -1;
    };
    auto k_off = [&](const Item& q, int t) -> size_t {
        return ((((size_t)t * g.E + q.e) * nn2 + q.k) * nstrips + q.strip) * ((size_t)L * CB);
    };
    // my item count (uniform per CTA) and the flattened (item, module) step sequence for the K^ ring
    long my_items = 0;
    if ((long)blockIdx.x < nitems) my_items = (nitems - 1 - (long)blockIdx.x) / (long)gridDim.x + 1;
    const long nsteps = my_items * T;
#ifndef JB_EMU
    auto issue_tile = [&](long item) {
        const Item q = decode(item);
        const int BR = src.BR;
        const int in_rows = g.w, in_p0 = g.m;
        int nch = (in_rows + BR - 1) / BR, issued = 0;
        for (int c = 0; c < nch; ++c) {
            int r0 = c * BR;
            if (r0 + BR > in_rows) r0 = in_rows - BR;
            int p0 = in_p0 + r0;
            if (p0 + BR > q.vlo && p0 < q.vhi) ++issued;
        }
        jb_mbar_expect_tx(&bars[0], (uint32_t)(issued * BR * CB * sizeof(cd)));
        for (int c = 0; c < nch; ++c) {
            int r0 = c * BR;
            if (r0 + BR > in_rows) r0 = in_rows - BR;
            int p0 = in_p0 + r0;
            if (p0 + BR > q.vlo && p0 < q.vhi)
                jb_tma_load_3d(tileA + (size_t)p0 * CB, &src.map, 2 * q.strip * CB, q.row0 + r0, (q.s * g.E + q.e) * n + q.i, &bars[0]);
        }
    };
    auto issue_k = [&](long step) {           // K^ slice of flattened step `step` -> ring slot step & 1
        const Item q = decode(item_of(step / T));
        const int ks = (int)(step & 1);
        jb_mbar_expect_tx(&bars[1 + ks], (uint32_t)(TILE * sizeof(cd)));
        jb_bulk_load_1d(kh0 + ks * TILE, khat + k_off(q, (int)(step % T)), (uint32_t)(TILE * sizeof(cd)), &bars[1 + ks]);
    };
    if (tid == 0) {
        jb_mbar_init(&bars[0], 1);
        jb_mbar_init(&bars[1], 1);
        jb_mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (my_items > 0) issue_tile(item_of(0));
        if (nsteps > 0) issue_k(0);
        if (nsteps > 1) issue_k(1);
    }
#endif
    JB_SYNC();
    for (long it = 0; it < my_items; ++it) {
        const Item q = decode(item_of(it));
#ifndef JB_EMU
        jb_mbar_wait<false>(&bars[0], (uint32_t)(it & 1));
#else
        JB_SYNC();
        for (int idx = tid; idx < g.w * CB; idx += NT) {
            int r = idx / CB, b = idx - r * CB, col = q.strip * CB + b, ir = q.row0 + r;
            tileA[(size_t)(g.m + r) * CB + b] = (col < PH && ir >= 0 && ir < src.N)
                ? src.base[((size_t)((q.s * g.E + q.e) * n + q.i) * src.N + ir) * PH + col] : cd_make(0.0, 0.0);
        }
        JB_SYNC();
#endif
        // ---- shared by all modules: stages 1, 2 and the forward half of stage 3
        C3::s1_fwd(tileA, tw1, tid, q.vlo, q.vhi, rw, g.m);
        JB_SYNC();
        C3::s2_fwd(tileA, tw2, tid);
        JB_SYNC();
        // SPEC_REG: the R3 spectrum values of a stage-3 thread stay in REGISTERS over the module loop (one shared-memory read
        // pass less per module) and tile A, free from here on, alternates with tile B as the inverse transform's tile: the
        // barrier that protected the single tile B against the next module's stage 3' goes away (2 block barriers per module
        // instead of 3).  The last module always works in tile B, so the next item's TMA load into tile A is issued where it was.
        constexpr bool SPEC_REG = JB_FWD_SPEC_REG && R3 <= JB_ADJ_ACC_REG;
        // (forward: only where a second CTA shares the SM -- the K^ slot is handed back one stage later in this mode, and at
        //  one CTA per SM the shorter prefetch lead costs more than the barrier saves: L = 672 cols_fwd 7.30 -> 7.39 ms,
        //  L = 336 0.840 -> 0.823 ms, r2z; again with the producer warp, r2G: L = 672 6.80 -> 7.63 ms)
        constexpr bool GROUP = (JB_MT_GROUP & 2) && SPEC_REG && C3::GROUPS && C3::CPS >= JB_MT_GROUP_FWD_MINCPS;
        // the ring refills are issued by the first lane of a warp that has NO item in the stage that follows the hand-back
        // (stage 2' with I2 items, stage 1' with I1 in group mode), so that the copy-engine commands stay off the critical
        // path of the warps the next barrier waits for; thread 0 when every warp is busy there
        constexpr int NEXT_ITEMS = GROUP ? C3::I1 : C3::I2;
        constexpr int ISSUER = JB_MT_ISSUER && ((NEXT_ITEMS + 31) / 32 * 32 < NT) ? (NEXT_ITEMS + 31) / 32 * 32 : 0;
        cd spec[R3];
        if (tid < C3::I3) {
            C3::s3_fwd(tileA, tid, spec);
            if (!SPEC_REG) {
                const int c = tid / CB, b = tid - c * CB;
                cd* base = tileA + (size_t)(R3 * c) * CB + b;
                static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; base[k3 * CB] = spec[k3]; });
            }
        }
        JB_SYNC();
        for (int t = 0; t < T; ++t) {
            const long step = it * T + t;
            const int ks = (int)(step & 1);
            const cd* kh = kh0 + ks * TILE;
            cd* tb = (SPEC_REG && ((T - 1 - t) & 1)) ? tileA : tileB;
#ifndef JB_EMU
            jb_mbar_wait<false>(&bars[1 + ks], (uint32_t)((step >> 1) & 1));
#else
            JB_SYNC();
            for (int idx = tid; idx < TILE; idx += NT) kh0[ks * TILE + idx] = khat[k_off(q, t) + idx];
            JB_SYNC();
#endif
            // ---- stage 3': spectrum * K^_t, inverse radix R3 -> this module's tile
            if (tid < C3::I3) {
                const int c = tid / CB, b = tid - c * CB;
                const cd* sa = tileA + (size_t)(R3 * c) * CB + b;
                cd* sb = tb + (size_t)(R3 * c) * CB + b;
                cd v[R3];
                static_for<0, R3>([&](auto kk) {
                    constexpr int k3 = kk;
                    v[k3] = cd_mul(SPEC_REG ? spec[k3] : sa[k3 * CB], kh[k3 * C3::I3 + tid]);
                });
                dft<R3, +1>(v);
                static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; sb[n3 * CB] = v[n3]; });
            }
#ifndef JB_EMU
            jb_fence_proxy_async();          // the K^ slot (and, after the last module, tile A) is handed back to the copy engine
#endif
            if (GROUP) {
                // stage 2' of a k1 only reads what the same group's stage 3' wrote: group barrier, then ONE block barrier
                // per module (behind stage 2'), which is also where the K^ slot / tile A are handed back
                const int grp = tid / C3::GT, gr = tid - grp * C3::GT;
                JB_GROUP_SYNC(tid < C3::I3, C3::NG, grp, C3::GT);
                if (tid < C3::I3)
                    for (int idx = gr; idx < C3::GI2; idx += C3::GT) C3::s2_inv(tb, tw2, grp * C3::GI2 + idx);
                JB_SYNC();
            } else {
                JB_SYNC();                    // tile complete; every warp is done with this K^ slot (and with tile A when t == T-1)
            }
#ifndef JB_EMU
            if (tid == ISSUER) {
                if (step + 2 < nsteps) issue_k(step + 2);
                if (t == T - 1 && it + 1 < my_items) issue_tile(item_of(it + 1));
            }
#endif
            if (!GROUP) {
                C3::s2_inv(tb, tw2, tid);
                JB_SYNC();
            }
            {
                const bool act = tid < C3::I1;
                cd v[R1];
                if (act) C3::s1_inv_load(tb, tw1, tid, v);
                if (!SPEC_REG) JB_SYNC();     // the single tile B may be overwritten by the next module's stage 3'
                if (act) {
                    dft<R1, +1>(v);
                    const int m_ = tid / CB, b = tid - m_ * CB;
                    if (b < q.ncols) {
                        const size_t plane = ((size_t)q.s * T + t) * g.E + q.e;
                        cd* ob = out + ((plane * nn2 + q.k) * (size_t)g.P) * PH + (size_t)q.strip * CB + b;
                        static_for<0, R1>([&](auto nn) {
                            constexpr int n1 = nn;
                            const int qr = M * n1 + m_;
                            if (qr < g.P) ob[(size_t)qr * PH] = v[n1];
                        });
                    }
                }
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// Adjoint column pass for T > 1 modules: the inverse column transform is linear, so the modules' products
// FFT(cotangent_t) * conj K^_t are summed in the spectral domain and ONE inverse transform per (sample, energy bin,
// patch, strip) writes the module-summed rows -- out: (S*E, n*n, w, PH); the c2r row pass behind it reads one spectrum
// instead of T.  Per module 4/7 of the FP64 work.  Shared memory keeps BOTH rings double-buffered (two input tiles, two
// K^ slices); the accumulator is R3 complex values per stage-3 thread, in REGISTERS for R3 <= JB_ADJ_ACC_REG, otherwise in
// a per-CTA scratch tile in global memory that never leaves L2 (each thread reads and writes only its own elements, loads
// issued ahead of the stage-3 butterfly).  Ring slots are handed back behind block barriers -- with JB_ADJ_DEFER behind the
// FIRST barrier of the next step -- by a warp that is idle in the following stage (no thread ever blocks on a "consumed"
// mbarrier); stage 2 and stage 3 synchronise per k1 group (JB_MT_GROUP).
template <class F3, int CB>
JB_GLOBAL void JB_LB((Cols3<F3, CB>::THREADS), (Cols3<F3, CB>::CPS))
k_cols3_adj_mt(const
#ifndef JB_EMU
               __grid_constant__
#endif
               JbTileSrc src, JbGeom g, cd* __restrict__ out, const cd* __restrict__ khat,
               const cd* __restrict__ tw1g, const cd* __restrict__ tw2g, cd* __restrict__ accg, int planes_se) {
    using C3 = Cols3<F3, CB>;
    constexpr int L = F3::L, R1 = F3::R1, R3 = F3::R3, M = F3::M;
    constexpr int TILE = L * CB;
    constexpr int NT = C3::THREADS;
    JB_SMEM(unsigned char, smraw);
    cd* tile0 = reinterpret_cast<cd*>(smraw);
    cd* kh0 = tile0 + 2 * TILE;
    cd* tw1 = kh0 + 2 * TILE;
    cd* tw2 = tw1 + L;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tw2 + M);     // [0..1] tile slot full, [2..3] K^ slot full
    const int tid = threadIdx.x;
    const int n = g.n, N = g.N, PH = g.PH, nn2 = n * n, T = g.T;
    const int nstrips = (PH + CB - 1) / CB;
    const long nitems = (long)planes_se * nn2 * nstrips;
    cd* acc = accg + (size_t)blockIdx.x * TILE;               // this CTA's accumulator, element k3*I3 + tid per thread
    for (int i = tid; i < L; i += NT) tw1[i] = tw1g[i];
    for (int i = tid; i < M; i += NT) tw2[i] = tw2g[i];
    struct Item { int strip, k, s, e, j, i, vlo, vhi, row0, ncols; };
    auto decode = [&](long item) -> Item {
        Item q;
        q.strip = (int)(item % nstrips);
        long r = item / nstrips;
        q.k = (int)(r % nn2);
        const int lp = (int)(r / nn2);
        q.s = lp / g.Es; q.e = g.e0 + (lp - q.s * g.Es);
        q.j = q.k % n; q.i = q.k / n;
        q.row0 = g.offs[q.j] - g.hp - g.m;                 // image row of tile row 0
        const int lo = -q.row0, hi = N - q.row0;
        q.vlo = lo > 0 ? lo : 0;
        q.vhi = hi < g.P ? hi : g.P;
        const int nc = PH - q.strip * CB;
        q.ncols = nc < CB ? nc : CB;
        return q;
    };
    auto item_of = [&](long it_) -> long {
        long v = (long)blockIdx.x + it_ * (long)gridDim.x;
        return v < nitems ? v : -1;
    };
    auto k_off = [&](const Item& q, int t) -> size_t {
        return ((((size_t)t * g.E + q.e) * nn2 + q.k) * nstrips + q.strip) * ((size_t)L * CB);
    };
    long my_items = 0;
    if ((long)blockIdx.x < nitems) my_items = (nitems - 1 - (long)blockIdx.x) / (long)gridDim.x + 1;
    const long nsteps = my_items * T;
#ifndef JB_EMU
    auto issue_tile = [&](long step) {
        const Item q = decode(item_of(step / T));
        const int t = (int)(step % T), ts = (int)(step & 1);
        cd* tile = tile0 + ts * TILE;
        const int BR = src.BR, in_rows = g.P;
        int nch = (in_rows + BR - 1) / BR, issued = 0;
        for (int c = 0; c < nch; ++c) {
            int r0 = c * BR;
            if (r0 + BR > in_rows) r0 = in_rows - BR;
            if (r0 + BR > q.vlo && r0 < q.vhi) ++issued;
        }
        jb_mbar_expect_tx(&bars[ts], (uint32_t)(issued * BR * CB * sizeof(cd)));
        const int slab = (((q.s * T + t) * g.E + q.e)) * n + q.i;
        for (int c = 0; c < nch; ++c) {
            int r0 = c * BR;
            if (r0 + BR > in_rows) r0 = in_rows - BR;
            if (r0 + BR > q.vlo && r0 < q.vhi)
                jb_tma_load_3d(tile + (size_t)r0 * CB, &src.map, 2 * q.strip * CB, q.row0 + r0, slab, &bars[ts]);
        }
    };
    auto issue_k = [&](long step) {
        const int ks = (int)(step & 1);
        jb_mbar_expect_tx(&bars[2 + ks], (uint32_t)(TILE * sizeof(cd)));
        jb_bulk_load_1d(kh0 + ks * TILE, khat + k_off(decode(item_of(step / T)), (int)(step % T)), (uint32_t)(TILE * sizeof(cd)),
                        &bars[2 + ks]);
    };
    if (tid == 0) {
        for (int b = 0; b < 4; ++b) jb_mbar_init(&bars[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (nsteps > 0) { issue_tile(0); issue_k(0); }
        if (nsteps > 1) { issue_tile(1); issue_k(1); }
    }
#endif
    JB_SYNC();
    for (long it = 0; it < my_items; ++it) {
        const Item q = decode(item_of(it));
        constexpr bool ACC_REG = R3 <= JB_ADJ_ACC_REG;
        constexpr bool GROUP = (JB_MT_GROUP & 1) && JB_ADJ_DEFER && C3::GROUPS;   // r2z: L = 672 cols_adj 9.44 -> 8.94 ms
        // deferred refills are issued by the first lane of a warp without a stage-2 item (group mode: the warps of group 0
        // behind its GI2 items; otherwise the warps behind the I2 items), thread 0 when there is none -- see k_cols3_fwd_mt
        constexpr int S2_BUSY = GROUP ? (C3::GI2 + 31) / 32 * 32 : (C3::I2 + 31) / 32 * 32;
        constexpr int ISSUER = !(JB_MT_ISSUER && JB_ADJ_DEFER) ? 0
                               : (GROUP && C3::I3 < NT) ? C3::I3               // a warp outside every group: idle in stages 2 and 3
                               : (S2_BUSY < (GROUP ? C3::GT : NT)) ? S2_BUSY : 0;
        cd a[R3];                 // spectral-domain sum over the modules: R3 values per stage-3 thread (registers if ACC_REG)
        for (int t = 0; t < T; ++t) {
            const long step = it * T + t;
            const int ts = (int)(step & 1);
            const bool lastm = t == T - 1;
            cd* tile = tile0 + ts * TILE;
            const cd* kh = kh0 + ts * TILE;
#ifndef JB_EMU
            jb_mbar_wait<false>(&bars[ts], (uint32_t)((step >> 1) & 1));
#else
            JB_SYNC();
            for (int idx = tid; idx < g.P * CB; idx += NT) {
                int r = idx / CB, b = idx - r * CB, col = q.strip * CB + b, ir = q.row0 + r;
                tile[(size_t)r * CB + b] = (col < PH && ir >= 0 && ir < src.N)
                    ? src.base[((size_t)((((q.s * T + t) * g.E + q.e)) * n + q.i) * src.N + ir) * PH + col] : cd_make(0.0, 0.0);
            }
            for (int idx = tid; idx < TILE; idx += NT) kh0[ts * TILE + idx] = khat[k_off(q, t) + idx];
            JB_SYNC();
#endif
            C3::s1_fwd(tile, tw1, tid, q.vlo, q.vhi, nullptr, 0);
            JB_SYNC();
#ifndef JB_EMU
            // every warp is past stage 1' of the previous item's last module / stage 3 of the previous step: when that step
            // was a LAST module its tile slot is free only now (its inverse stages used it).  JB_ADJ_DEFER: ALL refills of
            // the previous step's slots are issued here, so that the block barrier behind stage 3 is only needed in front of
            // the inverse stages (last module) -- 2 barriers per module instead of 3, for a third of a step less prefetch lead
            if (JB_ADJ_DEFER) {
                if (tid == ISSUER && step >= 1 && step + 1 < nsteps) { issue_k(step + 1); issue_tile(step + 1); }
            } else {
                if (tid == 0 && step >= 1 && ((step - 1) % T) == T - 1 && step + 1 < nsteps) issue_tile(step + 1);
            }
#endif
            if (GROUP) {
                const int grp = tid / C3::GT, gr = tid - grp * C3::GT;
                if (tid < C3::I3)
                    for (int idx = gr; idx < C3::GI2; idx += C3::GT) C3::s2_fwd(tile, tw2, grp * C3::GI2 + idx);
                JB_GROUP_SYNC(tid < C3::I3, C3::NG, grp, C3::GT);
            } else {
                C3::s2_fwd(tile, tw2, tid);
                JB_SYNC();
            }
#ifndef JB_EMU
            jb_mbar_wait<false>(&bars[2 + ts], (uint32_t)((step >> 1) & 1));
#endif
            if (tid < C3::I3) {
                if (!ACC_REG && t > 0) static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; a[k3] = jb_ld_cg(acc + k3 * C3::I3 + tid); });
                cd v[R3];
                C3::s3_fwd(tile, tid, v);
                static_for<0, R3>([&](auto kk) {
                    constexpr int k3 = kk;
                    cd y = cd_mulc(v[k3], kh[k3 * C3::I3 + tid]);
                    if (t > 0) y = cd_add(a[k3], y);          // modules in ascending order: fixed summation order
                    v[k3] = y;
                });
                if (!lastm) {
                    if (ACC_REG) static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; a[k3] = v[k3]; });
                    else static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; acc[k3 * C3::I3 + tid] = v[k3]; });
                } else {          // last module: the sum goes straight into the inverse transform, in this tile slot
                    dft<R3, +1>(v);
                    const int c = tid / CB, b = tid - c * CB;
                    cd* base = tile + (size_t)(R3 * c) * CB + b;
                    static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; base[n3 * CB] = v[n3]; });
                }
            }
#ifndef JB_EMU
            jb_fence_proxy_async();
            if (!JB_ADJ_DEFER || lastm) JB_SYNC();   // stage 3 done everywhere: this step's K^ slot (and, unless it feeds the inverse, its tile slot) is free
            if (!JB_ADJ_DEFER && tid == 0 && step + 2 < nsteps) {
                issue_k(step + 2);
                if (!lastm) issue_tile(step + 2);
            }
#else
            JB_SYNC();
#endif
            if (lastm) {
                C3::s2_inv(tile, tw2, tid);
                JB_SYNC();
                const bool act = tid < C3::I1;
                cd v[R1];
                if (act) C3::s1_inv_load(tile, tw1, tid, v);
#ifndef JB_EMU
                jb_fence_proxy_async();      // (the slot is refilled behind the first block barrier of the next step)
#endif
                if (act) {
                    dft<R1, +1>(v);
                    const int m_ = tid / CB, b = tid - m_ * CB;
                    if (b < q.ncols) {
                        cd* ob = out + ((((size_t)q.s * g.E + q.e) * nn2 + q.k) * (size_t)g.w) * PH + (size_t)q.strip * CB + b;
                        static_for<0, R1>([&](auto nn) {
                            constexpr int n1 = nn;
                            const int qr = M * n1 + m_ - g.m;
                            if (qr >= 0 && qr < g.w) ob[(size_t)qr * PH] = v[n1];
                        });
                    }
                }
            }
        }
    }
}

// plan time: forward 3-stage column transform of the PSF row spectra, normalised, stored as [k3][k1*R2+k2][b].
// grid (nstrips, n*n, TE), block JB_C3_THREADS, dyn smem = L*CB*16 + (L+M)*16.
template <class F3, int CB>
JB_GLOBAL void JB_LB((Cols3<F3, CB>::THREADS), 1)
k_prep_cols3(JbGeom g, const cd* __restrict__ scratch, const double* __restrict__ sums, const cd* __restrict__ tw1g,
             const cd* __restrict__ tw2g, cd* __restrict__ khat, int per_column) {
    using C3 = Cols3<F3, CB>;
    constexpr int L = F3::L, R3 = F3::R3, M = F3::M;
    constexpr int JB_C3_THREADS = C3::THREADS;
    JB_SMEM(cd, tile);
    cd* tw1 = tile + L * CB;
    cd* tw2 = tw1 + L;
    const int tid = threadIdx.x;
    const int strip = blockIdx.x, k = blockIdx.y, te = blockIdx.z;
    const int nn2 = g.n * g.n, PH = g.PH, dx = g.dx, w = g.w;
    const int nstrips = gridDim.x, col0 = strip * CB;
    const size_t pk = (size_t)te * nn2 + k;
    const cd* ib = scratch + pk * w * PH;
    cd* kb = khat + (pk * nstrips + strip) * ((size_t)L * CB);
    for (int i = tid; i < L; i += JB_C3_THREADS) tw1[i] = tw1g[i];
    for (int i = tid; i < M; i += JB_C3_THREADS) tw2[i] = tw2g[i];
    // rows of the L-periodic prepared kernel: offset d in [0,dx) -> row d ; d in [-dx,0) -> row L+d ; scratch row q = d+dx
    for (int idx = tid; idx < L * CB; idx += JB_C3_THREADS) {
        int p = idx / CB, b = idx - p * CB, col = col0 + b, q = -1;
        if (p < dx) q = p + dx;
        else if (p >= L - dx) q = p - L + dx;
        tile[idx] = (q >= 0 && col < PH) ? ib[(size_t)q * PH + col] : cd_make(0.0, 0.0);
    }
    JB_SYNC();
    C3::s1_fwd(tile, tw1, tid, 0, L, nullptr, 0);
    JB_SYNC();
    C3::s2_fwd(tile, tw2, tid);
    JB_SYNC();
    if (tid < C3::I3) {
        cd v[R3];
        C3::s3_fwd(tile, tid, v);
        const double scale = 1.0 / (sums[pk] * (double)L * (double)L);
        if (!per_column) {
            static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; kb[k3 * C3::I3 + tid] = cd_scale(v[k3], scale); });
        } else {
            // layout of the warp-private column pass (cols_wp.cuh): (TE, n*n, PH, L) with L = [k3][k1*R2 + k2]
            const int c = tid / CB, b = tid - c * CB, col = col0 + b;
            if (col < PH) {
                cd* kc = khat + (pk * PH + col) * (size_t)L;
                static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; kc[k3 * (F3::R1 * F3::R2) + c] = cd_scale(v[k3], scale); });
            }
        }
    }
}
