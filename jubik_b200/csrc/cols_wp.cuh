// cols_wp.cuh -- column pass with WARP-PRIVATE transforms: every warp owns one spectrum column of one patch.
//
// Why (r2k, tools/micro/stage_overlap.cu on B200): one FFT stage run the way cols3.cuh runs it -- every thread loads R
// values from shared memory, R-point butterfly, stores, BLOCK BARRIER -- takes 1399 (2 x 256 threads) to 1522 (1 x 512)
// cycles where its shared-memory traffic alone takes 1087 and its FP64 work alone 796; WITHOUT the barrier the same
// instruction stream takes 1100 = max(shared memory, FP64): the two pipes overlap perfectly as soon as warps are free to
// drift apart.  The barriers are needed there because a CB-column tile is transformed by the whole CTA.  Here a column is
// private to a warp: the three forward stages, the product with the resident PSF spectrum and the three inverse stages
// run on the warp's own padded tile with warp-level syncs only (the layout of rows5.cuh: odd pitches, conflict-free in
// every stage), 16 free-running warps per SM.
//   * input : the warp gathers its column (16 bytes per image row) with cp.async straight into the padded layout;
//   * K^    : stored per column ([plane][patch][column][k3][k1*R2+k2]), warmed in L2 by one bulk prefetch at the start of
//             the unit and read into registers at the top of stage 3 (no staging ring, no TMA tile);
//   * output: stored from registers row by row (16 bytes per row; the neighbouring columns' stores come from the
//             neighbouring warps and merge in L2).
// Index maps as in cols3.cuh: n = M*n1 + R3*n2 + n3, k = k1 + R1*(k2 + R2*k3), position (a, b, c) at a*BP + b*R3P + c.
#pragma once
#include "cols3.cuh"
#include "rows_fast.cuh"

#ifndef JB_EMU
JB_D void jb_cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#else
inline void jb_cp_async_wait0() {}
#endif

template <class F3>
struct ColsWp {
    static constexpr int L = F3::L, R1 = F3::R1, R2 = F3::R2, R3 = F3::R3, M = F3::M;
    static constexpr int R3P = R3 | 1;              // odd pitch of a k2 block
    static constexpr int BP = (R2 * R3P) | 1;       // odd pitch of a k1 block
    static constexpr int SEQ = R1 * BP;             // cd per column
    static constexpr int I1 = M, I2 = R1 * R3, I3 = R1 * R2;
    static constexpr int N1 = (I1 + 31) / 32, N2 = (I2 + 31) / 32, N3 = (I3 + 31) / 32;
    static JB_HD int addr_of_sample(int p) {
        int n1 = p / M, mm = p - n1 * M, n2 = mm / R3, n3 = mm - n2 * R3;
        return n1 * BP + n2 * R3P + n3;
    }
};

// dynamic shared memory: twiddles (L + M cd), window row weights (w doubles, forward), NW private columns
template <class F3>
inline size_t jb_cols_wp_smem(int w, int nw) {
    return (size_t)(F3::L + F3::M) * sizeof(cd) + (size_t)((w + 1) & ~1) * sizeof(double) + (size_t)nw * ColsWp<F3>::SEQ * sizeof(cd);
}

// grid: persistent, block = NW warps.  src: row spectra (slabs, N, PH) -- forward: rowsF, adjoint: rowsA; out: per-patch
// rows (planes, n*n, out_rows, PH); khat: per-column layout (TE, n*n, PH, L).
template <class F3, bool ADJ, int NW>
JB_GLOBAL void JB_LB(NW * 32, 2)
k_cols_wp(JbGeom g, const cd* __restrict__ src, cd* __restrict__ out, const cd* __restrict__ khat,
          const cd* __restrict__ tw1g, const cd* __restrict__ tw2g, const double* __restrict__ wgt, int planes) {
    using W = ColsWp<F3>;
    constexpr int L = W::L, R1 = W::R1, R2 = W::R2, R3 = W::R3, M = W::M, BP = W::BP, R3P = W::R3P;
    JB_SMEM(unsigned char, smraw);
    cd* tw1 = reinterpret_cast<cd*>(smraw);
    cd* tw2 = tw1 + L;
    double* rw = reinterpret_cast<double*>(tw2 + M);
    cd* bufs = reinterpret_cast<cd*>(rw + ((g.w + 1) & ~1));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < L; i += NW * 32) tw1[i] = tw1g[i];
    for (int i = tid; i < M; i += NW * 32) tw2[i] = tw2g[i];
    if (!ADJ) for (int i = tid; i < g.w; i += NW * 32) rw[i] = wgt[i];
    JB_SYNC();
    cd* buf = bufs + (size_t)wid * W::SEQ;
    const int n = g.n, N = g.N, PH = g.PH, nn2 = n * n;
    const long nunits = (long)planes * nn2 * PH;
    const long stride = (long)gridDim.x * NW;
    const long niter = (nunits + stride - 1) / stride;
    const int in_p0 = ADJ ? 0 : g.m, in_rows = ADJ ? g.P : g.w, out_p0 = ADJ ? g.m : 0, out_rows = ADJ ? g.w : g.P;
    for (long it = 0; it < niter; ++it) {
        const long unit = it * stride + (long)blockIdx.x * NW + wid;
        const bool live = unit < nunits;      // (the warp syncs below are executed by every warp: emulation needs that)
        int vlo = 0, vhi = 0;
        const cd* kb = khat;
        cd* ob = out;
        if (live) {
            const int c = (int)(unit % PH);
            const long r = unit / PH;
            const int k = (int)(r % nn2), lp = (int)(r / nn2);
            int plane, te, se;
            jb_map_ste(g, lp, plane, te, se);
            const int j = k % n, i = k / n;
            const int row0 = g.offs[j] - g.hp - (ADJ ? g.m : 0);      // image row of tile row in_p0
            const int lo = -row0, hi = N - row0;
            vlo = in_p0 + (lo > 0 ? lo : 0);
            vhi = in_p0 + (hi < in_rows ? hi : in_rows);
            const cd* sb = src + ((size_t)((ADJ ? plane : se) * n + i) * N) * PH + c;
            ob = out + (((size_t)plane * nn2 + k) * out_rows) * PH + c;
            kb = khat + (((size_t)te * nn2 + k) * PH + c) * L;
            // ---- gather the column: tile row p <- image row row0 + p - in_p0, straight into the padded layout
            for (int p = vlo + lane; p < vhi; p += 32)
                jb_cp_async16(buf + W::addr_of_sample(p), sb + (size_t)(row0 + p - in_p0) * PH);
#ifndef JB_EMU
            if (lane == 0) jb_bulk_prefetch_l2(kb, (uint32_t)(L * sizeof(cd)));
#endif
        }
        jb_cp_async_commit();
        jb_cp_async_wait0();
        JB_SYNCWARP();
        // ---- stage 1 forward: radix R1 over n1, * W_L^{m k1}; rows outside [vlo, vhi) are zero; forward: * row weight
        if (live) static_for<0, W::N1>([&](auto rr) {
            const int m_ = lane + 32 * (int)rr;
            if (m_ < W::I1) {
                const int n2 = m_ / R3, n3 = m_ - n2 * R3;
                cd* b = buf + n2 * R3P + n3;
                cd v[R1];
                static_for<0, R1>([&](auto nn) {
                    constexpr int n1 = nn;
                    const int p = M * n1 + m_;
                    cd t = cd_make(0.0, 0.0);
                    if (p >= vlo && p < vhi) {
                        t = b[n1 * BP];
                        if (!ADJ) t = cd_scale(t, rw[p - g.m]);
                    }
                    v[n1] = t;
                });
                dft<R1, -1>(v);
                static_for<0, R1>([&](auto kk) {
                    constexpr int k1 = kk;
                    cd y = v[k1];
                    if constexpr (k1 > 0) y = cd_mul(y, tw1[k1 * M + m_]);
                    b[k1 * BP] = y;
                });
            }
        });
        JB_SYNCWARP();
        // ---- stage 2 forward: radix R2 over n2, * W_M^{n3 k2}
        if (live) static_for<0, W::N2>([&](auto rr) {
            const int c2 = lane + 32 * (int)rr;
            if (c2 < W::I2) {
                const int k1 = c2 / R3, n3 = c2 - k1 * R3;
                cd* b = buf + k1 * BP + n3;
                cd v[R2];
                static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; v[n2] = b[n2 * R3P]; });
                dft<R2, -1>(v);
                static_for<0, R2>([&](auto kk) {
                    constexpr int k2 = kk;
                    cd y = v[k2];
                    if constexpr (k2 > 0) y = cd_mul(y, tw2[k2 * R3 + n3]);
                    b[k2 * R3P] = y;
                });
            }
        });
        JB_SYNCWARP();
        // ---- stage 3: forward radix R3, * K^ (conj for the adjoint), inverse radix R3 -- the spectrum stays in registers
        if (live) static_for<0, W::N3>([&](auto rr) {
            const int c3 = lane + 32 * (int)rr;
            if (c3 < W::I3) {
                const int k1 = c3 / R2, k2 = c3 - k1 * R2;
                cd* b = buf + k1 * BP + k2 * R3P;
                cd v[R3], kq[R3];
                static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; kq[k3] = jb_ld_nc(kb + k3 * W::I3 + c3); });
                static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; v[n3] = b[n3]; });
                dft<R3, -1>(v);
                static_for<0, R3>([&](auto kk) {
                    constexpr int k3 = kk;
                    v[k3] = ADJ ? cd_mulc(v[k3], kq[k3]) : cd_mul(v[k3], kq[k3]);
                });
                dft<R3, +1>(v);
                static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; b[n3] = v[n3]; });
            }
        });
        JB_SYNCWARP();
        // ---- stage 2 inverse
        if (live) static_for<0, W::N2>([&](auto rr) {
            const int c2 = lane + 32 * (int)rr;
            if (c2 < W::I2) {
                const int k1 = c2 / R3, n3 = c2 - k1 * R3;
                cd* b = buf + k1 * BP + n3;
                cd v[R2];
                static_for<0, R2>([&](auto kk) {
                    constexpr int k2 = kk;
                    cd y = b[k2 * R3P];
                    if constexpr (k2 > 0) y = cd_mulc(y, tw2[k2 * R3 + n3]);
                    v[k2] = y;
                });
                dft<R2, +1>(v);
                static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; b[n2 * R3P] = v[n2]; });
            }
        });
        JB_SYNCWARP();
        // ---- stage 1 inverse: the result rows leave from registers
        if (live) static_for<0, W::N1>([&](auto rr) {
            const int m_ = lane + 32 * (int)rr;
            if (m_ < W::I1) {
                const int n2 = m_ / R3, n3 = m_ - n2 * R3;
                const cd* b = buf + n2 * R3P + n3;
                cd v[R1];
                static_for<0, R1>([&](auto kk) {
                    constexpr int k1 = kk;
                    cd y = b[k1 * BP];
                    if constexpr (k1 > 0) y = cd_mulc(y, tw1[k1 * M + m_]);
                    v[k1] = y;
                });
                dft<R1, +1>(v);
                static_for<0, R1>([&](auto nn) {
                    constexpr int n1 = nn;
                    const int q = M * n1 + m_ - out_p0;
                    if (q >= 0 && q < out_rows) ob[(size_t)q * PH] = v[n1];
                });
            }
        });
        JB_SYNCWARP();   // the next unit's gather overwrites the tile
    }
}
