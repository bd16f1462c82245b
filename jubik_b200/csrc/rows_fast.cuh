// rows_fast.cuh -- the four row kernels for two-pass transform lengths, written for low instruction
// overhead: every per-item address is hoisted into a few base pointers, every per-element access is
// `pointer + compile-time offset`, the direct/mirrored half-spectrum choice is resolved at compile time
// (one run-time case per item), inter-pass twiddles come from two per-item seeds by recurrence, and the
// row-patch lists are unrolled to JB_MAXQ entries.  Same maths and data layout as kernels.cuh (which
// keeps the generic single-pass versions used by the tiny test lengths).
#pragma once
#include "kernels.cuh"

#define JB_MAXQ 4  // row-patches j that can cover one image row: ceil(P/dx) (+1 with clamped offsets)

// Hermitian half-spectrum element k = k1 + R1*K2 of a length-L real transform lives at index k when
// 2k <= L, else at L-k conjugated.  For K2 fixed at compile time only 2*K2 == R2 depends on k1.
template <class FP, int K2>
JB_D bool jb_spec_direct(int k1) {
    if constexpr (2 * K2 < FP::R2) return true;
    else if constexpr (2 * K2 == FP::R2) return k1 == 0;
    else return false;
}

// ---- shared pass bodies -------------------------------------------------------------------
// r2c pass 1 tail: real radix-R1 butterfly, twiddle W_L^{n2 k1} (seeds W^{n2}, W^{2 n2}), store to the exchange buffer
template <class FP>
JB_D void jb_r2c_pass1(const double (&x)[FP::R1], cd* dst /* sm + b*LS + n2 */, cd w1, cd w2) {
    constexpr int H1 = FP::H1, R2P = FP::R2P;
    cd Y[H1];
    rdft_fwd<FP::R1>(x, Y);
    cd po = w1, pe = w2;
    static_for<0, H1>([&](auto kk) {
        constexpr int k1 = kk;
        cd y = Y[k1];
        if constexpr (k1 > 0) {
            if constexpr (k1 & 1) { y = cd_mul(y, po); if constexpr (k1 + 2 < H1) po = cd_mul(po, w2); }
            else { y = cd_mul(y, pe); if constexpr (k1 + 2 < H1) pe = cd_mul(pe, w2); }
        }
        dst[k1 * R2P] = y;
    });
}
// r2c pass 2 head: load the k1-th row of the exchange buffer and transform over n2
template <class FP>
JB_D void jb_r2c_pass2(const cd* src /* sm + b*LS + k1*R2P */, cd (&u)[FP::R2]) {
    static_for<0, FP::R2>([&](auto nn) { constexpr int n2 = nn; u[n2] = src[n2]; });
    dft<FP::R2, -1>(u);
}
// c2r pass A tail: inverse radix-R2 over k2, twiddle W_L^{-n2 k1} (seeds W^{-k1}, W^{-2 k1}), store
template <class FP>
JB_D void jb_c2r_passA(cd (&u)[FP::R2], cd* dst /* sm + b*LS + k1*R2P */, cd w1, cd w2) {
    dft<FP::R2, +1>(u);
    cd po = w1, pe = w2;
    static_for<0, FP::R2>([&](auto nn) {
        constexpr int n2 = nn;
        cd y = u[n2];
        if constexpr (n2 > 0) {
            if constexpr (n2 & 1) { y = cd_mul(y, po); if constexpr (n2 + 2 < FP::R2) po = cd_mul(po, w2); }
            else { y = cd_mul(y, pe); if constexpr (n2 + 2 < FP::R2) pe = cd_mul(pe, w2); }
        }
        dst[n2] = y;
    });
}
// c2r pass B head: gather the Hermitian k1 column and run the real inverse butterfly
template <class FP>
JB_D void jb_c2r_passB(const cd* src /* sm + b*LS + n2 */, double (&x)[FP::R1]) {
    cd U[FP::H1];
    static_for<0, FP::H1>([&](auto kk) { constexpr int k1 = kk; U[k1] = src[k1 * FP::R2P]; });
    rdft_inv<FP::R1>(U, x);
}

// Row-patch table of the CTA's image rows in shared memory: for row rr, entry q: element offset of
// (patch row-block j, row u) inside one plane of the spectrum buffer *excluding* the i*n*rows*PH term,
// plus the window row weight.  cnt is clamped to JB_MAXQ (plan creation checks the geometry).
struct JbRowTab {
    int* cnt;        // [rpc]
    long long* off;  // [rpc*JB_MAXQ]
    double* wr;      // [rpc*JB_MAXQ]
};
JB_D size_t jb_rowtab_bytes(int rpc) {
    return jb_align16(sizeof(int) * rpc) + sizeof(long long) * rpc * JB_MAXQ + sizeof(double) * rpc * JB_MAXQ;
}
JB_D JbRowTab jb_rowtab_at(unsigned char* p, int rpc) {
    JbRowTab t;
    t.cnt = reinterpret_cast<int*>(p);
    t.off = reinterpret_cast<long long*>(p + jb_align16(sizeof(int) * rpc));
    t.wr = reinterpret_cast<double*>(t.off + rpc * JB_MAXQ);
    return t;
}
JB_D void jb_rowtab_build(const JbGeom& g, const JbRowTab& t, int r0, int rpc, int shift, int rows, const double* wgt) {
    for (int rr = jb_tid(); rr < rpc; rr += jb_nthr()) {
        int r = r0 + rr, c = 0;
        if (r < g.N) {
            for (int j = 0; j < g.n && c < JB_MAXQ; ++j) {
                int u = r + shift - g.offs[j];
                if (u >= 0 && u < rows) {
                    t.off[rr * JB_MAXQ + c] = ((long long)j * rows + u) * g.PH;
                    t.wr[rr * JB_MAXQ + c] = wgt ? wgt[u] : 1.0;
                    ++c;
                }
            }
        }
        t.cnt[rr] = c;
        for (int q = c; q < JB_MAXQ; ++q) { t.off[rr * JB_MAXQ + q] = 0; t.wr[rr * JB_MAXQ + q] = 0.0; }
    }
}

// store the Hermitian half spectrum held as u[k2] = X[k1 + R1*k2] to one contiguous row (k = 0..L/2)
template <class FP>
JB_D void jb_store_half_row(cd* row, const cd (&u)[FP::R2], int k1) {
    constexpr int L = FP::L, R1 = FP::R1, R2 = FP::R2;
    const bool mirror_ok = (k1 != 0) && (2 * k1 != R1);
    cd* pd = row + k1;        // direct:  + R1*k2
    cd* pm = row + (L - k1);  // mirror:  - R1*k2
    static_for<0, R2>([&](auto kk) {
        constexpr int k2 = kk;
        if (jb_spec_direct<FP, k2>(k1)) pd[R1 * k2] = u[k2];
        else if (mirror_ok) pm[-R1 * k2] = cd_conj(u[k2]);
    });
}

// ------------------------------------------------------------------------------------------------------------
// Persistent r2c row kernel with the NEXT image row prefetched into shared memory (cp.async, no registers held).
// ncu of the one-row-per-CTA predecessors (r1o): 58 % of the stall samples sit on the first use of the 24 strided global
// loads of pass 1 and 10 % on EXIT (stores draining) -- four one-row CTAs per SM cannot hide a DRAM round trip.  Here a
// CTA walks the (plane, image row) list with a stride of gridDim.x; while it transforms row k the 8 KB of row k+1 land
// in the other half of a two-slot ring, and pass 1 reads its window elements from shared memory.
//   ADJ = false: sky rows, window weights, w-wide windows at offs[i]-hp            -> rowsF (S*E, n, N, PH)
//   ADJ = true : cotangent rows (exposure already folded in), P-wide windows       -> rowsA (STE, n, N, PH)
// dyn smem = n*LS*16 (exchange) + 2*N*8 (row ring) + w*8 (forward: window weights) or 2*N*8 (adjoint with exposure).
#ifndef JB_EMU
JB_D void jb_cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
JB_D void jb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
JB_D void jb_cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
#else
inline void jb_cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
inline void jb_cp_async_commit() {}
inline void jb_cp_async_wait1() {}
#endif

template <class FP, bool ADJ>
JB_GLOBAL void JB_LB(256, 2) k_rows_in_pf(JbGeom g, const double* __restrict__ in, const double* __restrict__ aux,
                                          const cd* __restrict__ tw_nk, cd* __restrict__ rows, int planes, int exp_in) {
    // aux: forward = window weights (w); adjoint = effective exposure (T,E,N,N) to multiply the cotangent with, or null
    // exp_in (forward): the input rows are log_sky; x = exp(log_sky) is formed in the shared-memory ring (sky_models.py:340-341)
    constexpr int R1 = FP::R1, R2 = FP::R2, H1 = FP::H1, LS = FP::LS, R2P = FP::R2P;
    JB_SMEM(unsigned char, smraw);
    const int n = g.n, N = g.N, m = g.m, w = g.w, hp = g.hp, PH = g.PH, P = g.P;
    cd* sm = reinterpret_cast<cd*>(smraw);
    double* ring = reinterpret_cast<double*>(smraw + jb_align16((size_t)n * LS * sizeof(cd)));
    double* ext = ring + 2 * (size_t)N;      // forward: window weights (w); adjoint with exposure: its two-slot row ring
    const bool with_expo = ADJ && aux != nullptr;
    if (!ADJ) for (int q = jb_tid(); q < w; q += jb_nthr()) ext[q] = aux[q];   // ordered by the first barrier
    const int total = planes * N;
    // launched plane -> plane index of the input / output buffers (and of the exposure)
    auto plane_of = [&](int lp, int& te) -> int {
        int a, b2, c2;
        if (ADJ) { jb_map_ste(g, lp, a, b2, c2); te = b2; return a; }
        te = 0;
        return jb_map_se(g, lp, a, b2);
    };
    auto prefetch = [&](int idx, int slot) {
        const int lp = idx / N, r = idx - lp * N;
        int te;
        const int pl = plane_of(lp, te);
        const double* src = in + ((size_t)pl * N + r) * N;
        if (!ADJ) { int s_, e_; jb_map_se(g, lp, s_, e_); src = in + jb_sky_row(g, s_, e_, r); }   // (padded log-sky: strided)
        double* dst = ring + (size_t)slot * N;
        for (int c = 2 * jb_tid(); c < N; c += 2 * jb_nthr()) jb_cp_async16(dst + c, src + c);
        if (with_expo) {
            const double* src2 = aux + ((size_t)te * N + r) * N;
            double* dst2 = ext + (size_t)slot * N;
            for (int c = 2 * jb_tid(); c < N; c += 2 * jb_nthr()) jb_cp_async16(dst2 + c, src2 + c);
        }
    };
    int idx = blockIdx.x;
    if (idx < total) prefetch(idx, 0);
    jb_cp_async_commit();
    for (int k = 0; idx < total; idx += gridDim.x, ++k) {
        const int nxt = idx + gridDim.x;
        if (nxt < total) prefetch(nxt, (k + 1) & 1);
        jb_cp_async_commit();
        jb_cp_async_wait1();          // this thread's copies of the current row have landed ...
        JB_SYNC();                    // ... and so have everybody else's
        double* row = ring + (size_t)(k & 1) * N;
        const double* erow = ext + (size_t)(k & 1) * N;
        if (!ADJ && exp_in) {
            for (int c = jb_tid(); c < N; c += jb_nthr()) row[c] = exp(row[c]);
            JB_SYNC();
        }
        const int lp = idx / N, r = idx - lp * N;
        int te_;
        const int pl = plane_of(lp, te_);
        for (int it = jb_tid(); it < n * R2; it += jb_nthr()) {
            const int i = it / R2, n2 = it - i * R2;
            double xv[R1];
            if (!ADJ) {
                // element p = R2*n1 + n2 of the zero-padded window row: q = p - m in [0,w), column c = offs[i]-hp+q
                const int q0 = n2 - m, c0 = g.offs[i] - hp + q0;
                static_for<0, R1>([&](auto nn) {
                    constexpr int n1 = nn;
                    const int q = q0 + R2 * n1, c = c0 + R2 * n1;
                    const bool ok = q >= 0 && q < w && c >= 0 && c < N;
                    xv[n1] = ok ? row[c] * ext[q] : 0.0;
                });
            } else {
                const int c0 = g.offs[i] - hp - m + n2;
                static_for<0, R1>([&](auto nn) {
                    constexpr int n1 = nn;
                    const int p = n2 + R2 * n1, c = c0 + R2 * n1;
                    const bool ok = p < P && c >= 0 && c < N;
                    double v = ok ? row[c] : 0.0;
                    if (with_expo && ok) v *= erow[c];
                    xv[n1] = v;
                });
            }
            jb_r2c_pass1<FP>(xv, sm + (size_t)i * LS + n2, jb_ldg(tw_nk + R2 + n2), jb_ldg(tw_nk + 2 * R2 + n2));
        }
        JB_SYNC();
        for (int it = jb_tid(); it < n * H1; it += jb_nthr()) {
            const int i = it / H1, k1 = it - i * H1;
            cd u[R2];
            jb_r2c_pass2<FP>(sm + (size_t)i * LS + k1 * R2P, u);
            jb_store_half_row<FP>(rows + (((size_t)pl * n + i) * N + r) * PH, u, k1);
        }
        JB_SYNC();   // exchange buffer and ring slot k&1 are free for the next round
    }
}

// load the half-spectrum element (k1, K2) as the weighted sum over the row-patch table entries
template <class FP, int K2, bool WEIGHTED>
JB_D cd jb_sum_spec(const cd* const (&pd)[JB_MAXQ], const cd* const (&pm)[JB_MAXQ], const double (&wr)[JB_MAXQ],
                    int cnt, int k1) {
    cd acc = cd_make(0.0, 0.0);
    const bool direct = jb_spec_direct<FP, K2>(k1);
    static_for<0, JB_MAXQ>([&](auto qq) {
        constexpr int q = qq;
        if (q < cnt) {
            cd v = direct ? pd[q][FP::R1 * K2] : pm[q][-FP::R1 * K2];
            if (WEIGHTED) { acc.x += wr[q] * v.x; acc.y += wr[q] * v.y; }
            else { acc.x += v.x; acc.y += v.y; }
        }
    });
    if (!direct) acc.y = -acc.y;
    return acc;
}

// --------------------------------------------- forward: spectra -> image rows (+ epilogue) (c2r)
// grid (ceil(N/rpc), S*T*E)
template <class FP>
JB_GLOBAL void JB_LB(256, 2) k_rows_out2(JbGeom g, const cd* __restrict__ bufA, const cd* __restrict__ tw_kn, JbEpi ep, int rpc) {
    constexpr int L = FP::L, R1 = FP::R1, R2 = FP::R2, H1 = FP::H1, LS = FP::LS, R2P = FP::R2P;
    JB_SMEM(unsigned char, smraw);
    const int n = g.n, N = g.N, m = g.m, hp = g.hp, PH = g.PH, P = g.P;
    const int B = rpc * n;
    size_t off = 0;
    cd* sm = reinterpret_cast<cd*>(smraw);
    off += jb_align16((size_t)B * LS * sizeof(cd));
    double* tile = reinterpret_cast<double*>(smraw + off);
    off += jb_align16((size_t)B * P * sizeof(double));
    double* red = reinterpret_cast<double*>(smraw + off);
    off += jb_align16((size_t)jb_nthr() * sizeof(double));
    JbRowTab tab = jb_rowtab_at(smraw + off, rpc);
    int ste, te, se_;
    jb_map_ste(g, blockIdx.y, ste, te, se_);
    const int r0 = blockIdx.x * rpc;
    jb_rowtab_build(g, tab, r0, rpc, hp + m, P, nullptr);
    JB_SYNC();
    const cd* ab = bufA + (size_t)ste * n * n * P * PH;
    for (int it = jb_tid(); it < B * H1; it += jb_nthr()) {
        int b = it / H1, k1 = it - b * H1;
        int rr = b / n, i = b - rr * n;
        const int cnt = tab.cnt[rr];
        const cd* base = ab + (size_t)i * n * P * PH;
        const cd* pd[JB_MAXQ];
        const cd* pm[JB_MAXQ];
        double wr[JB_MAXQ];
        static_for<0, JB_MAXQ>([&](auto qq) {
            constexpr int q = qq;
            pd[q] = base + tab.off[rr * JB_MAXQ + q] + k1;
            pm[q] = base + tab.off[rr * JB_MAXQ + q] + (L - k1);
            wr[q] = 1.0;
        });
        cd w1 = cd_conj(jb_ldg(tw_kn + R1 + k1)), w2 = cd_conj(jb_ldg(tw_kn + 2 * R1 + k1));
        cd u[R2];
        static_for<0, R2>([&](auto kk) { constexpr int k2 = kk; u[k2] = jb_sum_spec<FP, k2, false>(pd, pm, wr, cnt, k1); });
        jb_c2r_passA<FP>(u, sm + (size_t)b * LS + k1 * R2P, w1, w2);
    }
    JB_SYNC();
    for (int it = jb_tid(); it < B * R2; it += jb_nthr()) {
        int b = it / R2, n2 = it - b * R2;
        double xv[R1];
        jb_c2r_passB<FP>(sm + (size_t)b * LS + n2, xv);
        double* tp = tile + (size_t)b * P + n2;
        static_for<0, R1>([&](auto nn) {
            constexpr int n1 = nn;
            if (R2 * n1 + n2 < P) tp[R2 * n1] = xv[n1];
        });
    }
    JB_SYNC();
    // overlap-add over column blocks i as a gather, then the epilogue
    const int shift = hp + m;
    double esum = 0.0;
    for (int idx = jb_tid(); idx < rpc * N; idx += jb_nthr()) {
        int rr = idx / N, c = idx - rr * N, r = r0 + rr;
        if (r >= N) continue;
        double conv = 0.0;
        const double* trow = tile + (size_t)rr * n * P + c + shift;
        for (int i = 0; i < n; ++i) {
            int p = c + shift - g.offs[i];
            if (p >= 0 && p < P) conv += trow[(size_t)i * P - g.offs[i]];
        }
        size_t o = ((size_t)ste * N + r) * N + c;
        size_t oe = ((size_t)te * N + r) * N + c;
        if (ep.mode == JB_EPI_POISSON) {
            const double ex = jb_ldg(ep.expo + oe);
            double gout = 0.0;
            if (ex != 0.0) {
                double lam = ex * conv, d = jb_ldg(ep.data + oe);
                esum += lam - d * log(lam);
                gout = ex * (1.0 - d / lam);
            }
            ep.out[o] = gout;
        } else {
            ep.out[o] = jb_epi_pixel(ep, conv, o, oe);
        }
    }
    if (ep.mode == JB_EPI_POISSON) {
        double tot = jb_block_sum(esum, red);
        if (jb_tid() == 0) ep.partial[(size_t)ste * gridDim.x + blockIdx.x] = tot;
    }
}

// ------------------------------------ adjoint: spectra -> sky-gradient rows, summed over modules (c2r)
// grid (ceil(N/rpc), S*E)
template <class FP>
JB_GLOBAL void JB_LB(256, 2) k_rows_adj_out2(JbGeom g, const cd* __restrict__ bufB, const double* __restrict__ wgt,
                               const cd* __restrict__ tw_kn, double* __restrict__ grad, const double* __restrict__ logsky,
                               int rpc) {
    constexpr int L = FP::L, R1 = FP::R1, R2 = FP::R2, H1 = FP::H1, LS = FP::LS, R2P = FP::R2P;
    JB_SMEM(unsigned char, smraw);
    const int n = g.n, N = g.N, m = g.m, hp = g.hp, PH = g.PH, w = g.w;
    const int B = rpc * n;
    size_t off = 0;
    cd* sm = reinterpret_cast<cd*>(smraw);
    off += jb_align16((size_t)B * LS * sizeof(cd));
    double* tile = reinterpret_cast<double*>(smraw + off);
    off += jb_align16((size_t)B * w * sizeof(double));
    JbRowTab tab = jb_rowtab_at(smraw + off, rpc);
    int s, e;
    const int r0 = blockIdx.x * rpc, se = jb_map_se(g, blockIdx.y, s, e);
    jb_rowtab_build(g, tab, r0, rpc, hp, w, wgt);
    JB_SYNC();
    const size_t plane_sz = (size_t)n * n * w * PH;
    for (int it = jb_tid(); it < B * H1; it += jb_nthr()) {
        int b = it / H1, k1 = it - b * H1;
        int rr = b / n, i = b - rr * n;
        const int cnt = tab.cnt[rr];
        double wr[JB_MAXQ];
        long long o[JB_MAXQ];
        static_for<0, JB_MAXQ>([&](auto qq) {
            constexpr int q = qq;
            wr[q] = tab.wr[rr * JB_MAXQ + q];
            o[q] = tab.off[rr * JB_MAXQ + q];
        });
        cd u[R2];
        static_for<0, R2>([&](auto kk) { constexpr int k2 = kk; u[k2] = cd_make(0.0, 0.0); });
        for (int t = 0; t < g.T; ++t) {
            const cd* base = bufB + ((size_t)(s * g.T + t) * g.E + e) * plane_sz + (size_t)i * n * w * PH;
            const cd* pd[JB_MAXQ];
            const cd* pm[JB_MAXQ];
            static_for<0, JB_MAXQ>([&](auto qq) {
                constexpr int q = qq;
                pd[q] = base + o[q] + k1;
                pm[q] = base + o[q] + (L - k1);
            });
            static_for<0, R2>([&](auto kk) {
                constexpr int k2 = kk;
                cd v = jb_sum_spec<FP, k2, true>(pd, pm, wr, cnt, k1);
                u[k2].x += v.x;
                u[k2].y += v.y;
            });
        }
        cd w1 = cd_conj(jb_ldg(tw_kn + R1 + k1)), w2 = cd_conj(jb_ldg(tw_kn + 2 * R1 + k1));
        jb_c2r_passA<FP>(u, sm + (size_t)b * LS + k1 * R2P, w1, w2);
    }
    JB_SYNC();
    for (int it = jb_tid(); it < B * R2; it += jb_nthr()) {
        int b = it / R2, n2 = it - b * R2;
        double xv[R1];
        jb_c2r_passB<FP>(sm + (size_t)b * LS + n2, xv);
        double* tp = tile + (size_t)b * w + (n2 - m);
        const double* pw = wgt + (n2 - m);
        static_for<0, R1>([&](auto nn) {
            constexpr int n1 = nn;
            int q = R2 * n1 + n2 - m;
            if (q >= 0 && q < w) tp[R2 * n1] = xv[n1] * jb_ldg(pw + R2 * n1);
        });
    }
    JB_SYNC();
    for (int idx = jb_tid(); idx < rpc * N; idx += jb_nthr()) {
        int rr = idx / N, c = idx - rr * N, r = r0 + rr;
        if (r >= N) continue;
        double acc = 0.0;
        const double* trow = tile + (size_t)rr * n * w + c + hp;
        for (int i = 0; i < n; ++i) {
            int v = c + hp - g.offs[i];
            if (v >= 0 && v < w) acc += trow[(size_t)i * w - g.offs[i]];
        }
        const size_t o = jb_sky_row(g, s, e, r) + c;
        if (logsky) acc *= exp(jb_ldg(logsky + o));
        grad[o] = acc;
    }
}
