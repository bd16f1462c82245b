// rows_pipe.cuh -- the two c2r row kernels as persistent, bulk-copy-fed pipelines (sm_100a).
//
// rows_out2 / rows_adj_out2 spend most of their time with every warp parked on the long scoreboard: each CTA
// first waits for ~75 KB of spectrum rows, then computes, and only 3-4 CTAs fit per SM to overlap the two.
// Here one CTA per SM loops over image rows; the spectrum rows an image row needs (n column blocks x up to Q
// row-patches, each one contiguous PH*16 B) are fetched two rows ahead with 1-D cp.async.bulk copies into a
// 2-stage shared-memory ring (mbarrier complete_tx), and the stage is handed back right after the first pass
// has summed it into registers.  Same maths as rows_fast.cuh.
#pragma once
#include "rows_fast.cuh"
#include "cols_pipe.cuh"

#define JB_RP_THREADS 256

struct JbRowInfo {  // per ring stage: which spectrum rows were staged
    int cnt;
    int pad;
    double wr[JB_MAXQ];
};

JB_D size_t jb_al128(size_t x) { return (x + 127) & ~(size_t)127; }

// column blocks i whose window of `len` samples starting at offs[i]-shift covers image column c form a
// contiguous range [lo, hi] (offs is non-decreasing); the same for every image row => tabulated once per CTA
JB_D void jb_col_ranges(const JbGeom& g, int shift, int len, unsigned char* lo, unsigned char* hi) {
    for (int c = threadIdx.x; c < g.N; c += JB_RP_THREADS) {
        int l = g.n, h = -1;
        for (int i = 0; i < g.n; ++i) {
            int p = c + shift - g.offs[i];
            if (p >= 0 && p < len) { if (i < l) l = i; h = i; }
        }
        lo[c] = (unsigned char)l;
        hi[c] = (unsigned char)(h + 1);   // exclusive
    }
}

// ------------------------------------------------------------------------------------------------------
// forward: spectra -> image rows (+ epilogue).  grid = #SMs (persistent), block = JB_RP_THREADS.
// item = plane*N + r.  bufA:(STE, n*n, P, PH).  Q = max row-patches per image row (<= JB_MAXQ).
template <class FP>
JB_GLOBAL void JB_LB(JB_RP_THREADS, 1)
k_rows_out3(JbGeom g, const cd* __restrict__ bufA, const cd* __restrict__ tw_kn, JbEpi ep, int planes, int Q) {
    constexpr int L = FP::L, R1 = FP::R1, R2 = FP::R2, H1 = FP::H1, LS = FP::LS, R2P = FP::R2P;
    JB_SMEM(unsigned char, smraw);
    const int n = g.n, N = g.N, m = g.m, hp = g.hp, PH = g.PH, P = g.P;
    const int tid = threadIdx.x;
    const size_t stage_cd = (size_t)n * Q * PH;
    size_t off = 0;
    cd* spec0 = reinterpret_cast<cd*>(smraw);
    off += jb_al128(2 * stage_cd * sizeof(cd));
    cd* sm = reinterpret_cast<cd*>(smraw + off);
    off += jb_al128((size_t)n * LS * sizeof(cd));
    double* tile = reinterpret_cast<double*>(smraw + off);
    off += jb_al128((size_t)n * P * sizeof(double));
    double* red = reinterpret_cast<double*>(smraw + off);
    off += jb_al128(JB_RP_THREADS * sizeof(double));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + off);  // full[2], done[2]
    off += 64;
    JbRowInfo* info = reinterpret_cast<JbRowInfo*>(smraw + off);
    off += 2 * sizeof(JbRowInfo);
    unsigned char* clo = smraw + off;
    unsigned char* chi = clo + N;
    const long nitems = (long)planes * N;
    const size_t nn2 = (size_t)n * n;
    const int shift = hp + m;
    jb_col_ranges(g, shift, P, clo, chi);

    // stage the spectrum rows of `item` into ring slot s_ (called by all lanes of warp 0)
    auto stage_rows = [&](long item_, int s_) {
        const int lane = tid & 31;
        const int lp = (int)(item_ / N), r = (int)(item_ - (long)lp * N);
        int ste, te_, se_;
        jb_map_ste(g, lp, ste, te_, se_);
        int jl[JB_MAXQ], ul[JB_MAXQ], cnt = 0;
        for (int j = 0; j < n && cnt < JB_MAXQ; ++j) {
            int u = r + shift - g.offs[j];
            if (u >= 0 && u < P) { jl[cnt] = j; ul[cnt] = u; ++cnt; }
        }
        cd* dst = spec0 + (size_t)s_ * stage_cd;
        const cd* ab = bufA + (size_t)ste * nn2 * P * PH;
#ifndef JB_EMU
        if (lane == 0) {
            info[s_].cnt = cnt;
            jb_mbar_expect_tx(&bars[s_], (uint32_t)((size_t)n * cnt * PH * sizeof(cd)));
        }
        __syncwarp();
        for (int c = lane; c < n * cnt; c += 32) {
            int i = c / cnt, q = c - i * cnt;
            jb_bulk_load_1d(dst + ((size_t)i * Q + q) * PH, ab + (((size_t)i * n + jl[q]) * P + ul[q]) * PH,
                            (uint32_t)(PH * sizeof(cd)), &bars[s_]);
        }
#else
        if (lane == 0) info[s_].cnt = cnt;
        for (int c = lane; c < n * cnt; c += 32) {
            int i = c / cnt, q = c - i * cnt;
            const cd* srow = ab + (((size_t)i * n + jl[q]) * P + ul[q]) * PH;
            cd* drow = dst + ((size_t)i * Q + q) * PH;
            for (int k = 0; k < PH; ++k) drow[k] = srow[k];
        }
#endif
    };

#ifndef JB_EMU
    if (tid == 0) {
        jb_mbar_init(&bars[0], 1);
        jb_mbar_init(&bars[1], 1);
        jb_mbar_init(&bars[2], JB_RP_THREADS / 32);
        jb_mbar_init(&bars[3], JB_RP_THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) {
        if ((long)blockIdx.x < nitems) stage_rows(blockIdx.x, 0);
        if ((long)blockIdx.x + gridDim.x < nitems) stage_rows((long)blockIdx.x + gridDim.x, 1);
    }
#endif
    double esum = 0.0;
    long item = blockIdx.x;
    for (unsigned it = 0; item < nitems; ++it, item += gridDim.x) {
        const int s = it & 1;
        const int lp = (int)(item / N), r = (int)(item - (long)lp * N);
        int ste, te, se_;
        jb_map_ste(g, lp, ste, te, se_);
        const cd* spec = spec0 + (size_t)s * stage_cd;
#ifndef JB_EMU
        jb_mbar_wait<false>(&bars[s], (it >> 1) & 1);
#else
        JB_SYNC();
        if (tid < 32) stage_rows(item, s);
        JB_SYNC();
#endif
        const int cnt = info[s].cnt;
        // pass A: sum the staged row-patch spectra, inverse radix-R2, twiddle, exchange
        for (int w = tid; w < n * H1; w += JB_RP_THREADS) {
            int i = w / H1, k1 = w - i * H1;
            const cd* pd[JB_MAXQ];
            const cd* pm[JB_MAXQ];
            double wr[JB_MAXQ];
            static_for<0, JB_MAXQ>([&](auto qq) {
                constexpr int q = qq;
                const cd* row = spec + ((size_t)i * Q + (q < Q ? q : 0)) * PH;
                pd[q] = row + k1;
                pm[q] = row + (L - k1);
                wr[q] = 1.0;
            });
            cd w1 = cd_conj(jb_ldg(tw_kn + R1 + k1)), w2 = cd_conj(jb_ldg(tw_kn + 2 * R1 + k1));
            cd u[R2];
            static_for<0, R2>([&](auto kk) { constexpr int k2 = kk; u[k2] = jb_sum_spec<FP, k2, false>(pd, pm, wr, cnt, k1); });
            jb_c2r_passA<FP>(u, sm + (size_t)i * LS + k1 * R2P, w1, w2);
        }
#ifndef JB_EMU
        // the ring slot is free as soon as every warp has summed it into registers: refill it two items ahead
        jb_fence_proxy_async();
        __syncwarp();
        if ((tid & 31) == 0) jb_mbar_arrive(&bars[2 + s]);
        if (tid < 32) {
            long nxt = item + 2 * (long)gridDim.x;
            if (nxt < nitems) {
                jb_mbar_wait<false>(&bars[2 + s], (it >> 1) & 1);
                stage_rows(nxt, s);
            }
        }
#endif
        JB_SYNC();
        for (int w = tid; w < n * R2; w += JB_RP_THREADS) {
            int i = w / R2, n2 = w - i * R2;
            double xv[R1];
            jb_c2r_passB<FP>(sm + (size_t)i * LS + n2, xv);
            double* tp = tile + (size_t)i * P + n2;
            static_for<0, R1>([&](auto nn) {
                constexpr int n1 = nn;
                if (R2 * n1 + n2 < P) tp[R2 * n1] = xv[n1];
            });
        }
        JB_SYNC();
        const size_t orow = ((size_t)ste * N + r) * N, erow = ((size_t)te * N + r) * N;
        for (int c = tid; c < N; c += JB_RP_THREADS) {
            double conv = 0.0;
            const double* trow = tile + c + shift;
            const int i1 = chi[c];
            for (int i = clo[c]; i < i1; ++i) conv += trow[i * P - g.offs[i]];
            if (ep.mode == JB_EPI_PSF) {
                ep.out[orow + c] = conv;
            } else {
                double ex = jb_ldg(ep.expo + erow + c);
                if (ep.mode == JB_EPI_RESP) {
                    ep.out[orow + c] = (ex != 0.0) ? ex * conv : 0.0;
                } else if (ep.mode == JB_EPI_POISSON) {
                    double gout = 0.0;
                    if (ex != 0.0) {
                        double lam = ex * conv, d = jb_ldg(ep.data + erow + c);
                        esum += lam - d * log(lam);
                        gout = ex * (1.0 - d / lam);
                    }
                    ep.out[orow + c] = gout;
                } else {
                    double gout = 0.0;
                    if (ex != 0.0) gout = ex * (ex * conv) / ep.lam[orow + c];
                    ep.out[orow + c] = gout;
                }
            }
        }
    }
    if (ep.mode == JB_EPI_POISSON) {
        // one partial per CTA and plane would need per-plane bookkeeping; with persistent CTAs a CTA's rows span
        // planes, so this kernel is only used with the split epilogue (PSF / RESP / METRIC modes) -- see capi.cu
        (void)esum; (void)red;
    }
}

// ------------------------------------------------------------------------------------------------------
// adjoint: spectra -> sky-gradient rows, summed over modules.  item = se*N + r; pipeline step = (item, t).
// bufB:(STE, n*n, w, PH).  grid = #SMs, block = JB_RP_THREADS.
template <class FP>
JB_GLOBAL void JB_LB(JB_RP_THREADS, 1)
k_rows_adj_out3(JbGeom g, const cd* __restrict__ bufB, const double* __restrict__ wgt, const cd* __restrict__ tw_kn,
                double* __restrict__ grad, int planes_se, int Q) {
    constexpr int L = FP::L, R1 = FP::R1, R2 = FP::R2, H1 = FP::H1, LS = FP::LS, R2P = FP::R2P;
    JB_SMEM(unsigned char, smraw);
    const int n = g.n, N = g.N, m = g.m, hp = g.hp, PH = g.PH, w = g.w, T = g.T;
    const int tid = threadIdx.x;
    const size_t stage_cd = (size_t)n * Q * PH;
    size_t off = 0;
    cd* spec0 = reinterpret_cast<cd*>(smraw);
    off += jb_al128(2 * stage_cd * sizeof(cd));
    cd* sm = reinterpret_cast<cd*>(smraw + off);
    off += jb_al128((size_t)n * LS * sizeof(cd));
    double* tile = reinterpret_cast<double*>(smraw + off);
    off += jb_al128((size_t)n * w * sizeof(double));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + off);
    off += 64;
    JbRowInfo* info = reinterpret_cast<JbRowInfo*>(smraw + off);
    off += 2 * sizeof(JbRowInfo);
    unsigned char* clo = smraw + off;
    unsigned char* chi = clo + N;
    jb_col_ranges(g, hp, w, clo, chi);
    const long nitems = (long)planes_se * N;
    const long nsteps_total = nitems * T;  // global step id = item*T + t ; this CTA owns items blockIdx.x + k*gridDim.x
    (void)nsteps_total;
    const size_t nn2 = (size_t)n * n;
    const size_t plane_sz = nn2 * w * PH;

    // local step index ls = it*T + t  <->  (item = blockIdx.x + it*gridDim.x, t)
    auto stage_rows = [&](long ls_, int s_) {
        const int lane = tid & 31;
        const long it_ = ls_ / T;
        const int t = (int)(ls_ - it_ * T);
        const long item_ = blockIdx.x + it_ * (long)gridDim.x;
        const int lp = (int)(item_ / N), r = (int)(item_ - (long)lp * N);
        int s_idx, e;
        jb_map_se(g, lp, s_idx, e);
        int jl[JB_MAXQ], ul[JB_MAXQ], cnt = 0;
        for (int j = 0; j < n && cnt < JB_MAXQ; ++j) {
            int u = r + hp - g.offs[j];
            if (u >= 0 && u < w) { jl[cnt] = j; ul[cnt] = u; ++cnt; }
        }
        cd* dst = spec0 + (size_t)s_ * stage_cd;
        const cd* bb = bufB + ((size_t)(s_idx * T + t) * g.E + e) * plane_sz;
        if (lane == 0) {
            info[s_].cnt = cnt;
            for (int q = 0; q < JB_MAXQ; ++q) info[s_].wr[q] = q < cnt ? wgt[ul[q]] : 0.0;
        }
#ifndef JB_EMU
        if (lane == 0) jb_mbar_expect_tx(&bars[s_], (uint32_t)((size_t)n * cnt * PH * sizeof(cd)));
        __syncwarp();
        for (int c = lane; c < n * cnt; c += 32) {
            int i = c / cnt, q = c - i * cnt;
            jb_bulk_load_1d(dst + ((size_t)i * Q + q) * PH, bb + (((size_t)i * n + jl[q]) * w + ul[q]) * PH,
                            (uint32_t)(PH * sizeof(cd)), &bars[s_]);
        }
#else
        for (int c = lane; c < n * cnt; c += 32) {
            int i = c / cnt, q = c - i * cnt;
            const cd* srow = bb + (((size_t)i * n + jl[q]) * w + ul[q]) * PH;
            cd* drow = dst + ((size_t)i * Q + q) * PH;
            for (int k = 0; k < PH; ++k) drow[k] = srow[k];
        }
#endif
    };
    // number of items / steps of this CTA
    const long my_items = (nitems > (long)blockIdx.x) ? (nitems - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long my_steps = my_items * T;
#ifndef JB_EMU
    if (tid == 0) {
        jb_mbar_init(&bars[0], 1);
        jb_mbar_init(&bars[1], 1);
        jb_mbar_init(&bars[2], JB_RP_THREADS / 32);
        jb_mbar_init(&bars[3], JB_RP_THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) {
        if (my_steps > 0) stage_rows(0, 0);
        if (my_steps > 1) stage_rows(1, 1);
    }
#endif
    long ls = 0;
    for (long it = 0; it < my_items; ++it) {
        const long item = blockIdx.x + it * (long)gridDim.x;
        const int lp = (int)(item / N), r = (int)(item - (long)lp * N);
        int s_i, e_i;
        const int se = jb_map_se(g, lp, s_i, e_i);
        // pass A, accumulated over the modules t: each step consumes one ring slot
        cd u[R2];  // valid for w = tid < n*H1 (one item per thread; plan creation guarantees n*H1 <= threads)
        static_for<0, R2>([&](auto kk) { constexpr int k2 = kk; u[k2] = cd_make(0.0, 0.0); });
        const int wi = tid;
        const bool actA = wi < n * H1;
        const int ia = wi / H1, k1a = wi - ia * H1;
        for (int t = 0; t < T; ++t, ++ls) {
            const int s = (int)(ls & 1);
            const cd* spec = spec0 + (size_t)s * stage_cd;
#ifndef JB_EMU
            jb_mbar_wait<false>(&bars[s], (unsigned)((ls >> 1) & 1));
#else
            JB_SYNC();
            if (tid < 32) stage_rows(ls, s);
            JB_SYNC();
#endif
            if (actA) {
                const int cnt = info[s].cnt;
                const cd* pd[JB_MAXQ];
                const cd* pm[JB_MAXQ];
                double wr[JB_MAXQ];
                static_for<0, JB_MAXQ>([&](auto qq) {
                    constexpr int q = qq;
                    const cd* row = spec + ((size_t)ia * Q + (q < Q ? q : 0)) * PH;
                    pd[q] = row + k1a;
                    pm[q] = row + (L - k1a);
                    wr[q] = info[s].wr[q];
                });
                static_for<0, R2>([&](auto kk) {
                    constexpr int k2 = kk;
                    cd v = jb_sum_spec<FP, k2, true>(pd, pm, wr, cnt, k1a);
                    u[k2].x += v.x;
                    u[k2].y += v.y;
                });
            }
#ifndef JB_EMU
            jb_fence_proxy_async();
            __syncwarp();
            if ((tid & 31) == 0) jb_mbar_arrive(&bars[2 + s]);
            if (tid < 32 && ls + 2 < my_steps) {
                jb_mbar_wait<false>(&bars[2 + s], (unsigned)((ls >> 1) & 1));
                stage_rows(ls + 2, s);
            }
#endif
        }
        if (actA) {
            cd w1 = cd_conj(jb_ldg(tw_kn + R1 + k1a)), w2 = cd_conj(jb_ldg(tw_kn + 2 * R1 + k1a));
            jb_c2r_passA<FP>(u, sm + (size_t)ia * LS + k1a * R2P, w1, w2);
        }
        JB_SYNC();
        for (int wv = tid; wv < n * R2; wv += JB_RP_THREADS) {
            int i = wv / R2, n2 = wv - i * R2;
            double xv[R1];
            jb_c2r_passB<FP>(sm + (size_t)i * LS + n2, xv);
            double* tp = tile + (size_t)i * w + (n2 - m);
            const double* pw = wgt + (n2 - m);
            static_for<0, R1>([&](auto nn) {
                constexpr int n1 = nn;
                int q = R2 * n1 + n2 - m;
                if (q >= 0 && q < w) tp[R2 * n1] = xv[n1] * jb_ldg(pw + R2 * n1);
            });
        }
        JB_SYNC();
        double* grow = grad + ((size_t)se * N + r) * N;
        for (int c = tid; c < N; c += JB_RP_THREADS) {
            double acc = 0.0;
            const double* trow = tile + c + hp;
            const int i1 = chi[c];
            for (int i = clo[c]; i < i1; ++i) acc += trow[i * w - g.offs[i]];
            grow[c] = acc;
        }
    }
}
