// rows3.cuh -- the two c2r row kernels on the three-stage complex transform (cols3.cuh's factorisation).
//
// The two-pass real transforms of rows_pipe.cuh have only n*13 / n*16 heavy work items per image row: with one
// persistent CTA per SM that is 3-4 busy warps and the FP64 latency is fully exposed.  Here the n real output
// rows of one image row are packed pairwise into NC = ceil(n/2) complex sequences
//     Z = X_a + i X_b   (Hermitian half spectra X_a, X_b extended to all L frequencies)
// and inverse-transformed with the three light stages (radix R3, R2, R1) -- R1*R2*NC, R1*R3*NC, M*NC items
// per stage (192/256/192 for n = 8, L = 384 = 8*6*8), real part = row a, imaginary part = row b.  Exact.
// The spectrum rows arrive through the same 2-stage cp.async.bulk ring as in rows_pipe.cuh; the stage that
// assembles Z (sum over row-patches / modules, Hermitian extension, packing) is spread over all threads.
//
// Tile layout: element (pos, c), pos = k1*M + R3*k2 + k3 for frequency k = k1 + R1*(k2 + R2*k3), stored at
// k1*BP + (pos - k1*M)*NC + c with BP = M*NC + 1 (one complex of skew per k1 block keeps the assembly stores,
// which walk k1 fastest, conflict-free).
#pragma once
#include "rows_pipe.cuh"
#include "cols3.cuh"

#define JB_R4_THREADS 256
#define JB_R4_CTAS_PER_SM 2   // two CTAs per SM, each with a single-slot ring: the loads of one overlap the other's stages
#define JB_R4_MAXEL 8   // tile elements assembled per thread: L*NC <= 8*threads

template <class F3>
struct Rows3 {
    static constexpr int L = F3::L, R1 = F3::R1, R2 = F3::R2, R3 = F3::R3, M = F3::M;

    // inverse stage 3 (radix R3 over k3), in place.  base = k1*BP + (R3*k2)*NC + b
    static JB_D void s3_inv(cd* tile, int base, int NC) {
        cd v[R3];
        static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; v[k3] = tile[base + k3 * NC]; });
        dft<R3, +1>(v);
        static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; tile[base + n3 * NC] = v[n3]; });
    }
    // inverse stage 2 (conj twiddle, radix R2 over k2), in place.  base = k1*BP + n3*NC + b
    static JB_D void s2_inv(cd* tile, const cd* tw2, int base, int n3, int NC) {
        cd v[R2];
        static_for<0, R2>([&](auto kk) {
            constexpr int k2 = kk;
            cd y = tile[base + (R3 * k2) * NC];
            if constexpr (k2 > 0) y = cd_mulc(y, tw2[k2 * R3 + n3]);
            v[k2] = y;
        });
        dft<R2, +1>(v);
        static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; tile[base + (R3 * n2) * NC] = v[n2]; });
    }
    // inverse stage 1 (conj twiddle, radix R1 over k1): v[n1] = z[M*n1 + m].  base = m*NC + b
    static JB_D void s1_inv(const cd* tile, const cd* tw1, int base, int m, int BP, cd (&v)[R1]) {
        static_for<0, R1>([&](auto kk) {
            constexpr int k1 = kk;
            cd y = tile[base + k1 * BP];
            if constexpr (k1 > 0) y = cd_mulc(y, tw1[k1 * M + m]);
            v[k1] = y;
        });
        dft<R1, +1>(v);
    }
    static JB_HD int pos_addr(int k, int c, int NC, int BP) {
        int k1 = k % R1, q = k / R1, k2 = q % R2, k3 = q / R2;
        return k1 * BP + (R3 * k2 + k3) * NC + c;
    }
};

// per-thread assembly descriptors (the element <-> thread map is the same for every image row)
struct JbAsm {
    int addr[JB_R4_MAXEL];   // tile address, -1 = none
    int src[JB_R4_MAXEL];    // offset of the half-spectrum element of sequence a inside a ring slot: ia*Q*PH + min(k, L-k)
    unsigned mir;            // bit e: frequency k > L/2 (use the conjugate)
    unsigned hasb;           // bit e: the pair has a second sequence (ia+1 < n)
};
#define JB_R4_MAXC 4  // image columns per thread with precomputed gather offsets (N <= 4*threads)

template <class F3>
JB_D void jb_asm_build(JbAsm& a, int n, int NC, int BP, int QPH) {
    a.mir = 0; a.hasb = 0;
    static_for<0, JB_R4_MAXEL>([&](auto ee) {
        constexpr int e = ee;
        int idx = (int)threadIdx.x + e * JB_R4_THREADS;
        a.addr[e] = -1; a.src[e] = 0;
        if (idx < F3::L * NC) {
            int c = idx / F3::L, k = idx - c * F3::L;
            a.addr[e] = Rows3<F3>::pos_addr(k, c, NC, BP);
            int kk = k;
            if (2 * k > F3::L) { kk = F3::L - k; a.mir |= 1u << e; }
            a.src[e] = 2 * c * QPH + kk;
            if (2 * c + 1 < n) a.hasb |= 1u << e;
        }
    });
}

// per-thread gather offsets into the real output tile: column c = tid + k*threads collects tr[i*len + c+shift-offs[i]]
// from the (<= JB_MAXQ) column blocks i that cover it; -1 = unused
JB_D void jb_gather_build(const JbGeom& g, int shift, int len, int (&goff)[JB_R4_MAXC][JB_MAXQ]) {
    static_for<0, JB_R4_MAXC>([&](auto kk) {
        constexpr int k = kk;
        int c = (int)threadIdx.x + k * JB_R4_THREADS, cntc = 0;
        static_for<0, JB_MAXQ>([&](auto jj) { constexpr int j = jj; goff[k][j] = -1; });
        if (c < g.N) {
            for (int i = 0; i < g.n; ++i) {
                int p = c + shift - g.offs[i];
                if (p >= 0 && p < len) {
                    int o = i * len + p;
                    static_for<0, JB_MAXQ>([&](auto jj) { constexpr int j = jj; if (cntc == j) goff[k][j] = o; });
                    ++cntc;
                }
            }
        }
    });
}
JB_D double jb_gather(const double* tr, const int (&go)[JB_MAXQ]) {
    double acc = 0.0;
    static_for<0, JB_MAXQ>([&](auto jj) { constexpr int j = jj; if (go[j] >= 0) acc += tr[go[j]]; });
    return acc;
}

// ------------------------------------------------------------------------------------------------------
// forward: spectra -> image rows (+ epilogue; not JB_EPI_POISSON).  grid = #SMs, block = JB_R4_THREADS.
template <class FP, class F3>
JB_GLOBAL void JB_LB(JB_R4_THREADS, JB_R4_CTAS_PER_SM)
k_rows_out4(JbGeom g, const cd* __restrict__ bufA, const cd* __restrict__ tw1g, const cd* __restrict__ tw2g, JbEpi ep,
            int planes, int Q) {
    using R = Rows3<F3>;
    constexpr int L = F3::L, R1 = F3::R1, R2 = F3::R2, R3 = F3::R3, M = F3::M;
    JB_SMEM(unsigned char, smraw);
    const int n = g.n, N = g.N, m = g.m, hp = g.hp, PH = g.PH, P = g.P;
    const int NC = (n + 1) / 2, BP = M * NC + 1;
    const int tid = threadIdx.x;
    const size_t stage_cd = (size_t)n * Q * PH;
    size_t off = 0;
    cd* spec0 = reinterpret_cast<cd*>(smraw);               // single ring slot
    off += jb_al128(stage_cd * sizeof(cd));
    cd* tile = reinterpret_cast<cd*>(smraw + off);
    double* tr = reinterpret_cast<double*>(smraw + off);   // real output rows [2*NC][P], ALIASES the tile
    off += jb_al128((size_t)R1 * BP * sizeof(cd));
    cd* tw1 = reinterpret_cast<cd*>(smraw + off);
    off += jb_al128((size_t)(L + M) * sizeof(cd));
    cd* tw2 = tw1 + L;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + off);
    off += 64;
    JbRowInfo* info = reinterpret_cast<JbRowInfo*>(smraw + off);
    off += 2 * sizeof(JbRowInfo);
    unsigned char* clo = smraw + off;
    unsigned char* chi = clo + N;
    const long nitems = (long)planes * N;
    const size_t nn2 = (size_t)n * n;
    const int shift = hp + m;
    for (int i = tid; i < L; i += JB_R4_THREADS) tw1[i] = tw1g[i];
    for (int i = tid; i < M; i += JB_R4_THREADS) tw2[i] = tw2g[i];
    for (int c = tid; c < N; c += JB_R4_THREADS) {
        int l = n, h = -1;
        for (int i = 0; i < n; ++i) {
            int p = c + shift - g.offs[i];
            if (p >= 0 && p < P) { if (i < l) l = i; h = i; }
        }
        clo[c] = (unsigned char)l;
        chi[c] = (unsigned char)(h + 1);
    }
    JbAsm as;
    const int QPH = Q * PH;
    jb_asm_build<F3>(as, n, NC, BP, QPH);
    int goff[JB_R4_MAXC][JB_MAXQ];
    jb_gather_build(g, shift, P, goff);
    // thread-constant stage indices
    const int I3 = R1 * R2 * NC, I2 = R1 * R3 * NC, I1 = M * NC;
    int base3, base2, n3_2, base1, m_1, b_1;
    {
        int c3 = tid / NC, b = tid - c3 * NC, k1 = c3 / R2, k2 = c3 - k1 * R2;
        base3 = k1 * BP + (R3 * k2) * NC + b;
        int c2 = tid / NC, k1b = c2 / R3;
        n3_2 = c2 - k1b * R3;
        base2 = k1b * BP + n3_2 * NC + b;
        m_1 = tid / NC; b_1 = b;
        base1 = m_1 * NC + b;
    }

    auto stage_rows = [&](long item_, int s_) {   // s_ is always 0 (single slot)
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
        jb_mbar_init(&bars[2], JB_R4_THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32 && (long)blockIdx.x < nitems) stage_rows(blockIdx.x, 0);
#else
    JB_SYNC();
#endif
    long item = blockIdx.x;
    for (unsigned it = 0; item < nitems; ++it, item += gridDim.x) {
        const int s = 0;
        const int lp = (int)(item / N), r = (int)(item - (long)lp * N);
        int ste, te, se_;
        jb_map_ste(g, lp, ste, te, se_);
        const cd* spec = spec0;
#ifndef JB_EMU
        jb_mbar_wait<false>(&bars[0], it & 1);
#else
        JB_SYNC();
        if (tid < 32) stage_rows(item, s);
        JB_SYNC();
#endif
        const int cnt = info[s].cnt;
        // assemble Z = X_a + i X_b (summed over row-patches, Hermitian-extended) into the tile
        static_for<0, JB_R4_MAXEL>([&](auto ee) {
            constexpr int e = ee;
            if (as.addr[e] >= 0) {
                const cd* ra = spec + as.src[e];
                const bool hasb = (as.hasb >> e) & 1u;
                const cd* rb = ra + (hasb ? QPH : 0);
                cd xa = cd_make(0.0, 0.0), xb = xa;
                static_for<0, JB_MAXQ>([&](auto qq) {
                    constexpr int q = qq;
                    if (q < cnt) {
                        xa = cd_add(xa, ra[q * PH]);
                        xb = cd_add(xb, rb[q * PH]);
                    }
                });
                if (!hasb) xb = cd_make(0.0, 0.0);
                if (as.mir & (1u << e)) { xa.y = -xa.y; xb.y = -xb.y; }
                tile[as.addr[e]] = cd_make(xa.x - xb.y, xa.y + xb.x);
            }
        });
#ifndef JB_EMU
        jb_fence_proxy_async();
        __syncwarp();
        if ((tid & 31) == 0) jb_mbar_arrive(&bars[2]);
        if (tid < 32) {
            long nxt = item + (long)gridDim.x;
            if (nxt < nitems) {
                jb_mbar_wait<false>(&bars[2], it & 1);
                stage_rows(nxt, 0);   // lands while this row runs its three stages and the gather
            }
        }
#endif
        JB_SYNC();
        if (tid < I3) R::s3_inv(tile, base3, NC);
        JB_SYNC();
        if (tid < I2) R::s2_inv(tile, tw2, base2, n3_2, NC);
        JB_SYNC();
        {
            cd v[R1];
            if (tid < I1) R::s1_inv(tile, tw1, base1, m_1, BP, v);
            JB_SYNC();   // the real rows overwrite the tile
            if (tid < I1) {
                double* ta = tr + (size_t)(2 * b_1) * P + m_1;
                static_for<0, R1>([&](auto nn) {
                    constexpr int n1 = nn;
                    if (M * n1 + m_1 < P) { ta[M * n1] = v[n1].x; ta[P + M * n1] = v[n1].y; }
                });
            }
        }
        JB_SYNC();
        const size_t orow = ((size_t)ste * N + r) * N, erow = ((size_t)te * N + r) * N;
        static_for<0, JB_R4_MAXC>([&](auto kk) {
            constexpr int kc = kk;
            const int c = tid + kc * JB_R4_THREADS;
            if (c < N) {
                double conv = jb_gather(tr, goff[kc]);
                if (ep.mode == JB_EPI_PSF) {
                    ep.out[orow + c] = conv;
                } else {
                    double ex = jb_ldg(ep.expo + erow + c);
                    if (ep.mode == JB_EPI_RESP) {
                        ep.out[orow + c] = (ex != 0.0) ? ex * conv : 0.0;
                    } else {
                        double gout = 0.0;
                        if (ex != 0.0) gout = ex * (ex * conv) / ep.lam[orow + c];
                        ep.out[orow + c] = gout;
                    }
                }
            }
        });
        JB_SYNC();   // the gather has read the aliased tile before the next row is assembled into it
    }
}

// ------------------------------------------------------------------------------------------------------
// adjoint: spectra -> sky-gradient rows, summed over modules.  grid = #SMs, block = JB_R4_THREADS.
template <class FP, class F3>
JB_GLOBAL void JB_LB(JB_R4_THREADS, JB_R4_CTAS_PER_SM)
k_rows_adj_out4(JbGeom g, const cd* __restrict__ bufB, const double* __restrict__ wgt, const cd* __restrict__ tw1g,
                const cd* __restrict__ tw2g, double* __restrict__ grad, int planes_se, int Q) {
    using R = Rows3<F3>;
    constexpr int L = F3::L, R1 = F3::R1, R2 = F3::R2, R3 = F3::R3, M = F3::M;
    JB_SMEM(unsigned char, smraw);
    const int n = g.n, N = g.N, m = g.m, hp = g.hp, PH = g.PH, w = g.w, T = g.T;
    const int NC = (n + 1) / 2, BP = M * NC + 1;
    const int tid = threadIdx.x;
    const size_t stage_cd = (size_t)n * Q * PH;
    size_t off = 0;
    cd* spec0 = reinterpret_cast<cd*>(smraw);               // single ring slot
    off += jb_al128(stage_cd * sizeof(cd));
    cd* tile = reinterpret_cast<cd*>(smraw + off);
    double* tr = reinterpret_cast<double*>(smraw + off);   // cropped, weighted real rows [2*NC][w], ALIASES the tile
    off += jb_al128((size_t)R1 * BP * sizeof(cd));
    cd* tw1 = reinterpret_cast<cd*>(smraw + off);
    off += jb_al128((size_t)(L + M) * sizeof(cd));
    cd* tw2 = tw1 + L;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + off);
    off += 64;
    JbRowInfo* info = reinterpret_cast<JbRowInfo*>(smraw + off);
    off += 2 * sizeof(JbRowInfo);
    unsigned char* clo = smraw + off;
    unsigned char* chi = clo + N;
    const long nitems = (long)planes_se * N;
    const size_t nn2 = (size_t)n * n;
    const size_t plane_sz = nn2 * w * PH;
    for (int i = tid; i < L; i += JB_R4_THREADS) tw1[i] = tw1g[i];
    for (int i = tid; i < M; i += JB_R4_THREADS) tw2[i] = tw2g[i];
    for (int c = tid; c < N; c += JB_R4_THREADS) {
        int l = n, h = -1;
        for (int i = 0; i < n; ++i) {
            int v = c + hp - g.offs[i];
            if (v >= 0 && v < w) { if (i < l) l = i; h = i; }
        }
        clo[c] = (unsigned char)l;
        chi[c] = (unsigned char)(h + 1);
    }
    JbAsm as;
    const int QPH = Q * PH;
    jb_asm_build<F3>(as, n, NC, BP, QPH);
    int goff[JB_R4_MAXC][JB_MAXQ];
    jb_gather_build(g, hp, w, goff);
    const int I3 = R1 * R2 * NC, I2 = R1 * R3 * NC, I1 = M * NC;
    int base3, base2, n3_2, base1, m_1, b_1;
    {
        int c3 = tid / NC, b = tid - c3 * NC, k1 = c3 / R2, k2 = c3 - k1 * R2;
        base3 = k1 * BP + (R3 * k2) * NC + b;
        int c2 = tid / NC, k1b = c2 / R3;
        n3_2 = c2 - k1b * R3;
        base2 = k1b * BP + n3_2 * NC + b;
        m_1 = tid / NC; b_1 = b;
        base1 = m_1 * NC + b;
    }
    // local step ls = it*T + t  <->  (item = blockIdx.x + it*gridDim.x, module t)
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
    const long my_items = (nitems > (long)blockIdx.x) ? (nitems - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long my_steps = my_items * T;
#ifndef JB_EMU
    if (tid == 0) {
        jb_mbar_init(&bars[0], 1);
        jb_mbar_init(&bars[2], JB_R4_THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32 && my_steps > 0) stage_rows(0, 0);
#else
    JB_SYNC();
#endif
    long ls = 0;
    for (long it = 0; it < my_items; ++it) {
        const long item = blockIdx.x + it * (long)gridDim.x;
        const int lp = (int)(item / N), r = (int)(item - (long)lp * N);
        int s_i, e_i;
        const int se = jb_map_se(g, lp, s_i, e_i);
        // assemble Z = sum_t sum_q w_q (X_a + i X_b); each thread owns the same tile elements for every module
        for (int t = 0; t < T; ++t, ++ls) {
            const int s = 0;
            const cd* spec = spec0;
#ifndef JB_EMU
            jb_mbar_wait<false>(&bars[0], (unsigned)(ls & 1));
#else
            JB_SYNC();
            if (tid < 32) stage_rows(ls, s);
            JB_SYNC();
#endif
            const int cnt = info[s].cnt;
            double wr[JB_MAXQ];
            static_for<0, JB_MAXQ>([&](auto qq) { constexpr int q = qq; wr[q] = info[s].wr[q]; });
            static_for<0, JB_R4_MAXEL>([&](auto ee) {
                constexpr int e = ee;
                if (as.addr[e] >= 0) {
                    const cd* ra = spec + as.src[e];
                    const bool hasb = (as.hasb >> e) & 1u;
                    const cd* rb = ra + (hasb ? QPH : 0);
                    cd xa = cd_make(0.0, 0.0), xb = xa;
                    static_for<0, JB_MAXQ>([&](auto qq) {
                        constexpr int q = qq;
                        if (q < cnt) {
                            cd va = ra[q * PH], vb = rb[q * PH];
                            xa.x += wr[q] * va.x; xa.y += wr[q] * va.y;
                            xb.x += wr[q] * vb.x; xb.y += wr[q] * vb.y;
                        }
                    });
                    if (!hasb) xb = cd_make(0.0, 0.0);
                    if (as.mir & (1u << e)) { xa.y = -xa.y; xb.y = -xb.y; }
                    cd z = cd_make(xa.x - xb.y, xa.y + xb.x);
                    if (t > 0) z = cd_add(z, tile[as.addr[e]]);
                    tile[as.addr[e]] = z;
                }
            });
#ifndef JB_EMU
            jb_fence_proxy_async();
            __syncwarp();
            if ((tid & 31) == 0) jb_mbar_arrive(&bars[2]);
            if (tid < 32 && ls + 1 < my_steps) {
                jb_mbar_wait<false>(&bars[2], (unsigned)(ls & 1));
                stage_rows(ls + 1, 0);
            }
#endif
        }
        JB_SYNC();
        if (tid < I3) R::s3_inv(tile, base3, NC);
        JB_SYNC();
        if (tid < I2) R::s2_inv(tile, tw2, base2, n3_2, NC);
        JB_SYNC();
        {
            cd v[R1];
            if (tid < I1) R::s1_inv(tile, tw1, base1, m_1, BP, v);
            JB_SYNC();   // the real rows overwrite the tile
            if (tid < I1) {
                double* ta = tr + (size_t)(2 * b_1) * w + (m_1 - m);
                const double* pw = wgt + (m_1 - m);
                static_for<0, R1>([&](auto nn) {
                    constexpr int n1 = nn;
                    int q = M * n1 + m_1 - m;
                    if (q >= 0 && q < w) {
                        double ww = jb_ldg(pw + M * n1);
                        ta[M * n1] = v[n1].x * ww;
                        ta[w + M * n1] = v[n1].y * ww;
                    }
                });
            }
        }
        JB_SYNC();
        double* grow = grad + ((size_t)se * N + r) * N;
        static_for<0, JB_R4_MAXC>([&](auto kk) {
            constexpr int kc = kk;
            const int c = tid + kc * JB_R4_THREADS;
            if (c < N) grow[c] = jb_gather(tr, goff[kc]);
        });
        JB_SYNC();   // the gather has read the aliased tile before the next row is assembled into it
    }
}
