// ops.cuh -- per-FFT-length launch table.  Each supported transform length is instantiated in its own
// translation unit (fft_L*.cu, one line each) so the lengths compile in parallel; capi.cu picks the
// table by length at plan creation.
#pragma once
#include "fft_list.h"
#include "kernels.cuh"
#include "rows_fast.cuh"
#include "cols_pipe.cuh"
#include "cols3.cuh"
#include "rows5.cuh"
#include "cols_wp.cuh"

#ifndef JB_COLS_WP
#define JB_COLS_WP 0   // 1: the column pass runs as warp-private transforms (cols_wp.cuh) instead of the TMA tile pipeline
#endif
#ifndef JB_COLS_WP_NW
#define JB_COLS_WP_NW 8
#endif
#ifndef JB_COLS_MT
#define JB_COLS_MT 1   // forward column pass shares the forward transform between the modules of a (sample, bin, patch)
#endif


struct JbFftOps {
    int L, R1, R2, CB;
    int ls;    // cd per real sequence in the exchange buffer
    int pipe;  // 0: plain column kernel only; 2: three-stage TMA pipeline (cols3.cuh)
    int q1, q2, q3;  // three-stage factors of the column pass (pipe == 2): twa/twb of cols_pipe / prep_cols
    int rq1, rq2, rq3;  // three-stage factors of the c2r row kernels (rows5_*): their own twiddle tables (fft_list.h)
    // fast != 0 selects the low-overhead kernels of rows_fast.cuh (two-pass lengths, <= JB_MAXQ row-patches per row,
    // no alias fold); fast == 0 the generic kernels of kernels.cuh (any geometry, period-P alias fold)
    cudaError_t (*rows_fwd)(const JbGeom&, const double* x, const double* wgt, const cd* tw_nk, cd* bufB,
                            int planes, int rpc, int threads, int fast, int exp_in, jb_stream_t);
    cudaError_t (*rows_adj_in)(const JbGeom&, const double* gin, const double* expo, const cd* tw_nk, cd* bufA,
                               int planes, int rpc, int threads, int fast, jb_stream_t);
    cudaError_t (*cols)(const JbGeom&, bool adj, const cd* in, cd* out, const cd* khat, const cd* tw_nk,
                        const double* wgt, int planes, int threads, jb_stream_t);
    // three-stage TMA pipeline (pipe == 2): twa/twb = stage-1/stage-2 twiddle tables
    // mt != 0 (several modules): forward -- the forward transform is shared by the modules; adjoint -- the modules are
    // summed in the spectral domain (accumulator tiles in `scratch`: cols_cps * ctas * L * CB elements) and `out` is the
    // module-summed layout (S*E, n*n, w, PH) (cols3.cuh, k_cols3_*_mt)
    cudaError_t (*cols_pipe)(const JbGeom&, bool adj, const JbTileSrc& src, cd* out, const cd* khat,
                             const cd* twa, const cd* twb, const double* wgt, int planes, int ctas, int mt, cd* scratch,
                             jb_stream_t);
    int cols_mt;   // the *_mt column kernels are compiled in
    int cols_cps;  // CTAs per SM of the column pipeline
    cudaError_t (*rows_out)(const JbGeom&, const cd* bufA, const cd* tw_kn, const JbEpi&, int planes, int rpc,
                            int threads, int fast, jb_stream_t);
    cudaError_t (*rows_adj_out)(const JbGeom&, const cd* bufB, const double* wgt, const cd* tw_kn, double* grad,
                                const double* logsky, int planes, int rpc, int threads, int fast, jb_stream_t);
    // dynamic shared memory of the generic row kernels (plan creation rejects geometries beyond 227 KB)
    size_t (*rows_generic_smem)(const JbGeom&, int rpc, int threads);
    // warp-private c2r kernels (rows5.cuh; need pipe == 2); fwd: ep.mode may be JB_EPI_POISSON (partials [STE][N])
    size_t (*rows5_smem)(const JbGeom&);
    int (*rows5_gather)(const JbGeom&, int adj, int maxq, int* host_tab);   // [maxq][N] table, returns slots used
    cudaError_t (*rows5_fwd)(const JbGeom&, const cd* bufA, const cd* tw1, const cd* tw2, const int* gtab, int gq,
                             const JbEpi&, int planes, int n_sm, jb_stream_t);
    cudaError_t (*rows5_adj)(const JbGeom&, const cd* bufB, const double* wgt, const cd* tw1, const cd* tw2,
                             const int* gtab, int gq, double* grad, const double* logsky, int planes_se, int n_sm,
                             jb_stream_t);
    // stand-alone FFT convolution (jubik_convolve): rows of length L, columns of length L
    cudaError_t (*conv_rows_r2c)(const double* in, cd* out, const cd* tw_nk, long nrows, jb_stream_t);
    cudaError_t (*conv_rows_c2r)(const cd* in, double* out, const cd* tw_kn, long nrows, double scale, jb_stream_t);
    cudaError_t (*conv_cols)(const cd* xs, cd* ys, cd* scratch, const cd* tw_nk, int PH, int lead, int x_lead, jb_stream_t);
    cudaError_t (*prep_rows)(const JbGeom&, const JbKsrc&, const cd* tw_nk, cd* scratch, int rpc, int threads,
                             jb_stream_t);
    // layout 0: K^ in the two-pass register order (tw = tw_nk); layout 2: three-stage order (twa/twb = tw1/tw2)
    cudaError_t (*prep_cols)(const JbGeom&, const cd* scratch, const double* sums, const cd* twa, const cd* twb,
                             cd* khat, int layout, int threads, jb_stream_t);
};

inline int jb_cdiv(int a, int b) { return (a + b - 1) / b; }
inline size_t jb_al16(size_t x) { return (x + 15) & ~(size_t)15; }

template <class FP, int CB, int PIPE, class F3, class F3R>
struct JbOpsImpl {
    static cudaError_t cols_pipe(const JbGeom& g, bool adj, const JbTileSrc& src, cd* out, const cd* khat,
                                 const cd* twa, const cd* twb, const double* wgt, int planes, int ctas, int mt, cd* scratch,
                                 jb_stream_t st) {
        if constexpr (PIPE == 2 && JB_COLS_WP) {
            constexpr int NW = JB_COLS_WP_NW;
            const size_t smem = jb_cols_wp_smem<F3>(g.w, NW);
            auto go = [&](auto kern) {
                int per_sm = 1;
#ifndef JB_EMU
                if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
#endif
                long grid = (long)per_sm * ctas;
                const long units = (long)planes * g.n * g.n * g.PH, need = (units + NW - 1) / NW;
                if (grid > need) grid = need;
                return jb_launch(kern, dim3((unsigned)grid), dim3(NW * 32), smem, st, g, src.base, out, khat, twa, twb, wgt, planes);
            };
            return adj ? go(k_cols_wp<F3, true, NW>) : go(k_cols_wp<F3, false, NW>);
        } else if constexpr (PIPE == 2) {
            size_t smem = 4 * (size_t)FP::L * CB * sizeof(cd) + (size_t)(F3::L + F3::M) * sizeof(cd) + 32 +
                          2 * sizeof(JbColInfo) + (size_t)g.w * sizeof(double) + 32 + 128;   // + the tile shift for odd margins
            const dim3 grid(ctas * Cols3<F3, CB>::CPS), block(Cols3<F3, CB>::THREADS);
#if JB_COLS_MT
            if (mt && g.T > 1 && planes % g.T == 0) {
                if (!adj) return jb_launch(k_cols3_fwd_mt<F3, CB>, grid, block, smem, st, src, g, out, khat, twa, twb, wgt, planes / g.T);
                if (scratch) return jb_launch(k_cols3_adj_mt<F3, CB>, grid, block, smem, st, src, g, out, khat, twa, twb, scratch, planes / g.T);
            }
#endif
            if (adj) return jb_launch(k_cols3_pipe<F3, CB, true>, grid, block, smem, st, src, g, out, khat, twa, twb, wgt,
                                      planes);
            return jb_launch(k_cols3_pipe<F3, CB, false>, grid, block, smem, st, src, g, out, khat, twa, twb, wgt, planes);
        } else {
            (void)g; (void)adj; (void)src; (void)out; (void)khat; (void)twa; (void)twb; (void)wgt;
            (void)planes; (void)ctas; (void)mt; (void)scratch; (void)st;
            return (cudaError_t)1;
        }
    }
    static size_t rowtab_bytes(int rpc) {
        return jb_al16(sizeof(int) * rpc) + (sizeof(long long) + sizeof(double)) * (size_t)rpc * JB_MAXQ + 16;
    }
    static size_t smem_rows_in(const JbGeom& g, int rpc) {
        size_t B = (size_t)rpc * g.n;
        return jb_al16(B * FP::LS * sizeof(cd)) + 16;
    }
    // generic c2r kernels: exchange buffer + real tile (TW wide) + reduction scratch + row-patch lists (2n per row)
    static size_t smem_rows_out_generic(const JbGeom& g, int rpc, int threads) {
        size_t B = (size_t)rpc * g.n;
        const int TW = g.fold ? g.L : g.P;
        return jb_al16(B * FP::LS * sizeof(cd)) + jb_al16(B * TW * sizeof(double)) + jb_al16((size_t)threads * sizeof(double)) +
               sizeof(int) * (rpc + 4 * (size_t)rpc * g.n) + 16;
    }
    static size_t smem_rows_adj_out_generic(const JbGeom& g, int rpc) {
        size_t B = (size_t)rpc * g.n;
        return jb_al16(B * FP::LS * sizeof(cd)) + jb_al16(B * g.w * sizeof(double)) + sizeof(int) * (rpc + 4 * (size_t)rpc * g.n) + 16;
    }
    static size_t rows_generic_smem(const JbGeom& g, int rpc, int threads) {
        size_t a = smem_rows_out_generic(g, rpc, threads), b = smem_rows_adj_out_generic(g, rpc), c = smem_rows_in(g, rpc);
        return a > b ? (a > c ? a : c) : (b > c ? b : c);
    }
    // persistent prefetching r2c kernel (k_rows_in_pf): used when its exchange buffer + row ring fit an SM
    template <bool ADJ>
    static bool rows_in_pf(const JbGeom& g, const double* in, const double* aux, const cd* tw_nk, cd* out, int planes,
                           int threads, int exp_in, jb_stream_t st, cudaError_t* err) {
        if constexpr (FP::R2 > 1) {
            const size_t smem = jb_al16((size_t)g.n * FP::LS * sizeof(cd)) + 2 * (size_t)g.N * sizeof(double) +
                                (ADJ ? (aux ? 2 * (size_t)g.N * sizeof(double) : 0) : (size_t)g.w * sizeof(double));
            const size_t lim = 200 * 1024;   // one CTA per SM is still a win (C4: cotangent rows 0.41 -> 0.16 ms, r1x)
            if ((g.N & 1) || smem > lim) return false;
            int per_sm = 1, n_sm = 148;
#ifndef JB_EMU
            static int cache_thr = -1, cache_cps = 1, cache_sm = 148;
            static size_t cache_smem = 0;
            if (cache_thr != threads || cache_smem != smem) {
                int dev = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&cache_sm, cudaDevAttrMultiProcessorCount, dev);
                if (smem > 48 * 1024)
                    cudaFuncSetAttribute(k_rows_in_pf<FP, ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cache_cps, k_rows_in_pf<FP, ADJ>, threads, smem) != cudaSuccess ||
                    cache_cps < 1) { cudaGetLastError(); cache_cps = 1; }
                cache_thr = threads; cache_smem = smem;
            }
            per_sm = cache_cps; n_sm = cache_sm;
#endif
            long ctas = (long)per_sm * n_sm;
            const long total = (long)planes * g.N;
            if (ctas > total) ctas = total;
            *err = jb_launch(k_rows_in_pf<FP, ADJ>, dim3((unsigned)ctas), dim3(threads), smem, st, g, in, aux, tw_nk, out, planes, exp_in);
            return true;
        } else {
            (void)g; (void)in; (void)aux; (void)tw_nk; (void)out; (void)planes; (void)threads; (void)exp_in; (void)st; (void)err;
            return false;
        }
    }
    static cudaError_t rows_fwd(const JbGeom& g, const double* x, const double* wgt, const cd* tw_nk, cd* bufB,
                                int planes, int rpc, int threads, int fast, int exp_in, jb_stream_t st) {
        dim3 grid(jb_cdiv(g.N, rpc), planes);
        cudaError_t e2;
        if (fast && rows_in_pf<false>(g, x, wgt, tw_nk, bufB, planes, threads, exp_in, st, &e2)) return e2;
        return jb_launch(k_rows_fwd<FP>, grid, dim3(threads), smem_rows_in(g, rpc), st, g, x, wgt, tw_nk, bufB, rpc, exp_in);
    }
    static cudaError_t rows_adj_in(const JbGeom& g, const double* gin, const double* expo, const cd* tw_nk,
                                   cd* bufA, int planes, int rpc, int threads, int fast, jb_stream_t st) {
        dim3 grid(jb_cdiv(g.N, rpc), planes);
        cudaError_t e2;
        if (fast && rows_in_pf<true>(g, gin, expo, tw_nk, bufA, planes, threads, 0, st, &e2)) return e2;
        return jb_launch(k_rows_adj_in<FP>, grid, dim3(threads), smem_rows_in(g, rpc), st, g, gin, expo, tw_nk, bufA, rpc);
    }
    static cudaError_t cols(const JbGeom& g, bool adj, const cd* in, cd* out, const cd* khat, const cd* tw_nk,
                            const double* wgt, int planes, int threads, jb_stream_t st) {
        dim3 grid(jb_cdiv(g.PH, CB), g.n * g.n, planes);
        size_t smem = (FP::R2 == 1) ? 16 : (size_t)FP::L * CB * sizeof(cd);
        if (adj) return jb_launch(k_cols<FP, CB, true>, grid, dim3(threads), smem, st, g, in, out, khat, tw_nk, wgt);
        return jb_launch(k_cols<FP, CB, false>, grid, dim3(threads), smem, st, g, in, out, khat, tw_nk, wgt);
    }
    static cudaError_t rows_out(const JbGeom& g, const cd* bufA, const cd* tw_kn, const JbEpi& ep, int planes,
                                int rpc, int threads, int fast, jb_stream_t st) {
        size_t B = (size_t)rpc * g.n;
        dim3 grid(jb_cdiv(g.N, rpc), planes);
        if constexpr (FP::R2 > 1) {
            if (fast) {
                size_t common = jb_al16(B * FP::LS * sizeof(cd)) + jb_al16(B * g.P * sizeof(double)) +
                                jb_al16((size_t)threads * sizeof(double));
                return jb_launch(k_rows_out2<FP>, grid, dim3(threads), common + rowtab_bytes(rpc), st, g, bufA, tw_kn, ep, rpc);
            }
        }
        return jb_launch(k_rows_out<FP>, grid, dim3(threads), smem_rows_out_generic(g, rpc, threads), st, g, bufA, tw_kn, ep, rpc);
    }
    static cudaError_t rows_adj_out(const JbGeom& g, const cd* bufB, const double* wgt, const cd* tw_kn,
                                    double* grad, const double* logsky, int planes, int rpc, int threads, int fast,
                                    jb_stream_t st) {
        size_t B = (size_t)rpc * g.n;
        dim3 grid(jb_cdiv(g.N, rpc), planes);
        if constexpr (FP::R2 > 1) {
            if (fast) {
                size_t common = jb_al16(B * FP::LS * sizeof(cd)) + jb_al16(B * g.w * sizeof(double));
                return jb_launch(k_rows_adj_out2<FP>, grid, dim3(threads), common + rowtab_bytes(rpc), st, g,
                                 bufB, wgt, tw_kn, grad, logsky, rpc);
            }
        }
        return jb_launch(k_rows_adj_out<FP>, grid, dim3(threads), smem_rows_adj_out_generic(g, rpc), st, g, bufB, wgt, tw_kn,
                         grad, logsky, rpc);
    }
    // sequences per warp of the c2r kernels: one for up to 8 packed pairs (n <= 16; more than 4 warps: the WIDE form)
    static int rows5_sw(const JbGeom& g) { return (g.n + 1) / 2 <= 8 ? 1 : 2; }
    static size_t rows5_smem(const JbGeom& g) {
        if constexpr (PIPE == 2) {
            const int sw = rows5_sw(g), NC = (g.n + 1) / 2, NW = (NC + sw - 1) / sw;
            if (NW * 32 > 512) return (size_t)1 << 30;
            return sw == 1 ? jb_rows5_smem<F3R, 1>(g.n) : jb_rows5_smem<F3R, 2>(g.n);
        } else {
            (void)g;
            return (size_t)1 << 30;
        }
    }
    static int rows5_gather(const JbGeom& g, int adj, int maxq, int* host_tab) {
        if constexpr (PIPE == 2) {
            return rows5_sw(g) == 1 ? jb_rows5_gather_table<F3R, 1>(g, adj != 0, maxq, host_tab)
                                    : jb_rows5_gather_table<F3R, 2>(g, adj != 0, maxq, host_tab);
        } else {
            (void)g; (void)adj; (void)maxq; (void)host_tab;
            return 0;
        }
    }
    template <class K>
    static int rows5_grid(K kern, int threads, size_t smem, long items, int n_sm) {
        int per_sm = 1;
#ifndef JB_EMU
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
#endif
        long gsz = (long)per_sm * n_sm;
        return (int)(gsz < items ? gsz : items);
    }
    template <int SW, bool ADJ>
    static cudaError_t rows5_launch(const JbGeom& g, const cd* src, const double* wgt, const cd* tw1, const cd* tw2,
                                    const int* gtab, int gq, const JbEpi& ep, double* grad, const double* logsky,
                                    int planes, int n_sm, jb_stream_t st) {
        const int NC = (g.n + 1) / 2, NW = (NC + SW - 1) / SW;
        const size_t smem = jb_rows5_smem<F3R, SW>(g.n);
        auto go = [&](auto kern) {
            int grid = rows5_grid(kern, NW * 32, smem, (long)planes * g.N, n_sm);
            return jb_launch(kern, dim3(grid), dim3(NW * 32), smem, st, g, src, tw1, tw2, wgt, gtab, gq, ep, grad, logsky, planes);
        };
        // heavy = transcendental epilogue: forward Poisson energy/cotangent, adjoint * exp(log_sky)
        const bool heavy = ADJ ? (logsky != nullptr) : (ep.mode == JB_EPI_POISSON);
        if constexpr (SW == 1) {
            if (NW > 4)   // 5..8 warps per CTA
                return heavy ? go(k_rows_c2r5<F3R, 1, ADJ, true, true>) : go(k_rows_c2r5<F3R, 1, ADJ, false, true>);
        }
        return heavy ? go(k_rows_c2r5<F3R, SW, ADJ, true>) : go(k_rows_c2r5<F3R, SW, ADJ, false>);
    }
    static cudaError_t rows5_fwd(const JbGeom& g, const cd* bufA, const cd* tw1, const cd* tw2, const int* gtab, int gq,
                                 const JbEpi& ep, int planes, int n_sm, jb_stream_t st) {
        if constexpr (PIPE == 2) {
            if (rows5_sw(g) == 1)
                return rows5_launch<1, false>(g, bufA, nullptr, tw1, tw2, gtab, gq, ep, nullptr, nullptr, planes, n_sm, st);
            return rows5_launch<2, false>(g, bufA, nullptr, tw1, tw2, gtab, gq, ep, nullptr, nullptr, planes, n_sm, st);
        } else {
            (void)g; (void)bufA; (void)tw1; (void)tw2; (void)gtab; (void)gq; (void)ep; (void)planes; (void)n_sm; (void)st;
            return (cudaError_t)1;
        }
    }
    static cudaError_t rows5_adj(const JbGeom& g, const cd* bufB, const double* wgt, const cd* tw1, const cd* tw2,
                                 const int* gtab, int gq, double* grad, const double* logsky, int planes_se, int n_sm,
                                 jb_stream_t st) {
        if constexpr (PIPE == 2) {
            JbEpi ep; memset(&ep, 0, sizeof(ep));
            if (rows5_sw(g) == 1)
                return rows5_launch<1, true>(g, bufB, wgt, tw1, tw2, gtab, gq, ep, grad, logsky, planes_se, n_sm, st);
            return rows5_launch<2, true>(g, bufB, wgt, tw1, tw2, gtab, gq, ep, grad, logsky, planes_se, n_sm, st);
        } else {
            (void)g; (void)bufB; (void)wgt; (void)tw1; (void)tw2; (void)gtab; (void)gq; (void)grad; (void)logsky; (void)planes_se; (void)n_sm; (void)st;
            return (cudaError_t)1;
        }
    }
    static int conv_row_threads(int rpc) {
        int items = rpc * (FP::R2 > 1 ? (FP::R2 > FP::H1 ? FP::R2 : FP::H1) : 1);
        int thr = ((items + 31) / 32) * 32;
        return thr < 32 ? 32 : (thr > 256 ? 256 : thr);
    }
    static cudaError_t conv_rows_r2c(const double* in, cd* out, const cd* tw_nk, long nrows, jb_stream_t st) {
        const int rpc = 8;
        return jb_launch(k_conv_rows_r2c<FP>, dim3((unsigned)((nrows + rpc - 1) / rpc)), dim3(conv_row_threads(rpc)),
                         jb_al16((size_t)rpc * FP::LS * sizeof(cd)) + 16, st, in, out, tw_nk, nrows, FP::L / 2 + 1, rpc);
    }
    static cudaError_t conv_rows_c2r(const cd* in, double* out, const cd* tw_kn, long nrows, double scale, jb_stream_t st) {
        const int rpc = 8;
        return jb_launch(k_conv_rows_c2r<FP>, dim3((unsigned)((nrows + rpc - 1) / rpc)), dim3(conv_row_threads(rpc)),
                         jb_al16((size_t)rpc * FP::LS * sizeof(cd)) + 16, st, in, out, tw_kn, nrows, FP::L / 2 + 1, rpc, scale);
    }
    static cudaError_t conv_cols(const cd* xs, cd* ys, cd* scratch, const cd* tw_nk, int PH, int lead, int x_lead,
                                 jb_stream_t st) {
        int ct = ((CB * (FP::R2 > 1 ? (FP::R1 > FP::R2 ? FP::R1 : FP::R2) : 1) + 31) / 32) * 32;
        if (ct > 128) ct = 128;
        size_t smem = (FP::R2 == 1) ? 16 : (size_t)FP::L * CB * sizeof(cd);
        return jb_launch(k_conv_cols<FP, CB>, dim3((unsigned)jb_cdiv(PH, CB), (unsigned)lead), dim3(ct), smem, st, xs, ys, scratch,
                         tw_nk, PH, x_lead);
    }
    static cudaError_t prep_rows(const JbGeom& g, const JbKsrc& ks, const cd* tw_nk, cd* scratch, int rpc,
                                 int threads, jb_stream_t st) {
        size_t smem = jb_al16((size_t)rpc * FP::LS * sizeof(cd)) + 16;
        return jb_launch(k_prep_rows<FP>, dim3(jb_cdiv(g.w, rpc), ks.T * ks.E * g.n * g.n), dim3(threads), smem, st,
                         g, ks, tw_nk, scratch, rpc);
    }
    static cudaError_t prep_cols(const JbGeom& g, const cd* scratch, const double* sums, const cd* twa, const cd* twb,
                                 cd* khat, int layout, int threads, jb_stream_t st) {
        dim3 grid(jb_cdiv(g.PH, CB), g.n * g.n, g.T * g.E);
        if constexpr (PIPE == 2) {
            if (layout == 2) {
                size_t smem3 = (size_t)FP::L * CB * sizeof(cd) + (size_t)(F3::L + F3::M) * sizeof(cd);
                return jb_launch(k_prep_cols3<F3, CB>, grid, dim3(Cols3<F3, CB>::THREADS), smem3, st, g, scratch, sums, twa, twb, khat,
                                 (int)JB_COLS_WP);
            }
        }
        size_t smem = (FP::R2 == 1) ? 16 : (size_t)FP::L * CB * sizeof(cd);
        return jb_launch(k_prep_cols<FP, CB>, grid, dim3(threads), smem, st, g, scratch, sums, twa, khat);
    }
    static const JbFftOps* table() {
        static const JbFftOps t = {FP::L, FP::R1, FP::R2, CB, FP::LS, (FP::R2 > 1 && PIPE == 2) ? 2 : 0, F3::R1, F3::R2, F3::R3,
                                   F3R::R1, F3R::R2, F3R::R3,
                                   rows_fwd, rows_adj_in, cols, cols_pipe, (int)(PIPE == 2 && JB_COLS_MT && !JB_COLS_WP),
                                   (int)Cols3<F3, CB>::CPS,
                                   rows_out, rows_adj_out, rows_generic_smem,
                                   rows5_smem, rows5_gather, rows5_fwd, rows5_adj, conv_rows_r2c, conv_rows_c2r, conv_cols,
                                   prep_rows, prep_cols};
        return &t;
    }
};

// factor order of the c2r row kernels: the column pass's unless fft_list.h (JB_ROWS_F3_LIST) or the variant builder
// (-DJB_INST_RQ1/2/3) names another one
template <int LEN, int Q1, int Q2, int Q3>
struct JbRowsF3 { using type = Fft3<LEN, Q1, Q2, Q3>; };
#ifdef JB_INST_RQ1
template <int Q1, int Q2, int Q3>
struct JbRowsF3<JB_INST_L, Q1, Q2, Q3> { using type = Fft3<JB_INST_L, JB_INST_RQ1, JB_INST_RQ2, JB_INST_RQ3>; };
#else
#define JB_Y_ROWS_F3(LEN, A, B, C) \
    template <int Q1, int Q2, int Q3> struct JbRowsF3<LEN, Q1, Q2, Q3> { using type = Fft3<LEN, A, B, C>; };
JB_ROWS_F3_LIST(JB_Y_ROWS_F3)
#endif
#define JB_DEFINE_FFT(LEN, R1, R2, CB, PIPE, Q1, Q2, Q3) \
    const JbFftOps* jb_fft_ops_##LEN() { \
        return JbOpsImpl<FftPlan<LEN, R1, R2>, CB, PIPE, Fft3<(PIPE == 2 ? LEN : 1), Q1, Q2, Q3>, \
                         typename JbRowsF3<(PIPE == 2 ? LEN : 1), Q1, Q2, Q3>::type>::table(); }
#define JB_DECLARE_FFT(LEN) const JbFftOps* jb_fft_ops_##LEN();
