// rows5.cuh -- the two c2r row kernels (spectra -> image rows) with WARP-PRIVATE transforms.
//
// An earlier generation (round 1, rows3.cuh, removed) ran one CTA-wide transform per image row: six block barriers
// around stages of one radix-8 butterfly per thread, a single warp issuing the bulk copies, and an assembly stage that
// was mostly address arithmetic (ncu r1f: 33 M warp instructions of which 7 M FP64, FP64 pipe 13 %, DRAM 27 %).  Here every warp owns SW packed
// complex sequences Z = X_a + i X_b (two Hermitian row spectra each, columns blocks i = 2c, 2c+1 of one image row):
//   * the lanes read the <= JB_MAXQ covering row-patches (x modules in the adjoint) straight from HBM with
//     coalesced 128-bit loads (no staging buffer, no mbarrier), sum them in registers and scatter Z[k], Z[L-k]
//     into the warp's own shared-memory tile;
//   * the three inverse stages (radix R3, R2, R1 of cols3.cuh's factorisation) run on that tile with warp-level
//     syncs only, conflict-free thanks to an odd row pitch;
//   * the real rows stay in place (sample p of rows a/b = re/im of tile element p), then ONE block barrier, the
//     overlap-add gather over the column blocks, the epilogue (PSF / response / metric / fused Poisson energy +
//     cotangent) and the coalesced store.
// A CTA = ceil(ceil(n/2)/SW) warps = one image row per iteration; many small CTAs per SM hide the load latency.
#pragma once
#include "rows_fast.cuh"
#include "cols3.cuh"

#ifndef JB_R5_B
#define JB_R5_B 3     // spectrum elements x row-patches whose loads a lane keeps in flight (per sequence)
#endif
#ifndef JB_R5_GU
#define JB_R5_GU 1    // unroll factor of the gather / epilogue pixel loop
#endif
#ifndef JB_R5_AU
#define JB_R5_AU 1    // unroll factor of the assembly loop over spectrum-element groups
#endif
#ifndef JB_R5_CPS
#define JB_R5_CPS 6   // CTAs per SM the SW = 1 kernel is register-limited to
#endif
#ifndef JB_R5_WIDE_T
#define JB_R5_WIDE_T 256   // launch-bound threads of the WIDE form (5..8 warps)
#endif
#ifndef JB_R5_BW
#define JB_R5_BW JB_R5_B   // load budget (see JB_R5_B) of the WIDE form
#endif
#ifndef JB_R5_APD
#define JB_R5_APD 2     // depth of the forward assembly pipeline: spectrum elements in flight per lane
#endif
#ifndef JB_R5_APIPE
#define JB_R5_APIPE 3   // 1: software-pipelined assembly in the forward WIDE kernel (r2v, C5 chunk: rows_out 6.79 -> 6.21 ms,
                        //    bitwise identical); 2: also in the narrow forward kernel (no gain at its 80 registers: 190.4 ->
                        //    190.9 us at C2); 3 (default): also over the modules of the adjoint WIDE kernel (rows_adj_out 2.05 -> 1.93 ms)
#endif
#ifndef JB_R5_APIPE_ADJ1
#define JB_R5_APIPE_ADJ1 1   // adjoint WIDE kernel on module-summed input (T == 1): element-pipelined assembly
#endif
#ifndef JB_R5_SPF
#define JB_R5_SPF 2   // spectrum prefetch (forward kernel): 1 = before, 2 = after the assembly of a row, one
                      // cp.async.bulk.prefetch.L2 per covering (row-patch, column block) spectrum row of the NEXT image row of
                      // this CTA.  r2b, B200: forward c2r pass 202.7 -> 191.1 us (C2, S = 2), 6.97 -> 6.77 ms (C5 chunk) with
                      // mode 2; mode 1 is slower (229 us: the prefetches queue in front of the row's own loads).  The adjoint
                      // kernel gathers T modules per row -- its prefetch volume overflows L2 (2.06 -> 2.61 ms) -- so it stays off.
#endif
#ifndef JB_R5_EPF
#define JB_R5_EPF 1   // 1: warm L2 with this row's exposure / counts (Poisson epilogue operands) before the spectra are
                      //    assembled (r1z: forward c2r pass 124.1 -> 119.8 us at C2, S = 1; bitwise identical results)
#endif

template <class F3, int SW>
struct Rows5 {
    static constexpr int L = F3::L, R1 = F3::R1, R2 = F3::R2, R3 = F3::R3, M = F3::M;
    static constexpr int R3P = R3 | 1;              // odd pitch of a k2 block: stage-3 items are conflict-free
    static constexpr int BP = (R2 * R3P) | 1;       // odd pitch of a k1 block: the assembly stores are conflict-free
    static constexpr int SEQ = R1 * BP;             // cd per sequence
    static constexpr int PHL = L / 2 + 1;
    static constexpr int NE = (PHL + 31) / 32;      // spectrum elements per lane and sequence
    static constexpr int I3 = R1 * R2, I2 = R1 * R3, I1 = M;   // items per sequence and stage
    static constexpr int N3 = (SW * I3 + 31) / 32, N2 = (SW * I2 + 31) / 32, N1 = (SW * I1 + 31) / 32;  // rounds
    static_assert(SEQ < 65535, "assembly addresses are packed in 16 bits");

    static JB_HD int addr_of_freq(int k) {          // frequency k = k1 + R1*(k2 + R2*k3)
        int k1 = k % R1, q = k / R1, k2 = q % R2, k3 = q / R2;
        return k1 * BP + k2 * R3P + k3;
    }
    static JB_HD int addr_of_sample(int p) {        // sample p = M*n1 + R3*n2 + n3
        int n1 = p / M, mm = p - n1 * M, n2 = mm / R3, n3 = mm - n2 * R3;
        return n1 * BP + n2 * R3P + n3;
    }
    // per-lane item descriptors of the three stages (the same for every image row), kept in shared memory as
    // itab[round][lane]: stage 3 rounds, then stage 2, then stage 1; value = tile offset | (twiddle index << 16),
    // -1 = idle lane
    static constexpr int NR = N3 + N2 + N1;
    static JB_D void build_items(int* itab, int tid, int nthr) {
        for (int idx = tid; idx < NR * 32; idx += nthr) {
            const int r = idx >> 5, lane = idx & 31;
            int v = -1;
            if (r < N3) {
                int it = lane + 32 * r;
                int s = it / I3, c = it - s * I3, k1 = c / R2, k2 = c - k1 * R2;
                if (it < SW * I3) v = s * SEQ + k1 * BP + k2 * R3P;
            } else if (r < N3 + N2) {
                int it = lane + 32 * (r - N3);
                int s = it / I2, c = it - s * I2, k1 = c / R3, n3 = c - k1 * R3;
                if (it < SW * I2) v = (s * SEQ + k1 * BP + n3) | (n3 << 16);
            } else {
                int it = lane + 32 * (r - N3 - N2);
                int s = it / I1, mm = it - s * I1, n2 = mm / R3, n3 = mm - n2 * R3;
                if (it < SW * I1) v = (s * SEQ + n2 * R3P + n3) | (mm << 16);
            }
            itab[idx] = v;
        }
    }
    // inverse stage 3: radix R3 over k3 (contiguous), items (k1, k2)
    static JB_D void s3(cd* tile, const int* __restrict__ itab, int lane) {
        static_for<0, N3>([&](auto rr) {
            constexpr int r = rr;
            const int d = itab[r * 32 + lane];
            if (d >= 0) {
                cd* b = tile + d;
                cd v[R3];
                static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; v[k3] = b[k3]; });
                dft<R3, +1>(v);
                static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; b[n3] = v[n3]; });
            }
        });
    }
    // inverse stage 2: conj twiddle W_M^{n3 k2}, radix R2 over k2, items (k1, n3)
    static JB_D void s2(cd* tile, const cd* __restrict__ tw2, const int* __restrict__ itab, int lane) {
        static_for<0, N2>([&](auto rr) {
            constexpr int r = rr;
            const int d = itab[(N3 + r) * 32 + lane];
            if (d >= 0) {
                cd* b = tile + (d & 0xffff);
                const cd* tw = tw2 + (d >> 16);
                cd v[R2];
                static_for<0, R2>([&](auto kk) {
                    constexpr int k2 = kk;
                    cd y = b[k2 * R3P];
                    if constexpr (k2 > 0) y = cd_mulc(y, jb_ldg(tw + k2 * R3));
                    v[k2] = y;
                });
                dft<R2, +1>(v);
                static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; b[n2 * R3P] = v[n2]; });
            }
        });
    }
    // inverse stage 1: conj twiddle W_L^{m k1}, radix R1 over k1, items m = R3*n2 + n3; the results (sample
    // M*n1 + m: re = row a, im = row b) go back in place.  ADJ: crop to [p0, p0+len) and multiply by wgt[p - p0].
    template <bool ADJ>
    static JB_D void s1(cd* tile, const cd* __restrict__ tw1, const double* __restrict__ wgt, int p0, int len,
                        const int* __restrict__ itab, int lane) {
        static_for<0, N1>([&](auto rr) {
            constexpr int r = rr;
            const int d = itab[(N3 + N2 + r) * 32 + lane];
            if (d >= 0) {
                cd* b = tile + (d & 0xffff);
                const int mm = d >> 16;
                const cd* tw = tw1 + mm;
                cd v[R1];
                static_for<0, R1>([&](auto kk) {
                    constexpr int k1 = kk;
                    cd y = b[k1 * BP];
                    if constexpr (k1 > 0) y = cd_mulc(y, jb_ldg(tw + k1 * M));
                    v[k1] = y;
                });
                dft<R1, +1>(v);
                static_for<0, R1>([&](auto nn) {
                    constexpr int n1 = nn;
                    cd y = v[n1];
                    if (ADJ) {
                        int q = M * n1 + mm - p0;
                        double ww = (q >= 0 && q < len) ? jb_ldg(wgt + q) : 0.0;
                        y = cd_scale(y, ww);
                    }
                    b[n1 * BP] = y;
                });
            }
        });
    }
};

// dynamic shared memory: NW tiles (+ one zero element the gather table points unused slots at), the packed assembly
// addresses [NE][32], reduction scratch
template <class F3, int SW>
inline size_t jb_rows5_smem(int n) {
    const int NC = (n + 1) / 2, NW = (NC + SW - 1) / SW;
    return ((size_t)NW * SW * Rows5<F3, SW>::SEQ + 1) * sizeof(cd) +
           (size_t)(Rows5<F3, SW>::NE + Rows5<F3, SW>::NR) * 32 * sizeof(unsigned) + (size_t)NW * 32 * sizeof(double);
}

// Gather table (plan time, host): for image column c the doubles of the CTA's tiles that overlap-add into it,
//   tab[q*N + c] = 2*((i/2)*SEQ + addr_of_sample(p + p0)) + (i & 1)  for the q-th window i with p = c + shift - offs[i]
//   in [0, len); unused slots point at the zero element behind the tiles.  Returns the slots per column (<= maxq), 0 on
//   overflow.
template <class F3, int SW>
inline int jb_rows5_gather_table(const JbGeom& g, bool adj, int maxq, int* tab) {
    using R = Rows5<F3, SW>;
    const int NC = (g.n + 1) / 2, NW = (NC + SW - 1) / SW;
    const int zero = 2 * (NW * SW * R::SEQ);
    const int shift = adj ? g.hp : g.hp + g.m, len = adj ? g.w : g.P, p0 = adj ? g.m : 0;
    int used = 0;
    for (int q = 0; q < maxq; ++q) for (int c = 0; c < g.N; ++c) tab[(size_t)q * g.N + c] = zero;
    for (int c = 0; c < g.N; ++c) {
        int cnt = 0;
        for (int i = 0; i < g.n; ++i) {
            int p = c + shift - g.offs[i];
            if (p < 0 || p >= len) continue;
            if (cnt >= maxq) return 0;
            tab[(size_t)cnt * g.N + c] = 2 * ((i >> 1) * R::SEQ + R::addr_of_sample(p + p0)) + (i & 1);
            ++cnt;
        }
        used = cnt > used ? cnt : used;
    }
    return used;
}

// one image row: summation over the covering row-patches (x modules) and assembly of the warp's SW sequences
//   pbase[s][0/1] : row 0 of column blocks ia / ib of the first source plane; tstride: elements between modules (ADJ)
// CNT (row-patches covering the row) is a compile-time constant and the loads of EG consecutive spectrum elements
// (EG x SW sequences x 2 rows x CNT row-patches) are all issued before the first use, so a warp pays one DRAM
// round trip per EG elements (the loads are volatile asm: the compiler keeps this order).
template <class F3, int SW, bool ADJ, int CNT, int EG>
JB_D void jb_rows5_assemble_c(cd* tile, const cd* const (&pbase)[SW][2], const bool (&has)[SW][2], size_t tstride, int T,
                              const int (&rowoff)[JB_MAXQ], const double (&wr)[JB_MAXQ],
                              const unsigned* __restrict__ atab, int lane) {
    using R = Rows5<F3, SW>;
    constexpr int NG = (R::NE + EG - 1) / EG;
    constexpr int AU = JB_R5_AU;
#pragma unroll AU
    for (int gi = 0; gi < NG; ++gi) {
        const int e0 = gi * EG;
        cd xa[EG][SW], xb[EG][SW];
        static_for<0, EG>([&](auto ee) {
            static_for<0, SW>([&](auto ss) { xa[ee][ss] = cd_make(0.0, 0.0); xb[ee][ss] = cd_make(0.0, 0.0); });
        });
        size_t toff = lane + 32 * e0;
        for (int t = 0; t < T; ++t, toff += tstride) {
            cd va[EG][SW][CNT], vb[EG][SW][CNT];
            static_for<0, EG>([&](auto ee) {
                constexpr int e = ee;
                if (lane + 32 * (e0 + e) < R::PHL) {
                    static_for<0, SW>([&](auto ss) {
                        constexpr int s = ss;
                        static_for<0, CNT>([&](auto qq) {
                            constexpr int q = qq;
                            va[e][s][q] = jb_ld_stream(pbase[s][0] + toff + 32 * e + rowoff[q]);
                            vb[e][s][q] = jb_ld_stream(pbase[s][1] + toff + 32 * e + rowoff[q]);
                        });
                    });
                }
            });
            static_for<0, EG>([&](auto ee) {
                constexpr int e = ee;
                if (lane + 32 * (e0 + e) < R::PHL) {
                    static_for<0, SW>([&](auto ss) {
                        constexpr int s = ss;
                        static_for<0, CNT>([&](auto qq) {
                            constexpr int q = qq;
                            if (ADJ) {
                                xa[e][s].x += wr[q] * va[e][s][q].x; xa[e][s].y += wr[q] * va[e][s][q].y;
                                xb[e][s].x += wr[q] * vb[e][s][q].x; xb[e][s].y += wr[q] * vb[e][s][q].y;
                            } else {
                                xa[e][s] = cd_add(xa[e][s], va[e][s][q]); xb[e][s] = cd_add(xb[e][s], vb[e][s][q]);
                            }
                        });
                    });
                }
            });
        }
        static_for<0, EG>([&](auto ee) {
            constexpr int e = ee;
            const int k = lane + 32 * (e0 + e);
            if (k < R::PHL) {
                const unsigned ad = atab[(e0 + e) * 32 + lane];
                const int adk = (int)(ad & 0xffffu), adm = (int)(ad >> 16);
                static_for<0, SW>([&](auto ss) {
                    constexpr int s = ss;
                    if (has[s][0]) {
                        cd a = xa[e][s], b = has[s][1] ? xb[e][s] : cd_make(0.0, 0.0);
                        cd* ts = tile + s * R::SEQ;
                        ts[adk] = cd_make(a.x - b.y, a.y + b.x);
                        if (adm != 0xffff) ts[adm] = cd_make(a.x + b.y, b.x - a.y);
                    }
                });
            }
        });
    }
}

// Forward, one sequence per warp: the same assembly with the loads of element e+1 in flight while element e is summed
// and scattered (two register buffers, fully unrolled) -- the warp pays the memory round trip once per row instead of
// once per element group.  Costs 8*CNT registers: used where the launch bound leaves them (JB_R5_APIPE).
template <class F3, int CNT, bool WGT = false>
JB_D void jb_rows5_assemble_pipe(cd* tile, const cd* pa, const cd* pb, bool hasb, const int (&rowoff)[JB_MAXQ],
                                 const unsigned* __restrict__ atab, int lane, const double* wr = nullptr) {
    using R = Rows5<F3, 1>;
    constexpr int D = JB_R5_APD;          // elements in flight (register buffers)
    cd va[D][CNT], vb[D][CNT];
    auto load = [&](auto ee) {
        constexpr int e = ee, buf = e % D;
        if (lane + 32 * e < R::PHL) {
            static_for<0, CNT>([&](auto qq) {
                constexpr int q = qq;
                va[buf][q] = jb_ld_stream(pa + lane + 32 * e + rowoff[q]);
                vb[buf][q] = jb_ld_stream(pb + lane + 32 * e + rowoff[q]);
            });
        }
    };
    auto consume = [&](auto ee) {
        constexpr int e = ee, buf = e % D;
        const int k = lane + 32 * e;
        if (k < R::PHL) {
            cd a = va[buf][0], b = vb[buf][0];
            if (WGT) { a = cd_scale(a, wr[0]); b = cd_scale(b, wr[0]); }     // adjoint: window row weights (same order of
            static_for<1, CNT>([&](auto qq) {                                  // operations as jb_rows5_assemble_pipe_adj)
                constexpr int q = qq;
                if (WGT) {
                    a.x += wr[q] * va[buf][q].x; a.y += wr[q] * va[buf][q].y;
                    b.x += wr[q] * vb[buf][q].x; b.y += wr[q] * vb[buf][q].y;
                } else {
                    a = cd_add(a, va[buf][q]); b = cd_add(b, vb[buf][q]);
                }
            });
            if (!hasb) b = cd_make(0.0, 0.0);
            const unsigned ad = atab[e * 32 + lane];
            const int adk = (int)(ad & 0xffffu), adm = (int)(ad >> 16);
            tile[adk] = cd_make(a.x - b.y, a.y + b.x);
            if (adm != 0xffff) tile[adm] = cd_make(a.x + b.y, b.x - a.y);
        }
    };
    static_for<0, (D - 1 < R::NE ? D - 1 : R::NE)>([&](auto ee) { load(ee); });
    static_for<0, R::NE>([&](auto ee) {
        constexpr int e = ee;
        if constexpr (e + D - 1 < R::NE) load(std::integral_constant<int, e + D - 1>{});
        consume(ee);
    });
}

// Adjoint, one sequence per warp: the T modules of an element are summed with the loads of module t+1 in flight while
// module t is accumulated (two register buffers, the module loop unrolled by two).
template <class F3, int CNT>
JB_D void jb_rows5_assemble_pipe_adj(cd* tile, const cd* pa, const cd* pb, bool hasb, size_t tstride, int T,
                                     const int (&rowoff)[JB_MAXQ], const double (&wr)[JB_MAXQ],
                                     const unsigned* __restrict__ atab, int lane) {
    using R = Rows5<F3, 1>;
    for (int e = 0; e < R::NE; ++e) {
        const int k = lane + 32 * e;
        if (k >= R::PHL) break;
        cd va[2][CNT], vb[2][CNT];
        cd a = cd_make(0.0, 0.0), b = cd_make(0.0, 0.0);
        auto load = [&](int t, auto bb) {
            constexpr int buf = bb;
            const size_t off = (size_t)t * tstride + k;
            static_for<0, CNT>([&](auto qq) {
                constexpr int q = qq;
                va[buf][q] = jb_ld_stream(pa + off + rowoff[q]);
                vb[buf][q] = jb_ld_stream(pb + off + rowoff[q]);
            });
        };
        auto consume = [&](auto bb) {
            constexpr int buf = bb;
            static_for<0, CNT>([&](auto qq) {
                constexpr int q = qq;
                a.x += wr[q] * va[buf][q].x; a.y += wr[q] * va[buf][q].y;
                b.x += wr[q] * vb[buf][q].x; b.y += wr[q] * vb[buf][q].y;
            });
        };
        load(0, std::integral_constant<int, 0>{});
        for (int t = 0; t < T; t += 2) {
            if (t + 1 < T) load(t + 1, std::integral_constant<int, 1>{});
            consume(std::integral_constant<int, 0>{});
            if (t + 2 < T) load(t + 2, std::integral_constant<int, 0>{});
            if (t + 1 < T) consume(std::integral_constant<int, 1>{});
        }
        if (!hasb) b = cd_make(0.0, 0.0);
        const unsigned ad = atab[e * 32 + lane];
        const int adk = (int)(ad & 0xffffu), adm = (int)(ad >> 16);
        tile[adk] = cd_make(a.x - b.y, a.y + b.x);
        if (adm != 0xffff) tile[adm] = cd_make(a.x + b.y, b.x - a.y);
    }
}

// loads in flight per lane and round trip: EG * SW * 2 * CNT, kept at <= 16 (64 registers)
template <class F3, int SW, bool ADJ, int BUDGET = JB_R5_B>
JB_D void jb_rows5_assemble(cd* tile, const cd* const (&pbase)[SW][2], const bool (&has)[SW][2], size_t tstride, int T,
                            int cnt, const int (&rowoff)[JB_MAXQ], const double (&wr)[JB_MAXQ],
                            const unsigned* __restrict__ atab, int lane) {
    constexpr int B = BUDGET / SW;   // EG * CNT budget
    if (cnt == 1) jb_rows5_assemble_c<F3, SW, ADJ, 1, (B >= 4 ? 4 : B)>(tile, pbase, has, tstride, T, rowoff, wr, atab, lane);
    else if (cnt == 2) jb_rows5_assemble_c<F3, SW, ADJ, 2, (B / 2 >= 1 ? B / 2 : 1)>(tile, pbase, has, tstride, T, rowoff, wr, atab, lane);
    else if (cnt == 3) jb_rows5_assemble_c<F3, SW, ADJ, 3, (B / 3 >= 1 ? B / 3 : 1)>(tile, pbase, has, tstride, T, rowoff, wr, atab, lane);
    else if (cnt == 4) jb_rows5_assemble_c<F3, SW, ADJ, 4, (B / 4 >= 1 ? B / 4 : 1)>(tile, pbase, has, tstride, T, rowoff, wr, atab, lane);
}

#ifndef JB_EMU
// L2 prefetch of every spectrum row that image-row item `nxt` will gather: (covering row-patch j) x (column block i)
// [x (module t) in the adjoint], PH*16 bytes each, spread over the CTA's threads
template <bool ADJ>
JB_D void jb_rows5_prefetch_next(const JbGeom& g, const cd* __restrict__ src, long nxt, long nitems, int ROWS, int shift,
                                 size_t plane_sz, size_t istride, int tid, int NT) {
    if (nxt >= nitems) return;
    const int N = g.N, n = g.n, PH = g.PH;
    const int lp = (int)(nxt / N), r = (int)(nxt - (long)lp * N);
    const cd* plane0;
    int T = 1;
    size_t tstride = 0;
    if (!ADJ) {
        int ste, te, se;
        jb_map_ste(g, lp, ste, te, se);
        plane0 = src + (size_t)ste * plane_sz;
    } else {
        int s_i, e_i;
        jb_map_se(g, lp, s_i, e_i);
        T = g.T;
        plane0 = src + ((size_t)(s_i * g.T) * g.E + e_i) * plane_sz;
        tstride = (size_t)g.E * plane_sz;
    }
    const int per_j = n * T;
    int base = 0;
    for (int j = 0; j < n; ++j) {
        const int u = r + shift - g.offs[j];
        if (u < 0 || u >= ROWS) continue;
        for (int idx = tid - base; idx < per_j; idx += NT) {
            if (idx < 0) continue;
            const int t = idx / n, i = idx - t * n;
            jb_bulk_prefetch_l2(plane0 + (size_t)t * tstride + (size_t)i * istride + (size_t)(j * ROWS + u) * PH,
                                (uint32_t)(PH * sizeof(cd)));
        }
        base = (base + per_j) % NT;
    }
}
#endif

// ------------------------------------------------------------------------------------------------------
// grid: persistent (any size), block = NW*32 threads, NW = ceil(ceil(n/2)/SW).
//   forward (ADJ = false): src = bufA (STE, n*n, P, PH); items = planes*N over launched (S,T,Es) planes; epilogue ep.
//   adjoint (ADJ = true) : src = bufB ((S,T,E), n*n, w, PH); items = planes*N over launched (S,Es) planes; the
//                          spectra of all modules are summed before the transform; out = grad (S,E,N,N).
//   gtab: gather table [GQ][N] (jb_rows5_gather_table).
// SW = 1: one sequence per warp, CTAs of <= 4 warps (n <= 8), 80 registers so that 6 CTAs = 24 warps fit an SM;
// SW = 2: two sequences per warp for wider patch grids (up to 16 warps per CTA).
// POIS: the transcendental epilogue is compiled in (it raises the register pressure, so the plain variants stay lean):
//   forward -- the fused Poisson energy / cotangent (log, divide); adjoint -- the VJP of the sky-model tail, * exp(log_sky).
// WIDE (SW = 1 only): CTAs of 5..8 warps (patch grids of 9..16 per axis, e.g. the 10x10 of the eROSITA-like configs):
// still one sequence per warp and 80 registers, so that e.g. 3 CTAs x 5 warps share an SM at L = 672 where the
// SW = 2 form (3 warps per CTA, 110-128 registers) keeps 9 warps resident.
template <class F3, int SW, bool ADJ, bool POIS, bool WIDE = false>
JB_GLOBAL void JB_LB((SW == 1 ? (WIDE ? JB_R5_WIDE_T : 128) : 512), (SW == 1 ? (WIDE ? 3 : JB_R5_CPS) : 1))
k_rows_c2r5(JbGeom g, const cd* __restrict__ src, const cd* __restrict__ tw1, const cd* __restrict__ tw2,
            const double* __restrict__ wgt, const int* __restrict__ gtab, int GQ, JbEpi ep, double* __restrict__ grad,
            const double* __restrict__ logsky, int planes) {
    using R = Rows5<F3, SW>;
    JB_SMEM(unsigned char, smraw);
    const int n = g.n, N = g.N, m = g.m, hp = g.hp, PH = g.PH;
    const int NC = (n + 1) / 2, NW = (NC + SW - 1) / SW;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NT = NW * 32;
    cd* tiles = reinterpret_cast<cd*>(smraw);
    unsigned* atab = reinterpret_cast<unsigned*>(tiles + (size_t)NW * SW * R::SEQ + 1);
    int* itab = reinterpret_cast<int*>(atab + R::NE * 32);
    double* red = reinterpret_cast<double*>(itab + R::NR * 32);
    cd* tile = tiles + (size_t)wid * SW * R::SEQ;
    const int ROWS = ADJ ? g.w : g.P;            // rows per patch (window length)
    const int shift = ADJ ? hp : hp + m;         // u = r + shift - offs[j]
    const int p0 = ADJ ? m : 0;                  // valid window samples [p0, p0 + ROWS)
    const size_t nn2 = (size_t)n * n;
    // source layout: per-patch rows (planes, n*n, ROWS, PH)
    const size_t plane_sz = nn2 * ROWS * PH;
    const size_t istride = (size_t)n * ROWS * PH;
    for (int i = tid; i < R::NE * 32; i += NT) {
        int k = i;   // e*32 + lane == frequency
        unsigned lo = (k < R::PHL) ? (unsigned)R::addr_of_freq(k) : 0u;
        unsigned hi = (k > 0 && 2 * k < R::L) ? (unsigned)R::addr_of_freq(R::L - k) : 0xffffu;
        atab[i] = lo | (hi << 16);
    }
    if (tid == 0) tiles[(size_t)NW * SW * R::SEQ] = cd_make(0.0, 0.0);
    R::build_items(itab, tid, NT);
    bool has[SW][2];
    static_for<0, SW>([&](auto ss) {
        constexpr int s = ss;
        const int ia = 2 * (wid * SW + s);
        has[s][0] = ia < n; has[s][1] = ia + 1 < n;
    });
    JB_SYNC();
    const long nitems = (long)planes * N;
    for (long it_ = blockIdx.x; it_ < nitems; it_ += gridDim.x) {
        const long item = it_;
        const int lp = (int)(item / N), r = (int)(item - (long)lp * N);
        // ---- sources of this row (uniform over the CTA)
        int rowoff[JB_MAXQ];
        double wr[JB_MAXQ];
        int cnt = 0;
        static_for<0, JB_MAXQ>([&](auto qq) { constexpr int q = qq; rowoff[q] = 0; wr[q] = 0.0; });
        for (int j = 0; j < n; ++j) {
            int u = r + shift - g.offs[j];
            if (u >= 0 && u < ROWS && cnt < JB_MAXQ) {
                static_for<0, JB_MAXQ>([&](auto qq) {
                    constexpr int q = qq;
                    if (cnt == q) { rowoff[q] = (j * ROWS + u) * PH; if (ADJ) wr[q] = jb_ldg(wgt + u); }
                });
                ++cnt;
            }
        }
        int ste = 0, te = 0, se = 0, T = 1, sky_s = 0, sky_e = 0;
        const cd* plane0;
        size_t tstride = 0;
        if (!ADJ) {
            jb_map_ste(g, lp, ste, te, se);
            plane0 = src + (size_t)ste * plane_sz;
        } else {
            int s_i, e_i;
            se = jb_map_se(g, lp, s_i, e_i);
            sky_s = s_i; sky_e = e_i;
            T = g.T;
            plane0 = src + ((size_t)(s_i * g.T) * g.E + e_i) * plane_sz;
            tstride = (size_t)g.E * plane_sz;
        }
        const cd* pbase[SW][2];
        static_for<0, SW>([&](auto ss) {
            constexpr int s = ss;
            const int ia = 2 * (wid * SW + s);
            pbase[s][0] = plane0 + (size_t)(has[s][0] ? ia : 0) * istride;
            pbase[s][1] = plane0 + (size_t)(has[s][1] ? ia + 1 : 0) * istride;
        });
#if !defined(JB_EMU) && JB_R5_EPF
        // the epilogue's operands (exposure, counts / lambda of this image row: N*8 bytes each) are independent of the
        // transform: one bulk L2 prefetch each now, so that the pixel loop behind the block barrier finds them in L2
        if (!ADJ && POIS && tid == 0) {
            jb_bulk_prefetch_l2(ep.expo + ((size_t)te * N + r) * N, (uint32_t)(N * sizeof(double)));
            jb_bulk_prefetch_l2(ep.data + ((size_t)te * N + r) * N, (uint32_t)(N * sizeof(double)));
        }
#endif
#if !defined(JB_EMU) && JB_R5_SPF == 1
        if (!ADJ) jb_rows5_prefetch_next<ADJ>(g, src, it_ + gridDim.x, nitems, ROWS, shift, plane_sz, istride, tid, NT);
#endif
#if JB_R5_APIPE >= 3
        if constexpr (ADJ && SW == 1 && WIDE) {
            // T == 1 (the module-summing column pass has already collapsed the modules): nothing to pipeline over the
            // modules, so the ELEMENTS are pipelined as in the forward kernel (JB_R5_APIPE & 4)
            if ((JB_R5_APIPE_ADJ1) && T == 1) {
                if (cnt == 1) jb_rows5_assemble_pipe<F3, 1, true>(tile, pbase[0][0], pbase[0][1], has[0][1], rowoff, atab, lane, wr);
                else if (cnt == 2) jb_rows5_assemble_pipe<F3, 2, true>(tile, pbase[0][0], pbase[0][1], has[0][1], rowoff, atab, lane, wr);
                else if (cnt == 3) jb_rows5_assemble_pipe<F3, 3, true>(tile, pbase[0][0], pbase[0][1], has[0][1], rowoff, atab, lane, wr);
                else if (cnt == 4) jb_rows5_assemble_pipe<F3, 4, true>(tile, pbase[0][0], pbase[0][1], has[0][1], rowoff, atab, lane, wr);
            } else
            if (cnt == 1) jb_rows5_assemble_pipe_adj<F3, 1>(tile, pbase[0][0], pbase[0][1], has[0][1], tstride, T, rowoff, wr, atab, lane);
            else if (cnt == 2) jb_rows5_assemble_pipe_adj<F3, 2>(tile, pbase[0][0], pbase[0][1], has[0][1], tstride, T, rowoff, wr, atab, lane);
            else if (cnt == 3) jb_rows5_assemble_pipe_adj<F3, 3>(tile, pbase[0][0], pbase[0][1], has[0][1], tstride, T, rowoff, wr, atab, lane);
            else if (cnt == 4) jb_rows5_assemble_pipe_adj<F3, 4>(tile, pbase[0][0], pbase[0][1], has[0][1], tstride, T, rowoff, wr, atab, lane);
        } else
#endif
#if JB_R5_APIPE
        if constexpr (!ADJ && SW == 1 && (WIDE || JB_R5_APIPE == 2 || JB_R5_APIPE == 4)) {
            if (cnt == 1) jb_rows5_assemble_pipe<F3, 1>(tile, pbase[0][0], pbase[0][1], has[0][1], rowoff, atab, lane);
            else if (cnt == 2) jb_rows5_assemble_pipe<F3, 2>(tile, pbase[0][0], pbase[0][1], has[0][1], rowoff, atab, lane);
            else if (cnt == 3) jb_rows5_assemble_pipe<F3, 3>(tile, pbase[0][0], pbase[0][1], has[0][1], rowoff, atab, lane);
            else if (cnt == 4) jb_rows5_assemble_pipe<F3, 4>(tile, pbase[0][0], pbase[0][1], has[0][1], rowoff, atab, lane);
        } else
#endif
        jb_rows5_assemble<F3, SW, ADJ, (WIDE ? JB_R5_BW : JB_R5_B)>(tile, pbase, has, tstride, T, cnt, rowoff, wr, atab, lane);
#if !defined(JB_EMU) && JB_R5_SPF == 2
        if (!ADJ) jb_rows5_prefetch_next<ADJ>(g, src, it_ + gridDim.x, nitems, ROWS, shift, plane_sz, istride, tid, NT);
#endif
        JB_SYNCWARP();
        R::s3(tile, itab, lane);
        JB_SYNCWARP();
        R::s2(tile, tw2, itab, lane);
        JB_SYNCWARP();
        R::template s1<ADJ>(tile, tw1, wgt, p0, ROWS, itab, lane);
        JB_SYNC();
        // ---- overlap-add gather over the column blocks, epilogue, store
        const double* tr = reinterpret_cast<const double*>(tiles);
        double esum = 0.0;
        const size_t orow = ADJ ? jb_sky_row(g, sky_s, sky_e, r) : ((size_t)ste * N + r) * N;
        const size_t erow = ((size_t)te * N + r) * N;
        constexpr int GU = JB_R5_GU;
#pragma unroll GU
        for (int c = tid; c < N; c += NT) {
            double conv = 0.0;
            static_for<0, JB_MAXQ>([&](auto qq) {
                constexpr int q = qq;
                if (q < GQ) conv += tr[jb_ldg(gtab + q * N + c)];
            });
            if (ADJ) {
                if (POIS) conv *= exp(jb_ldg(logsky + orow + c));   // VJP of x = exp(log_sky) (sky_models.py:340-341)
                grad[orow + c] = conv;
            } else if (POIS) {  // JB_EPI_POISSON: energy partial + cotangent expo * (1 - d / lambda)
                const double ex = jb_ldg(ep.expo + erow + c);
                double o = 0.0;
                if (ex != 0.0) {
                    const double lam = ex * conv, d = jb_ldg(ep.data + erow + c);
                    esum += lam - d * log(lam);
                    o = ex * (1.0 - d / lam);
                }
                ep.out[orow + c] = o;
            } else {
                ep.out[orow + c] = jb_epi_pixel(ep, conv, orow + c, erow + c);
            }
        }
        if (!ADJ && POIS) {
            double tot = jb_block_sum(esum, red);   // fixed-order tree; ends with a block barrier
            if (tid == 0) ep.partial[(size_t)ste * N + r] = tot;
        } else {
            JB_SYNC();   // the gather has read the tiles before the next row is assembled into them
        }
    }
}
