// rows6.cuh -- the two r2c row kernels (image rows -> row spectra) in the warp-private style of rows5.cuh.
//
// One CTA = one image row per iteration.  The row (sky row, or cotangent row [x exposure]) is read once, coalesced,
// into shared memory; every warp packs the windows of SW pairs of column blocks (2c, 2c+1) as z = x_a + i x_b into
// its own tile (forward: w samples x column weight at offset m, zero padded; adjoint: the P-wide window), runs the
// three FORWARD stages (radix R1, R2, R3 with the twiddles of cols3.cuh) with warp-level syncs only, separates the two
// Hermitian spectra  X_a[k] = (Z[k] + conj Z[L-k]) / 2,  X_b[k] = (Z[k] - conj Z[L-k]) / 2i  and stores the two rows
// with coalesced 128-bit stores.  Same tile layout, item descriptors and launch shape as the c2r kernels, 24 warps
// per SM instead of the 16 of rows_fast.cuh's two-pass kernels; one block barrier per row (double-buffered row).
#pragma once
#include "rows5.cuh"

template <class F3, int SW>
struct Rows6 {
    using R = Rows5<F3, SW>;
    static constexpr int L = F3::L, R1 = F3::R1, R2 = F3::R2, R3 = F3::R3, M = F3::M;
    // forward stage 1: radix R1 over n1, then * W_L^{m k1}; items m = R3*n2 + n3
    static JB_D void s1(cd* tile, const cd* __restrict__ tw1, const int* __restrict__ itab, int lane) {
        static_for<0, R::N1>([&](auto rr) {
            constexpr int r = rr;
            const int d = itab[(R::N3 + R::N2 + r) * 32 + lane];
            if (d >= 0) {
                cd* b = tile + (d & 0xffff);
                const cd* tw = tw1 + (d >> 16);
                cd v[R1];
                static_for<0, R1>([&](auto nn) { constexpr int n1 = nn; v[n1] = b[n1 * R::BP]; });
                dft<R1, -1>(v);
                static_for<0, R1>([&](auto kk) {
                    constexpr int k1 = kk;
                    cd y = v[k1];
                    if constexpr (k1 > 0) y = cd_mul(y, jb_ldg(tw + k1 * M));
                    b[k1 * R::BP] = y;
                });
            }
        });
    }
    // forward stage 2: radix R2 over n2, then * W_M^{n3 k2}; items (k1, n3)
    static JB_D void s2(cd* tile, const cd* __restrict__ tw2, const int* __restrict__ itab, int lane) {
        static_for<0, R::N2>([&](auto rr) {
            constexpr int r = rr;
            const int d = itab[(R::N3 + r) * 32 + lane];
            if (d >= 0) {
                cd* b = tile + (d & 0xffff);
                const cd* tw = tw2 + (d >> 16);
                cd v[R2];
                static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; v[n2] = b[n2 * R::R3P]; });
                dft<R2, -1>(v);
                static_for<0, R2>([&](auto kk) {
                    constexpr int k2 = kk;
                    cd y = v[k2];
                    if constexpr (k2 > 0) y = cd_mul(y, jb_ldg(tw + k2 * R3));
                    b[k2 * R::R3P] = y;
                });
            }
        });
    }
    // forward stage 3: radix R3 over n3 -> Z[k1 + R1*(k2 + R2*k3)] at k1*BP + k2*R3P + k3; items (k1, k2)
    static JB_D void s3(cd* tile, const int* __restrict__ itab, int lane) {
        static_for<0, R::N3>([&](auto rr) {
            constexpr int r = rr;
            const int d = itab[r * 32 + lane];
            if (d >= 0) {
                cd* b = tile + d;
                cd v[R3];
                static_for<0, R3>([&](auto nn) { constexpr int n3 = nn; v[n3] = b[n3]; });
                dft<R3, -1>(v);
                static_for<0, R3>([&](auto kk) { constexpr int k3 = kk; b[k3] = v[k3]; });
            }
        });
    }
};

// dynamic shared memory: NW tiles, the packed spectrum addresses [NE][32], the item descriptors, the sample-address
// table [L] (16 bit) and two copies of the image row
template <class F3, int SW>
inline size_t jb_rows6_smem(int n, int N) {
    using R = Rows5<F3, SW>;
    const int NC = (n + 1) / 2, NW = (NC + SW - 1) / SW;
    return (size_t)NW * SW * R::SEQ * sizeof(cd) + (size_t)(R::NE + R::NR) * 32 * sizeof(unsigned) +
           (((size_t)F3::L * sizeof(unsigned short) + 15) & ~(size_t)15) + 2 * (size_t)N * sizeof(double);
}

// grid: persistent, block = NW*32.
//   ADJ = false: in = x (S,E,N,N); planes over (S,Es); out = rowsF (S*E, n, N, PH)
//   ADJ = true : in = cotangent (S,T,E,N,N) [x expo (T,E,N,N) if given]; planes over (S,T,Es); out = rowsA (STE, n, N, PH)
template <class F3, int SW, bool ADJ>
JB_GLOBAL void JB_LB((SW == 1 ? 128 : 512), (SW == 1 ? JB_R5_CPS : 1))
k_rows_r2c6(JbGeom g, const double* __restrict__ in, const double* __restrict__ expo, const double* __restrict__ wgt,
            const cd* __restrict__ tw1, const cd* __restrict__ tw2, cd* __restrict__ out, int planes) {
    using R = Rows5<F3, SW>;
    using Q = Rows6<F3, SW>;
    constexpr int L = F3::L;
    JB_SMEM(unsigned char, smraw);
    const int n = g.n, N = g.N, m = g.m, hp = g.hp, PH = g.PH, w = g.w, P = g.P;
    const int NC = (n + 1) / 2, NW = (NC + SW - 1) / SW;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NT = NW * 32;
    cd* tiles = reinterpret_cast<cd*>(smraw);
    unsigned* atab = reinterpret_cast<unsigned*>(tiles + (size_t)NW * SW * R::SEQ);
    int* itab = reinterpret_cast<int*>(atab + R::NE * 32);
    unsigned short* ptab = reinterpret_cast<unsigned short*>(itab + R::NR * 32);
    double* xrow = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ptab) +
                                             (((size_t)L * sizeof(unsigned short) + 15) & ~(size_t)15));
    cd* tile = tiles + (size_t)wid * SW * R::SEQ;
    for (int i = tid; i < R::NE * 32; i += NT) {
        int k = i;
        unsigned lo = (k < R::PHL) ? (unsigned)R::addr_of_freq(k) : 0u;
        unsigned hi = (k > 0 && 2 * k < L) ? (unsigned)R::addr_of_freq(L - k) : 0xffffu;
        atab[i] = lo | (hi << 16);
    }
    for (int p = tid; p < L; p += NT) ptab[p] = (unsigned short)R::addr_of_sample(p);
    R::build_items(itab, tid, NT);
    // window geometry: sample p of column block i is image column offs[i] - back + p for p in [plo, phi), else zero
    const int back = ADJ ? hp + m : hp + m;     // forward: column of window sample q = p - m is offs[i] - hp + q
    const int plo = ADJ ? 0 : m, phi = ADJ ? P : m + w;
    int ca[SW][2];
    bool has[SW][2];
    static_for<0, SW>([&](auto ss) {
        constexpr int s = ss;
        const int ia = 2 * (wid * SW + s);
        has[s][0] = ia < n; has[s][1] = ia + 1 < n;
        ca[s][0] = has[s][0] ? g.offs[ia] - back : 0;
        ca[s][1] = has[s][1] ? g.offs[ia + 1] - back : 0;
    });
    const long nitems = (long)planes * N;
    int buf = 0;
    for (long item = blockIdx.x; item < nitems; item += gridDim.x, buf ^= 1) {
        const int lp = (int)(item / N), r = (int)(item - (long)lp * N);
        int ste = 0, te = 0, se = 0, s_i, e_i;
        size_t oplane;
        const double* src;
        const double* ex = nullptr;
        if (!ADJ) {
            se = jb_map_se(g, lp, s_i, e_i);
            src = in + ((size_t)se * N + r) * N;
            oplane = (size_t)se * n;
        } else {
            jb_map_ste(g, lp, ste, te, se);
            src = in + ((size_t)ste * N + r) * N;
            if (expo) ex = expo + ((size_t)te * N + r) * N;
            oplane = (size_t)ste * n;
        }
        double* xr = xrow + (size_t)buf * N;
        for (int c = tid; c < N; c += NT) {
            double v = src[c];
            if (ADJ && ex) v *= jb_ldg(ex + c);
            xr[c] = v;
        }
        JB_SYNC();   // also: every warp has finished reading the other row buffer (used two rows ago)
        // ---- pack the windows of the warp's pairs: z[p] = x_a[p] + i x_b[p]
        static_for<0, SW>([&](auto ss) {
            constexpr int s = ss;
            if (has[s][0]) {
                cd* ts = tile + s * R::SEQ;
                for (int p = lane; p < L; p += 32) {
                    double za = 0.0, zb = 0.0;
                    if (p >= plo && p < phi) {
                        const double wq = ADJ ? 1.0 : jb_ldg(wgt + (p - m));
                        const int c0 = ca[s][0] + p, c1 = ca[s][1] + p;
                        if (c0 >= 0 && c0 < N) za = xr[c0] * wq;
                        if (has[s][1] && c1 >= 0 && c1 < N) zb = xr[c1] * wq;
                    }
                    ts[ptab[p]] = cd_make(za, zb);
                }
            }
        });
        JB_SYNCWARP();
        Q::s1(tile, tw1, itab, lane);
        JB_SYNCWARP();
        Q::s2(tile, tw2, itab, lane);
        JB_SYNCWARP();
        Q::s3(tile, itab, lane);
        JB_SYNCWARP();
        // ---- separate the two Hermitian spectra and store the rows
        static_for<0, SW>([&](auto ss) {
            constexpr int s = ss;
            if (has[s][0]) {
                const cd* ts = tile + s * R::SEQ;
                const int ia = 2 * (wid * SW + s);
                cd* oa = out + ((oplane + ia) * N + r) * PH;
                cd* ob = oa + (size_t)N * PH;
#pragma unroll
                for (int e = 0; e < R::NE; ++e) {
                    const int k = lane + 32 * e;
                    if (k < R::PHL) {
                        const unsigned ad = atab[e * 32 + lane];
                        const int adk = (int)(ad & 0xffffu), adm = (int)(ad >> 16);
                        const cd zk = ts[adk], zm = ts[adm == 0xffff ? adk : adm];
                        oa[k] = cd_make(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
                        if (has[s][1]) ob[k] = cd_make(0.5 * (zk.y + zm.y), 0.5 * (zm.x - zk.x));
                    }
                }
            }
        });
        JB_SYNCWARP();   // the tile is refilled by this warp only after all its lanes have read it
    }
}
