// kernels.cuh -- the hot path of J-UBIK's instrument response as sm_100a kernels.
//
// Reference chain (all /root/reference): jubik/convolve.py:66-145 linpatch_convolve =
//   slice_patches(margin 0) :87  ->  * bilinear weights :88-90  ->  zero-pad m :92-101  ->
//   circular FFT convolution with the prepared PSF :104-131,148-176  ->  transpose of
//   slice_patches(margin m) :135-143 (overlap-add)  ->  crop m :144,
// then jubik/response.py:44 (exposure), :110 (readout mask) and the Poisson energy
// sum(lambda) - d log(lambda) (NIFTy jft.Poissonian; formula demos/x-ray-imaging.py:313).
//
// Decomposition used here (separable 2-D FFT with the intermediate spectra in L2/HBM):
//   forward   R : k_rows_fwd (r2c of windowed sky rows) -> k_cols (column FFT * K^ * column IFFT)
//                 -> k_rows_out (c2r + overlap-add as a GATHER + exposure/mask/Poisson epilogue)
//   adjoint R^T : k_rows_adj_in (r2c of cotangent rows) -> k_cols (conj K^) -> k_rows_adj_out
//                 (c2r + crop + weights + overlap-add gather, summed over modules)
// Linearity is exploited twice: (1) the sky-row transform of window column-block i is shared by
// the (<=2) row-patches j that contain the row -- one FFT, stored ONCE per (row, i); the column pass
// reads overlapping row windows of that buffer and applies the row weight itself;
// (2) the spectra of all row-patches j (and modules t, in the adjoint) that land on one output
// row are summed BEFORE the inverse row transform -- one c2r per (row, i) instead of one per
// patch row.  Overlap-add is a deterministic gather; no atomics, no memset.
//
// Patch k = i*n + j sits at row offset offs[j], column offset offs[i] (meshgrid "xy",
// convolve.py:59-63).
#pragma once
#include "fft_pass.cuh"

#define JB_MAXN 64

struct JbGeom {
    int N, n, m, dx, w, hp, P;  // dx=int(N/n), w=2dx, hp=dx//2, P=w+2m  (convolve.py:43-51,104-105)
    int L, PH;                  // FFT length (>= P) and stored spectrum columns L/2+1
    int S, T, E;
    int e0, Es;                 // energy-bin chunk processed by a launch: bins [e0, e0+Es) (whole range by default);
                                // launches index planes over (S, T, Es) / (S, Es) and map to the true plane ids
    int fold, ext;              // period-P alias fold (L > P and PSF half-support s > m): ext = s - m rows/columns at either
                                // end of the P-window alias onto [P, P+ext) / [L-ext, L) of the length-L transform
    int offs[JB_MAXN];          // clamped window starts in the hp-padded frame (lax.dynamic_slice)
    // sky-space strides in doubles (sample, energy plane, row).  0 = contiguous (S,E,N,N).  The sky-model tail
    // exp(log_sky[:E, :N, :N]) (sky_models.py:340-341) reads the PADDED log-sky in place with these, and its VJP writes
    // the padded gradient: the crop costs nothing.
    long long sky_smp, sky_pln;
    int sky_row;
};
// element offset of sky-space row r of (sample s, energy e)
JB_HD size_t jb_sky_row(const JbGeom& g, int s, int e, int r) {
    if (g.sky_row == 0) return (((size_t)s * g.E + e) * g.N + r) * (size_t)g.N;
    return (size_t)s * (size_t)g.sky_smp + (size_t)e * (size_t)g.sky_pln + (size_t)r * (size_t)g.sky_row;
}

// Period-P alias fold (hard part 2; reference semantics jubik/convolve.py:92-131,165-176: the convolution is
// CIRCULAR with period P, so PSF tails beyond the margin wrap around).  With a transform length L >= w + 2s the length-L
// result is the LINEAR convolution (no wrap), laid out circularly: positions [-ext, 0) sit at [L-ext, L) and positions
// [P, P+ext) follow the window.  The period-P result is out[t] = lin[t] + lin[t+P] (t < ext) + lin[t-P] (t >= P-ext);
// the transpose feeds the periodic extension of the P-window.  jb_fold_src maps a length-L position p to the window
// position it aliases (or -1 if p lies in the unused gap): p in [0,P) -> p, [P,P+ext) -> p-P, [L-ext,L) -> p-L+P.
JB_HD int jb_fold_src(const JbGeom& g, int p) {
    if (p < g.P) return p;
    if (!g.fold) return -1;
    if (p < g.P + g.ext) return p - g.P;
    if (p >= g.L - g.ext) return p - g.L + g.P;
    return -1;
}
// second length-L position that folds onto window position t (-1: none)
JB_HD int jb_fold_alias(const JbGeom& g, int t) {
    if (!g.fold) return -1;
    if (t < g.ext) return t + g.P;
    if (t >= g.P - g.ext) return t - g.P + g.L;
    return -1;
}

// launched plane index over (S, T, Es) -> ste = (s*T+t)*E+e, te = t*E+e, se = s*E+e
JB_D void jb_map_ste(const JbGeom& g, int lp, int& ste, int& te, int& se) {
    const int TEs = g.T * g.Es;
    const int s = lp / TEs, r = lp - s * TEs;
    const int t = r / g.Es, e = g.e0 + (r - t * g.Es);
    ste = (s * g.T + t) * g.E + e;
    te = t * g.E + e;
    se = s * g.E + e;
}
// launched plane index over (S, Es) -> se (and s, e)
JB_D int jb_map_se(const JbGeom& g, int lp, int& s, int& e) {
    s = lp / g.Es;
    e = g.e0 + (lp - s * g.Es);
    return s * g.E + e;
}

// Epilogues of the forward chain (conv = linpatch_convolve(x)[s,t,e,r,c], ex = effective exposure, 0 == masked):
//   PSF      out = conv                                      (psf_func)
//   RESP     out = ex * conv                                 (R, dense)
//   POISSON  out = ex * (1 - d / lambda), lambda = ex * conv; energy partial += lambda - d log lambda
//   MFAC     out = ex / conv = ex^2 / lambda                 (linearisation point of the metric, jubik_poisson_set_point)
//   METRIC   out = fac * conv, fac = ex^2 / lambda(x0)       (R^T-ready  ex * (R v) / lambda)
enum { JB_EPI_PSF = 0, JB_EPI_RESP = 1, JB_EPI_POISSON = 2, JB_EPI_METRIC = 3, JB_EPI_MFAC = 4 };

struct JbEpi {
    int mode;
    const double* expo;   // (T,E,N,N) thresholded exposure * eff_area * pixel_area ; 0 == masked
    const double* data;   // (T,E,N,N) counts as fp64
    const double* fac;    // (S,T,E,N,N) for JB_EPI_METRIC: ex^2 / lambda at the linearisation point, 0 == masked
    double* out;          // (S,T,E,N,N)
    double* partial;      // per-CTA energy partials, [STE][gridDim.x]
};

// one pixel of the non-Poisson epilogues
JB_D double jb_epi_pixel(const JbEpi& ep, double conv, size_t o, size_t oe) {
    if (ep.mode == JB_EPI_PSF) return conv;
    if (ep.mode == JB_EPI_METRIC) return jb_ldg(ep.fac + o) * conv;
    const double ex = jb_ldg(ep.expo + oe);
    if (ex == 0.0) return 0.0;
    return ep.mode == JB_EPI_RESP ? ex * conv : ex / conv;
}

// Row-patch list of an image row: all j with u = r + shift - offs[j] in [0, len).
// Stored in shared memory as (j, row) pairs, 2n slots per image row; `row` is the row inside the patch's block of the
// spectrum buffer.  fold (forward direction only): the aliased row of the same patch is a second source.
JB_D void jb_build_jlist(const JbGeom& g, int r0, int rpc, int shift, int len, bool fold, int* jcnt, int* jj, int* ju) {
    const int cap = 2 * g.n;
    for (int rr = jb_tid(); rr < rpc; rr += jb_nthr()) {
        int r = r0 + rr, c = 0;
        if (r < g.N) {
            for (int j = 0; j < g.n; ++j) {
                int u = r + shift - g.offs[j];
                if (u >= 0 && u < len) {
                    jj[rr * cap + c] = j; ju[rr * cap + c] = u; ++c;
                    int a = fold ? jb_fold_alias(g, u) : -1;
                    if (a >= 0) { jj[rr * cap + c] = j; ju[rr * cap + c] = a; ++c; }
                }
            }
        }
        jcnt[rr] = c;
    }
}

JB_D size_t jb_align16(size_t x) { return (x + 15) & ~(size_t)15; }

// ------------------------------------------------------------------ forward: sky rows -> spectra
// grid (ceil(N/rpc), S*E).  x:(S,E,N,N).  rowsF:(S*E, n, N, PH): r2c of the column-weighted, zero-padded window
// row of column block i -- stored once per (row, i); the row weight is applied by the column pass.
template <class FP>
JB_GLOBAL void k_rows_fwd(JbGeom g, const double* __restrict__ x, const double* __restrict__ wgt,
                          const cd* __restrict__ tw_nk, cd* __restrict__ rowsF, int rpc, int exp_in) {
    JB_SMEM(unsigned char, smraw);
    const int B = rpc * g.n;
    cd* sm = reinterpret_cast<cd*>(smraw);
    int s_, e_;
    const int r0 = blockIdx.x * rpc, se = jb_map_se(g, blockIdx.y, s_, e_);
    const int n = g.n, N = g.N, m = g.m, w = g.w, hp = g.hp, PH = g.PH;
    auto ld = [&](int b, int p) -> double {
        int rr = b / n, i = b - rr * n, r = r0 + rr;
        int q = p - m;
        int c = g.offs[i] - hp + q;
        if (r >= N || q < 0 || q >= w || c < 0 || c >= N) return 0.0;
        double v = jb_ldg(x + jb_sky_row(g, s_, e_, r) + c);
        if (exp_in) v = exp(v);      // sky-model tail x = exp(log_sky) (sky_models.py:340-341)
        return v * jb_ldg(wgt + q);
    };
    auto st = [&](int b, int k, cd X) {
        int rr = b / n, i = b - rr * n, r = r0 + rr;
        if (r < N) rowsF[(((size_t)se * n + i) * N + r) * PH + k] = X;
    };
    rows_r2c<FP>(sm, tw_nk, B, ld, st);
}

// ---------------------------------------------------------- adjoint: cotangent rows -> spectra
// grid (ceil(N/rpc), S*T*E).  gin:(S,T,E,N,N) (optionally * expo).  rowsA:(STE, n, N, PH): r2c of the P-wide
// window of cotangent row r for column block i, stored once per (row, i).
template <class FP>
JB_GLOBAL void k_rows_adj_in(JbGeom g, const double* __restrict__ gin, const double* __restrict__ expo,
                             const cd* __restrict__ tw_nk, cd* __restrict__ rowsA, int rpc) {
    JB_SMEM(unsigned char, smraw);
    const int B = rpc * g.n;
    cd* sm = reinterpret_cast<cd*>(smraw);
    int ste, te, se_;
    jb_map_ste(g, blockIdx.y, ste, te, se_);
    const int r0 = blockIdx.x * rpc;
    const int n = g.n, N = g.N, m = g.m, hp = g.hp, PH = g.PH;
    const double* gp = gin + (size_t)ste * N * N;
    const double* ep = expo ? expo + (size_t)te * N * N : nullptr;
    auto ld = [&](int b, int p) -> double {
        int rr = b / n, i = b - rr * n, r = r0 + rr;
        const int ps = jb_fold_src(g, p);       // window position (periodic extension under the alias fold)
        int c = g.offs[i] - hp - m + ps;
        if (r >= N || ps < 0 || c < 0 || c >= N) return 0.0;
        double v = jb_ldg(gp + (size_t)r * N + c);
        if (ep) v *= jb_ldg(ep + (size_t)r * N + c);
        return v;
    };
    auto st = [&](int b, int k, cd X) {
        int rr = b / n, i = b - rr * n, r = r0 + rr;
        if (r < N) rowsA[(((size_t)ste * n + i) * N + r) * PH + k] = X;
    };
    rows_r2c<FP>(sm, tw_nk, B, ld, st);
}

// --------------------------------------------------------------------------- column pass
// grid (nstrips, n*n, planes).  One CTA owns CB spectrum columns of one patch: forward column FFT,
// multiply by the resident PSF spectrum (conj for the adjoint), inverse column FFT.
//   forward : in  rowsF[s*E+e][i] rows offs[j]-hp + (p-m),  p in [m, m+w), times the row weight wgt[p-m]
//             out bufA rows p in [0,P)
//   adjoint : in  rowsA[plane][i] rows offs[j]-hp-m + p,    p in [0,P)
//             out bufB rows p-m, p in [m, m+w)
// Window rows outside the image [0,N) are zero (the overlapping windows of neighbouring patches read the same
// stored rows: nothing is duplicated in HBM).
// Khat layout per (t,e,patch): [strip][k2][k1][b]  -- exactly the register distribution after the
// forward transform, so the load is perfectly coalesced and needs no index arithmetic.
template <class FP, int CB, bool ADJ>
JB_GLOBAL void k_cols(JbGeom g, const cd* __restrict__ in, cd* __restrict__ out,
                      const cd* __restrict__ khat, const cd* __restrict__ tw_nk, const double* __restrict__ wgt) {
    JB_SMEM(cd, sm);
    constexpr int L = FP::L, R1 = FP::R1, R2 = FP::R2;
    const int strip = blockIdx.x, k = blockIdx.y;
    int plane, te, se;
    jb_map_ste(g, blockIdx.z, plane, te, se);
    const int n = g.n, N = g.N, m = g.m, w = g.w, hp = g.hp, PH = g.PH, P = g.P;
    const int j = k % n, i = k / n;
    const size_t nn2 = (size_t)n * n;
    const int nstrips = gridDim.x;
    const int col0 = strip * CB;
    int in_p0, in_rows, row0, out_p0, out_rows;
    size_t in_plane;
    if (!ADJ) {
        in_p0 = m; in_rows = w; out_p0 = 0; out_rows = g.fold ? L : P;   // fold: all L rows go out, the row pass sums the aliases
        in_plane = (size_t)se;
        row0 = g.offs[j] - hp;
    } else {
        in_p0 = 0; in_rows = P; out_p0 = m; out_rows = w;
        in_plane = plane;
        row0 = g.offs[j] - hp - m;
    }
    const cd* ib = in + ((in_plane * n + i) * (size_t)N) * PH;
    cd* ob = out + (((size_t)plane * nn2 + k) * out_rows) * PH;
    const cd* kb = khat + (((size_t)te * nn2 + k) * nstrips + strip) * ((size_t)L * CB);
    auto ld = [&](int p, int b) -> cd {
        int col = col0 + b, u = p - in_p0;
        if (ADJ) u = jb_fold_src(g, p);          // periodic extension of the P cotangent rows under the alias fold
        int r = row0 + u;
        if (col >= PH || u < 0 || u >= in_rows || r < 0 || r >= N) return cd_make(0.0, 0.0);
        cd v = ib[(size_t)r * PH + col];
        if (!ADJ) v = cd_scale(v, jb_ldg(wgt + u));
        return v;
    };
    auto mid = [&](int k1, int k2, int b, cd X) -> cd {
        cd kv = jb_ld_stream(kb + ((size_t)(k2 * R1 + k1)) * CB + b);
        return ADJ ? cd_mulc(X, kv) : cd_mul(X, kv);
    };
    auto st = [&](int p, int b, cd v) {
        int col = col0 + b;
        if (col < PH && p >= out_p0 && p < out_p0 + out_rows) ob[(size_t)(p - out_p0) * PH + col] = v;
    };
    cols_cfft<FP, CB, true>(sm, tw_nk, ld, mid, st);
}

// fixed-order shared-memory tree: the same inputs always give the same bits
template <bool MAX>
JB_D double jb_block_reduce(double v, double* red) {
    int t = jb_tid(), nt = jb_nthr();
    red[t] = v;
    JB_SYNC();
    int p2 = 1;
    while (p2 * 2 <= nt) p2 *= 2;
    if (t + p2 < nt) red[t] = MAX ? (red[t + p2] > red[t] ? red[t + p2] : red[t]) : red[t] + red[t + p2];
    JB_SYNC();
    for (int sft = p2 / 2; sft > 0; sft >>= 1) {
        if (t < sft) red[t] = MAX ? (red[t + sft] > red[t] ? red[t + sft] : red[t]) : red[t] + red[t + sft];
        JB_SYNC();
    }
    double r = red[0];
    JB_SYNC();
    return r;
}
JB_D double jb_block_sum(double v, double* red) { return jb_block_reduce<false>(v, red); }

// --------------------------------------------- forward: spectra -> image rows (+ epilogue)
// grid (ceil(N/rpc), S*T*E).  Sums the spectra of every row-patch j that covers image row r,
// one c2r per (row, i), overlap-add over i as a gather, then the exposure/mask/Poisson epilogue.
template <class FP>
JB_GLOBAL void k_rows_out(JbGeom g, const cd* __restrict__ bufA, const cd* __restrict__ tw_kn,
                          JbEpi ep, int rpc) {
    JB_SMEM(unsigned char, smraw);
    const int B = rpc * g.n;
    const int n = g.n, N = g.N, m = g.m, hp = g.hp, PH = g.PH, P = g.P;
    size_t off = 0;
    cd* sm = reinterpret_cast<cd*>(smraw);
    const int TW = g.fold ? FP::L : P;       // real samples kept per sequence; OR = rows per patch in bufA
    const int OR = TW;
    off += jb_align16((size_t)B * FP::LS * sizeof(cd));
    double* tile = reinterpret_cast<double*>(smraw + off);
    off += jb_align16((size_t)B * TW * sizeof(double));
    double* red = reinterpret_cast<double*>(smraw + off);
    off += jb_align16((size_t)jb_nthr() * sizeof(double));
    int* jcnt = reinterpret_cast<int*>(smraw + off);
    int* jj = jcnt + rpc;
    int* ju = jj + 2 * rpc * n;
    int ste, te, se_;
    jb_map_ste(g, blockIdx.y, ste, te, se_);
    const int r0 = blockIdx.x * rpc;
    jb_build_jlist(g, r0, rpc, hp + m, P, g.fold != 0, jcnt, jj, ju);
    JB_SYNC();
    const size_t nn2 = (size_t)n * n;
    const cd* ab = bufA + (size_t)ste * nn2 * OR * PH;
    auto ld = [&](int b, int k) -> cd {
        int rr = b / n, i = b - rr * n;
        cd acc = cd_make(0.0, 0.0);
        for (int q = 0; q < jcnt[rr]; ++q) {
            int j = jj[rr * 2 * n + q], u = ju[rr * 2 * n + q];
            acc = cd_add(acc, ab[(((size_t)i * n + j) * OR + u) * PH + k]);
        }
        return acc;
    };
    auto st = [&](int b, int p, double v) {
        if (p < TW) tile[(size_t)b * TW + p] = v;
    };
    rows_c2r<FP>(sm, tw_kn, B, ld, st);
    JB_SYNC();
    double esum = 0.0;
    for (int idx = jb_tid(); idx < rpc * N; idx += jb_nthr()) {
        int rr = idx / N, c = idx - rr * N, r = r0 + rr;
        if (r >= N) continue;
        double conv = 0.0;
        for (int i = 0; i < n; ++i) {
            int p = c - (g.offs[i] - hp - m);
            if (p >= 0 && p < P) {
                conv += tile[(size_t)(rr * n + i) * TW + p];
                const int a = jb_fold_alias(g, p);
                if (a >= 0) conv += tile[(size_t)(rr * n + i) * TW + a];
            }
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

// ------------------------------- adjoint: spectra -> sky-gradient rows (sum over modules t)
// grid (ceil(N/rpc), S*E).  bufB:(STE, n*n, w, PH).  grad:(S,E,N,N).
template <class FP>
JB_GLOBAL void k_rows_adj_out(JbGeom g, const cd* __restrict__ bufB, const double* __restrict__ wgt,
                              const cd* __restrict__ tw_kn, double* __restrict__ grad, const double* __restrict__ logsky,
                              int rpc) {
    JB_SMEM(unsigned char, smraw);
    const int B = rpc * g.n;
    const int n = g.n, N = g.N, m = g.m, hp = g.hp, PH = g.PH, w = g.w;
    size_t off = 0;
    cd* sm = reinterpret_cast<cd*>(smraw);
    off += jb_align16((size_t)B * FP::LS * sizeof(cd));
    double* tile = reinterpret_cast<double*>(smraw + off);
    off += jb_align16((size_t)B * w * sizeof(double));
    int* jcnt = reinterpret_cast<int*>(smraw + off);
    int* jj = jcnt + rpc;
    int* ju = jj + 2 * rpc * n;
    int s, e;
    const int r0 = blockIdx.x * rpc, se = jb_map_se(g, blockIdx.y, s, e);
    jb_build_jlist(g, r0, rpc, hp, w, false, jcnt, jj, ju);
    JB_SYNC();
    const size_t nn2 = (size_t)n * n;
    auto ld = [&](int b, int k) -> cd {
        int rr = b / n, i = b - rr * n;
        cd acc = cd_make(0.0, 0.0);
        for (int t = 0; t < g.T; ++t) {
            const cd* bb = bufB + ((size_t)(s * g.T + t) * g.E + e) * nn2 * w * PH;
            for (int q = 0; q < jcnt[rr]; ++q) {
                int j = jj[rr * 2 * n + q], u0 = ju[rr * 2 * n + q];
                double wr = jb_ldg(wgt + u0);
                cd v = bb[(((size_t)i * n + j) * w + u0) * PH + k];
                acc.x += wr * v.x;
                acc.y += wr * v.y;
            }
        }
        return acc;
    };
    auto st = [&](int b, int p, double v) {
        int q = p - m;
        if (q >= 0 && q < w) tile[(size_t)b * w + q] = v * jb_ldg(wgt + q);
    };
    rows_c2r<FP>(sm, tw_kn, B, ld, st);
    JB_SYNC();
    for (int idx = jb_tid(); idx < rpc * N; idx += jb_nthr()) {
        int rr = idx / N, c = idx - rr * N, r = r0 + rr;
        if (r >= N) continue;
        double acc = 0.0;
        for (int i = 0; i < n; ++i) {
            int v = c - (g.offs[i] - hp);
            if (v >= 0 && v < w) acc += tile[(size_t)(rr * n + i) * w + v];
        }
        const size_t o = jb_sky_row(g, s, e, r) + c;
        if (logsky) acc *= exp(jb_ldg(logsky + o));   // VJP of x = exp(log_sky): d/dlog_sky = grad_x * x
        grad[o] = acc;
    }
}

// ============================================================================ plan-time kernels
// PSF preparation, convolve.py:104-131: fftshift -> centre crop (2dx)^2 -> zero-pad m ->
// ifftshift -> / (sum * dvol).  In offset coordinates d in [-dx, dx): value(d_row, d_col) =
// kernel[d mod N] (FULL layout, origin-centred) or cut[d + dx] (CROPPED layout, centred).
struct JbKsrc {
    const double* ptr;
    int cropped;  // 0: (T, n*n, E, N, N) origin-centred ; 1: (T, n*n, E, w, w) centred
    int T, E;
};

JB_D double jb_kval(const JbGeom& g, const JbKsrc& ks, int t, int k, int e, int dr, int dc) {
    size_t base = ((size_t)t * g.n * g.n + k) * ks.E + e;
    if (ks.cropped) {
        return ks.ptr[(base * g.w + (dr + g.dx)) * g.w + (dc + g.dx)];
    } else {
        int a = dr < 0 ? dr + g.N : dr, b = dc < 0 ? dc + g.N : dc;
        return ks.ptr[(base * g.N + a) * g.N + b];
    }
}

// one CTA per (te, k): sum of the cropped kernel and its half-support (max |d| with value != 0)
JB_GLOBAL_S void k_prep_sum(JbGeom g, JbKsrc ks, double* __restrict__ sums, int* __restrict__ support) {
    JB_SMEM(double, red);
    const int k = blockIdx.x, te = blockIdx.y, t = te / ks.E, e = te - t * ks.E;
    double acc = 0.0;
    int sup = 0;
    for (int idx = jb_tid(); idx < g.w * g.w; idx += jb_nthr()) {
        int qr = idx / g.w, qc = idx - qr * g.w;
        int dr = qr - g.dx, dc = qc - g.dx;
        double v = jb_kval(g, ks, t, k, e, dr, dc);
        acc += v;
        if (v != 0.0) {
            int ar = dr < 0 ? -dr : dr, ac = dc < 0 ? -dc : dc;
            int a = ar > ac ? ar : ac;
            sup = a > sup ? a : sup;
        }
    }
    double tot = jb_block_sum(acc, red);
    double mx = jb_block_reduce<true>((double)sup, red);
    if (jb_tid() == 0) {
        sums[(size_t)te * g.n * g.n + k] = tot;
        support[(size_t)te * g.n * g.n + k] = (int)mx;
    }
}

// grid (ceil(w/rpc), TE*n*n): r2c of the w non-zero rows of the prepared kernel (period L).
// scratch:(TE*n*n, w, PH), row q <-> offset d_row = q - dx.
template <class FP>
JB_GLOBAL void k_prep_rows(JbGeom g, JbKsrc ks, const cd* __restrict__ tw_nk, cd* __restrict__ scratch, int rpc) {
    JB_SMEM(cd, sm);
    constexpr int L = FP::L;
    const int q0 = blockIdx.x * rpc, pk = blockIdx.y;
    const int nn2 = g.n * g.n;
    const int te = pk / nn2, k = pk - te * nn2, t = te / ks.E, e = te - t * ks.E;
    auto ld = [&](int b, int p) -> double {
        int q = q0 + b;
        if (q >= g.w) return 0.0;
        int dc;
        if (p < g.dx) dc = p;
        else if (p >= L - g.dx) dc = p - L;
        else return 0.0;
        return jb_kval(g, ks, t, k, e, q - g.dx, dc);
    };
    auto st = [&](int b, int kk, cd X) {
        int q = q0 + b;
        if (q < g.w) scratch[((size_t)pk * g.w + q) * g.PH + kk] = X;
    };
    rows_r2c<FP>(sm, tw_nk, rpc, ld, st);
}

// grid (nstrips, n*n, TE): forward column FFT of the kernel row spectra, normalise by
// 1/(sum * L^2) (the 1/L^2 of the inverse 2-D FFT is folded in here), store in k_cols' layout.
template <class FP, int CB>
JB_GLOBAL void k_prep_cols(JbGeom g, const cd* __restrict__ scratch, const double* __restrict__ sums,
                           const cd* __restrict__ tw_nk, cd* __restrict__ khat) {
    JB_SMEM(cd, sm);
    constexpr int L = FP::L, R1 = FP::R1;
    const int strip = blockIdx.x, k = blockIdx.y, te = blockIdx.z;
    const int nn2 = g.n * g.n, PH = g.PH, dx = g.dx, w = g.w;
    const int nstrips = gridDim.x, col0 = strip * CB;
    const size_t pk = (size_t)te * nn2 + k;
    const cd* ib = scratch + pk * w * PH;
    cd* kb = khat + (pk * nstrips + strip) * ((size_t)L * CB);
    const double scale = 1.0 / (sums[pk] * (double)L * (double)L);
    auto ld = [&](int p, int b) -> cd {
        int col = col0 + b, q;
        if (p < dx) q = p + dx;
        else if (p >= L - dx) q = p - L + dx;
        else return cd_make(0.0, 0.0);
        if (col >= PH) return cd_make(0.0, 0.0);
        return ib[(size_t)q * PH + col];
    };
    auto mid = [&](int k1, int k2, int b, cd X) -> cd {
        kb[((size_t)(k2 * R1 + k1)) * CB + b] = cd_scale(X, scale);
        return X;
    };
    auto st = [&](int, int, cd) {};
    cols_cfft<FP, CB, false>(sm, tw_nk, ld, mid, st);
}

// left-sqrt-metric epilogue: out = ex * t / sqrt(lambda) = sqrt(fac) * t with fac = ex^2 / lambda (0 where masked);
// NIFTy's Poissonian left_sqrt_metric t / sqrt(lambda) pulled back through R, the ex factor being
// R^T's exposure multiply
JB_GLOBAL_S void k_lsm_epi(const double* __restrict__ t, const double* __restrict__ fac, double* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * jb_nthr() + jb_tid(); i < n; i += (size_t)gridDim.x * jb_nthr()) {
        const double f = fac[i];
        out[i] = (f != 0.0) ? sqrt(f) * t[i] : 0.0;
    }
}

// sample-sharded multi-GPU evaluation: flat = [ sum_s grad[s] (E,N,N) | energies of all ranks' samples (nslots) ].
// Folds `cnt` gradient elements of the S local samples (element i of sample s at grad[s*sstride + i]) into dst[0..cnt) in a
// fixed order.  This rank's energies were already written into its slots [slot0, slot0+S) of `slots` by k_reduce_energy;
// the other slots are zeroed here so that ONE all-reduce(sum) delivers the sample-summed gradient and every energy.
// grad == nullptr: the gradient of the single local sample already sits in dst; slots == nullptr: leave the energies alone
// (the per-energy-bin calls of the overlapped collective clear them with the last bin only).
JB_GLOBAL_S void k_fold_samples(const double* __restrict__ grad, int S, size_t cnt, size_t sstride, double* __restrict__ dst,
                                int accumulate, double* __restrict__ slots, int slot0, int nown, int nslots) {
    const size_t t0 = (size_t)blockIdx.x * jb_nthr() + jb_tid(), nt = (size_t)gridDim.x * jb_nthr();
    if (grad) {
        const cd* g2 = reinterpret_cast<const cd*>(grad);   // 128-bit accesses (cnt and sstride are multiples of N*N, N even)
        cd* f2 = reinterpret_cast<cd*>(dst);
        const size_t h = cnt / 2, hs = sstride / 2;
        for (size_t i = t0; i < h; i += nt) {
            cd a = g2[i];
            if (accumulate) { cd o = f2[i]; a.x = o.x + a.x; a.y = o.y + a.y; }   // earlier chunks first: fixed order
            for (int s = 1; s < S; ++s) { cd b = g2[(size_t)s * hs + i]; a.x += b.x; a.y += b.y; }
            f2[i] = a;
        }
    }
    if (slots)
        for (size_t j = t0; j < (size_t)nslots; j += nt)
            if ((int)j < slot0 || (int)j >= slot0 + nown) slots[j] = 0.0;
}

// energy[s] = sum over (t,e) planes and row blocks of the per-CTA partials; one CTA per sample
JB_GLOBAL_S void k_reduce_energy(const double* __restrict__ partial, int per_sample, double* __restrict__ energy) {
    JB_SMEM(double, red);
    const int s = blockIdx.x;
    double acc = 0.0;
    for (int i = jb_tid(); i < per_sample; i += jb_nthr()) acc += partial[(size_t)s * per_sample + i];
    double tot = jb_block_sum(acc, red);
    if (jb_tid() == 0) energy[s] = tot;
}

// expo_eff = (expo < cut ? 0 : expo) * eff_area[t,e] * pixel_area   (response.py:37-44,
// erosita_response.py:327-328,349-350); data int64 -> fp64
JB_GLOBAL_S void k_prep_expo(const double* __restrict__ expo, const double* __restrict__ eff, double cut,
                           int has_cut, double pixel_area, size_t plane, size_t total, double* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * jb_nthr() + jb_tid(); i < total; i += (size_t)gridDim.x * jb_nthr()) {
        double v = expo[i];
        if (has_cut && v < cut) v = 0.0;
        double f = pixel_area;
        if (eff) f *= eff[i / plane];
        out[i] = (v == 0.0) ? 0.0 : v * f;
    }
}

// Content fingerprint of a kernel array (jubik_plan_ensure_kernel: "did the caller pass a different PSF?"): two 64-bit
// wrap-around sums over the bit patterns, the second weighted by the odd number 2*index+1, one pair per CTA in out[2*block]
// (the host adds the pairs; integer sums do not depend on the order).
JB_GLOBAL_S void k_fingerprint(const double* __restrict__ a, size_t total, unsigned long long* __restrict__ out) {
    JB_SMEM(unsigned long long, red);
    unsigned long long h1 = 0, h2 = 0;
    const unsigned long long* bits = reinterpret_cast<const unsigned long long*>(a);
    for (size_t i = (size_t)blockIdx.x * jb_nthr() + jb_tid(); i < total; i += (size_t)gridDim.x * jb_nthr()) {
        const unsigned long long v = bits[i];
        h1 += v;
        h2 += v * (2ull * (unsigned long long)i + 1ull);
    }
    const int t = jb_tid(), nt = jb_nthr();
    red[t] = h1; red[nt + t] = h2;
    JB_SYNC();
    for (int sft = nt / 2; sft > 0; sft >>= 1) {
        if (t < sft) { red[t] += red[t + sft]; red[nt + t] += red[nt + t + sft]; }
        JB_SYNC();
    }
    if (t == 0) { out[2 * blockIdx.x] = red[0]; out[2 * blockIdx.x + 1] = red[nt]; }
}

JB_GLOBAL_S void k_i64_to_f64(const long long* __restrict__ in, size_t total, double* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * jb_nthr() + jb_tid(); i < total; i += (size_t)gridDim.x * jb_nthr())
        out[i] = (double)in[i];
}

// ------------------------------------------------------------------------------------------------------
// Stand-alone circular FFT convolution (jubik/convolve.py:148-176: dvol * Re ifftn(fftn(x) * fftn(y)) over the last two
// axes) for arrays whose side lengths are compiled transform lengths: rows r2c (both operands) -> column pass (forward
// column FFT of y into a scratch slice in register order, forward column FFT of x, product, inverse column FFT) -> rows
// c2r with dvol / (n0*n1) folded in.  Row kernels are instantiated for the row length n1, the column kernel for n0.
template <class FP>
JB_GLOBAL void k_conv_rows_r2c(const double* __restrict__ in, cd* __restrict__ out, const cd* __restrict__ tw_nk,
                               long nrows, int PH, int rpc) {
    JB_SMEM(cd, sm);
    const long r0 = (long)blockIdx.x * rpc;
    const int B = (int)(nrows - r0 < rpc ? nrows - r0 : rpc);
    auto ld = [&](int b, int p) -> double { return jb_ldg(in + (size_t)(r0 + b) * FP::L + p); };
    auto st = [&](int b, int k, cd X) { out[(size_t)(r0 + b) * PH + k] = X; };
    rows_r2c<FP>(sm, tw_nk, B, ld, st);
}
template <class FP>
JB_GLOBAL void k_conv_rows_c2r(const cd* __restrict__ in, double* __restrict__ out, const cd* __restrict__ tw_kn,
                               long nrows, int PH, int rpc, double scale) {
    JB_SMEM(cd, sm);
    const long r0 = (long)blockIdx.x * rpc;
    const int B = (int)(nrows - r0 < rpc ? nrows - r0 : rpc);
    auto ld = [&](int b, int k) -> cd { return in[(size_t)(r0 + b) * PH + k]; };
    auto st = [&](int b, int p, double v) { out[(size_t)(r0 + b) * FP::L + p] = scale * v; };
    rows_c2r<FP>(sm, tw_kn, B, ld, st);
}
// grid (nstrips, lead); xs (x_lead, L, PH), ys (lead, L, PH) row spectra; the result overwrites the strip of ys.
// scratch: (lead, nstrips, L*CB) column spectra of y in the register order of cols_cfft.
template <class FP, int CB>
JB_GLOBAL void k_conv_cols(const cd* __restrict__ xs, cd* __restrict__ ys, cd* __restrict__ scratch,
                           const cd* __restrict__ tw_nk, int PH, int x_lead) {
    JB_SMEM(cd, sm);
    constexpr int L = FP::L, R1 = FP::R1;
    const int strip = blockIdx.x, l = blockIdx.y, col0 = strip * CB;
    const cd* xb = xs + (size_t)(x_lead == 1 ? 0 : l) * L * PH;
    cd* yb = ys + (size_t)l * L * PH;
    cd* sc = scratch + ((size_t)l * gridDim.x + strip) * ((size_t)L * CB);
    auto ld_y = [&](int p, int b) -> cd { int c = col0 + b; return c < PH ? yb[(size_t)p * PH + c] : cd_make(0.0, 0.0); };
    auto ld_x = [&](int p, int b) -> cd { int c = col0 + b; return c < PH ? xb[(size_t)p * PH + c] : cd_make(0.0, 0.0); };
    auto keep = [&](int k1, int k2, int b, cd Y) -> cd { sc[(size_t)(k2 * R1 + k1) * CB + b] = Y; return Y; };
    auto none = [&](int, int, cd) {};
    cols_cfft<FP, CB, false>(sm, tw_nk, ld_y, keep, none);
    JB_SYNC();   // (each thread reads back exactly the slots it wrote; the barrier separates the two uses of sm)
    auto mul = [&](int k1, int k2, int b, cd X) -> cd { return cd_mul(X, sc[(size_t)(k2 * R1 + k1) * CB + b]); };
    auto st = [&](int p, int b, cd v) { int c = col0 + b; if (c < PH) yb[(size_t)p * PH + c] = v; };
    cols_cfft<FP, CB, true>(sm, tw_nk, ld_x, mul, st);
}

// ------------------------------------------------------------------------------------------------------
// Device PSF builder: erosita_psf_utils.py:198-264 psf_interpolator with :94-171 get_psf.  For patch k
// and energy e the PSF tables of the bracketing off-axis radii (host-computed weights wts[e][k][q], radii ascending,
// table index perm[e][q]) are rotated to the patch's position angle (rot[k] = (cos phi, sin phi)) and bilinearly
// resampled (map_coordinates order=1, mode='nearest') at every kernel pixel.  Output (n*n, E, side, side): side == N is
// the reference layout (origin-centred, i.e. rolled by N/2), side == 2dx the centred crop convolve.py:107-111 extracts --
// the (T, n*n, E, N, N) array (17-70 GB at the BASELINE configurations) never has to exist.
struct JbPsfInterp {
    int E, nrad, px, py, N, nn2, side, cropped, dx;
    double d0, d1, delta0, delta1;
    const double* psfs;   // (E, nrad, px, py)
    const double* wts;    // (E, nn2, nrad)
    const double* cid;    // (E, nrad, 2) sorted like the weights
    const int* perm;      // (E, nrad): table index of sorted radius q
    const double* rot;    // (nn2, 2)
};
JB_GLOBAL_S void k_psf_interp(JbPsfInterp q, double* __restrict__ out) {
    const int k = blockIdx.y, e = blockIdx.z;
    const size_t npix = (size_t)q.side * q.side;
    const double cs = q.rot[2 * k], sn = q.rot[2 * k + 1];
    const double* wt = q.wts + ((size_t)e * q.nn2 + k) * q.nrad;
    double* ob = out + ((size_t)k * q.E + e) * npix;
    for (size_t idx = (size_t)blockIdx.x * jb_nthr() + jb_tid(); idx < npix; idx += (size_t)gridDim.x * jb_nthr()) {
        const int u = (int)(idx / q.side), v = (int)(idx - (size_t)u * q.side);
        int dr, dc;
        if (q.cropped) { dr = u - q.dx; dc = v - q.dx; }
        else { dr = u < q.N / 2 ? u : u - q.N; dc = v < q.N / 2 ? v : v - q.N; }
        const double d_dec = (double)dr * q.d0, d_ra = -1.0 * ((double)dc * q.d1);
        const double x = d_ra * cs + d_dec * sn, y = d_dec * cs - d_ra * sn;      // (d_ra + i d_dec) * exp(-i phi)
        double acc = 0.0;
        for (int r = 0; r < q.nrad; ++r) {
            const double w = wt[r];
            if (w == 0.0) continue;
            const double* tab = q.psfs + ((size_t)e * q.nrad + q.perm[e * q.nrad + r]) * ((size_t)q.px * q.py);
            const double c0 = q.cid[((size_t)e * q.nrad + r) * 2], c1 = q.cid[((size_t)e * q.nrad + r) * 2 + 1];
            double a = (x + c0 * q.delta0) / q.delta0, b = (y + c1 * q.delta1) / q.delta1;
            a = a < 0.0 ? 0.0 : (a > (double)(q.px - 1) ? (double)(q.px - 1) : a);     // mode='nearest'
            b = b < 0.0 ? 0.0 : (b > (double)(q.py - 1) ? (double)(q.py - 1) : b);
            const int i0 = (int)floor(a), j0 = (int)floor(b);
            const int i1 = i0 + 1 < q.px ? i0 + 1 : q.px - 1, j1 = j0 + 1 < q.py ? j0 + 1 : q.py - 1;
            const double fa = a - (double)i0, fb = b - (double)j0;
            const double v00 = tab[(size_t)i0 * q.py + j0], v01 = tab[(size_t)i0 * q.py + j1];
            const double v10 = tab[(size_t)i1 * q.py + j0], v11 = tab[(size_t)i1 * q.py + j1];
            acc += w * ((1.0 - fa) * ((1.0 - fb) * v00 + fb * v01) + fa * ((1.0 - fb) * v10 + fb * v11));
        }
        ob[idx] = acc;
    }
}
