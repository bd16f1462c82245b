// capi.cu -- the C ABI declared in include/jubik_b200.h: plan creation (PSF spectra precomputed and kept
// resident, exposure/mask/data folded into dense device planes) and the per-call kernel sequences.
#include "ops.cuh"
#include "fft_list.h"
#include "../../include/jubik_b200.h"
#include <algorithm>
#include <functional>
#include <string>
#include <vector>
#include <new>

#define JB_X_DECL(L, R1, R2, CB, PIPE, Q1, Q2, Q3) JB_DECLARE_FFT(L)
JB_FFT_LIST(JB_X_DECL)
#undef JB_X_DECL

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#ifndef JB_EMU
int dev_malloc(void** p, size_t n) { return cudaMalloc(p, n) == cudaSuccess ? 0 : 1; }
void dev_free(void* p) { if (p) cudaFree(p); }
int h2d(void* d, const void* h, size_t n, jb_stream_t st) { return cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, st) != cudaSuccess; }
int d2h(void* h, const void* d, size_t n, jb_stream_t st) { return cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st) != cudaSuccess; }
int stream_sync(jb_stream_t st) { return cudaStreamSynchronize(st) != cudaSuccess; }
std::string cuda_msg(cudaError_t e) { return std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e); }
int device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
#else
int dev_malloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 1; }
void dev_free(void* p) { free(p); }
int h2d(void* d, const void* h, size_t n, jb_stream_t) { memcpy(d, h, n); return 0; }
int d2h(void* h, const void* d, size_t n, jb_stream_t) { memcpy(h, d, n); return 0; }
int stream_sync(jb_stream_t) { return 0; }
std::string cuda_msg(cudaError_t) { return "emu"; }
int device_count() { return 1; }
#endif

// 3-D fp64 TMA view of a row-spectrum buffer (slabs, N, PH) complex: box = CB columns x BR rows x 1 slab.
int make_tile_src(JbTileSrc* ts, const cd* base, size_t slabs, int N, int PH, int CB, int BR, std::string* err) {
    memset(ts, 0, sizeof(*ts));
    ts->base = base; ts->PH = PH; ts->BR = BR; ts->N = N;
#ifndef JB_EMU
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            *err = "cuTensorMapEncodeTiled is not available from the driver";
            return 1;
        }
        encode = (encode_fn)fn;
    }
    cuuint64_t gdim[3] = {(cuuint64_t)2 * PH, (cuuint64_t)N, (cuuint64_t)slabs};
    cuuint64_t gstr[2] = {(cuuint64_t)PH * sizeof(cd), (cuuint64_t)N * PH * sizeof(cd)};
    cuuint32_t box[3] = {(cuuint32_t)(2 * CB), (cuuint32_t)BR, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&ts->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r); return 1; }
#else
    (void)slabs; (void)CB; (void)err;
#endif
    return 0;
}

const JbFftOps* find_ops(int L) {
#define JB_X_FIND(LL, R1, R2, CB, PIPE, Q1, Q2, Q3) if (L == LL) return jb_fft_ops_##LL();
    JB_FFT_LIST(JB_X_FIND)
#undef JB_X_FIND
    return nullptr;
}
const int k_lengths[] = {
#define JB_X_LEN(LL, R1, R2, CB, PIPE, Q1, Q2, Q3) LL,
    JB_FFT_LIST(JB_X_LEN)
#undef JB_X_LEN
};
}  // namespace

struct jubik_plan {
    JbGeom g;
    const JbFftOps* ops = nullptr;
    int S_max = 0;
    int exact = 0, support = 0;
    int has_expo = 0, has_data = 0;
    int rpc_in = 1, rpc_out = 1, thr_rows = 128, thr_cols = 128;
    long long launches = 0;
    // device memory
    cd* khat = nullptr;
    cd* tw_nk = nullptr;
    cd* tw_kn = nullptr;
    cd* tw1 = nullptr;         // three-stage column pass: W_L^{m k1}  [k1*M + m]
    cd* tw2 = nullptr;         //                          W_M^{n3 k2} [k2*R3 + n3]
    double* wgt = nullptr;
    double* expo = nullptr;    // (T,E,N,N) effective exposure, 0 == masked
    double* data = nullptr;    // (T,E,N,N)
    // bufA: forward  column-pass output (STE, n^2, P, PH) -> rows_out;  adjoint row spectra rowsA (STE, n, N, PH)
    // bufB: adjoint  column-pass output (STE, n^2, w, PH) -> rows_adj_out; forward row spectra rowsF (S*E, n, N, PH)
    // (each direction uses one buffer as its de-duplicated input and the other as its per-patch output)
    cd* bufA = nullptr;
    cd* bufB = nullptr;
    double* gbuf = nullptr;    // (S_max,T,E,N,N) cotangent / lambda scratch
    double* lbuf = nullptr;    // (S_max,T,E,N,N) lambda for the metric (allocated on first use)
    double* partial = nullptr; // energy partials
    double* xexp = nullptr;    // exp(log sky) of the sky-model tail entry point (allocated on first use)
    double* xdev = nullptr;    // staging for the host-buffer entry point (allocated on first use)
    double* gdev = nullptr;
    double* edev = nullptr;
    size_t bytes_khat = 0, bytes_ws = 0, bytes_total = 0;
    int nblk_out = 0;
    int use_pipe = 0, n_sm = 1;
    int split_epi = 1, nblk_epi = 1;  // Poisson/metric epilogue as a separate elementwise kernel
    int rows_r3_out = 0, rows_r3_adj = 0;  // packed-complex three-stage c2r kernels (rows3.cuh)
    int rows6 = 0;                         // warp-private r2c kernels (rows6.cuh)
    int sweep = 0;                         // column pass accumulates over the row-patches (cols3.cuh SWEEP + rows5 presum)
    int chunk_e = 0;                       // fused energy+gradient runs one energy bin at a time (L2 residency)
    int cols_pp = 0;                       // column pass as an anti-phase pair of groups per SM (cols3_pp.cuh)
    int dev_planes = 0;                    // ... one (sample, bin) plane at a time on two alternating streams
    int rows5 = 0;                         // warp-private c2r kernels with the fused Poisson epilogue (rows5.cuh)
    int* gtab_fwd = nullptr;               // rows5 overlap-add gather tables [gq][N]
    int* gtab_adj = nullptr;
    unsigned* counter = nullptr;           // last-CTA-done counter of the fused energy reduction
    int gq_fwd = 0, gq_adj = 0;
    int rows_pipe_out = 0, rows_pipe_adj = 0, qmax = 0;  // persistent bulk-copy-fed c2r kernels (rows_pipe.cuh)
    int fast_rows = 0;         // rows_fast.cuh kernels (two-pass length and <= JB_MAXQ row-patches per image row)
    JbTileSrc src_fwd, src_adj;  // TMA views of bufB / bufA for the pipelined column pass
    // second spectrum workspace (one sample) of the host-buffer entry point's second compute stream (first use)
    cd* bufA2 = nullptr;
    cd* bufB2 = nullptr;
    double* gbuf2 = nullptr;
    int ws2_ready = 0;
    JbTileSrc src_fwd2, src_adj2;
    // per-kernel timing (jubik_poisson_profile): events recorded after every launch while enabled
    int prof_on = 0, prof_n = 0;
#ifndef JB_EMU
    cudaEvent_t prof_ev[48] = {};
    // host-buffer entry point: copy streams and per-energy-bin events (created on first use)
    cudaStream_t st_in = nullptr, st_out = nullptr, st_c2 = nullptr;
    cudaEvent_t *ev_in = nullptr, *ev_c = nullptr, ev_start = nullptr;
#endif
};

static inline void jb_mark(jubik_plan* p, jb_stream_t st) {
#ifndef JB_EMU
    if (p->prof_on && p->prof_n < 48) cudaEventRecord(p->prof_ev[p->prof_n++], st);
#else
    (void)p; (void)st;
#endif
}

#define JB_CK(call)                                                           \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return fail(JUBIK_ERR_CUDA, std::string(#call) + " -> " + cuda_msg(e_)); \
    } while (0)
#define JB_LAUNCHED(p, call) do { JB_CK(call); ++(p)->launches; jb_mark(p, st); } while (0)

extern "C" {

const char* jubik_last_error(void) { return g_err.c_str(); }
int jubik_version(void) { return 100; }
int jubik_fft_length_supported(int length) { return find_ops(length) != nullptr; }
int jubik_device_count(void) { return device_count(); }

int jubik_plan_destroy(jubik_plan* p) {
    if (!p) return JUBIK_OK;
    dev_free(p->khat); dev_free(p->tw_nk); dev_free(p->tw_kn); dev_free(p->tw1); dev_free(p->tw2); dev_free(p->wgt); dev_free(p->expo);
    dev_free(p->data); dev_free(p->bufA); dev_free(p->bufB); dev_free(p->bufA2); dev_free(p->bufB2); dev_free(p->gbuf2); dev_free(p->gbuf); dev_free(p->lbuf);
    dev_free(p->counter); dev_free(p->xexp); dev_free(p->partial); dev_free(p->gtab_fwd); dev_free(p->gtab_adj); dev_free(p->xdev); dev_free(p->gdev); dev_free(p->edev);
#ifndef JB_EMU
    if (p->st_in) {
        for (int e = 0; e < p->S_max * p->g.E; ++e) { cudaEventDestroy(p->ev_in[e]); cudaEventDestroy(p->ev_c[e]); }
        delete[] p->ev_in; delete[] p->ev_c;
        cudaEventDestroy(p->ev_start);
        cudaStreamDestroy(p->st_in); cudaStreamDestroy(p->st_out); if (p->st_c2) cudaStreamDestroy(p->st_c2);
    }
    for (int i = 0; i < 48; ++i) if (p->prof_ev[i]) cudaEventDestroy(p->prof_ev[i]);
#endif
    delete p;
    return JUBIK_OK;
}

static int alloc(jubik_plan* p, void** ptr, size_t bytes, const char* what) {
    if (dev_malloc(ptr, bytes)) return fail(JUBIK_ERR_ALLOC, std::string("device allocation failed for ") + what +
                                                                 " (" + std::to_string(bytes) + " bytes)");
    p->bytes_total += bytes;
    return 0;
}

static void choose_launch_shapes(jubik_plan* p) {
    const JbGeom& g = p->g;
    const JbFftOps* o = p->ops;
    // rows kernels: aim for >= 8 sequences per CTA, <= 256 threads; items per pass ~ B*R2 and B*(R1/2+1)
    int rpc = 1;
    while (rpc * g.n < 8 && rpc < g.N) rpc *= 2;
    if (const char* ev = getenv("JUBIK_B200_RPC")) { int v = atoi(ev); if (v >= 1 && v <= 8) rpc = v; }   // tuning hook
    p->rpc_in = p->rpc_out = rpc;
    int B = rpc * g.n;
    int items = B * (o->R2 > 1 ? (o->R2 > o->R1 / 2 + 1 ? o->R2 : o->R1 / 2 + 1) : 1);
    int thr = ((items + 31) / 32) * 32;
    if (thr < 64) thr = 64;
    if (thr > 256) thr = 256;
    p->thr_rows = thr;
    int citems = o->CB * (o->R2 > 1 ? (o->R1 > o->R2 ? o->R1 : o->R2) : 1);
    int ct = ((citems + 31) / 32) * 32;
    if (ct < 32) ct = 32;
    if (ct > 128) ct = 128;
    p->thr_cols = ct;
}

int jubik_plan_create(jubik_plan** out, const jubik_plan_desc* d, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (!out || !d) return fail(JUBIK_ERR_ARG, "null plan/desc pointer");
    *out = nullptr;
    if (device_count() < 1) return fail(JUBIK_ERR_CUDA, "no CUDA device: jubik_b200 has no CPU fallback");
    if (d->n0 != d->n1) return fail(JUBIK_ERR_UNSUPPORTED, "only square skies are supported (the reference's window "
                                    "weights use shape[0] for both axes, convolve.py:20-25)");
    const int N = d->n0, n = d->n_patches, m = d->margin;
    if (N < 4 || (N & 1)) return fail(JUBIK_ERR_ARG, "spatial size must be even and >= 4 (odd sizes break the "
                                      "reference's centre crop, convolve.py:107-111)");
    if (n < 3) return fail(JUBIK_ERR_ARG, "n_patches_per_axis must be >= 3 (n <= 2 gives an empty kernel crop in the "
                           "reference, convolve.py:107-111)");
    if (n > JB_MAXN) return fail(JUBIK_ERR_UNSUPPORTED, "n_patches_per_axis > 64");
    if (m < 1) return fail(JUBIK_ERR_ARG, "margin must be >= 1 (margin 0 makes the reference's [m:-m] crop empty)");
    if (d->n_energy < 1 || d->n_modules < 1 || d->max_samples < 1) return fail(JUBIK_ERR_ARG, "E, T, S_max must be >= 1");
    if (!d->kernel) return fail(JUBIK_ERR_ARG, "kernel pointer is null");
    if (d->has_exposure_cut && d->exposure_cut < 0) return fail(JUBIK_ERR_ARG, "exposure_cut should be non-negative or None!");
    const int dx = N / n;  // int(N/n), convolve.py:104
    if (dx < 1) return fail(JUBIK_ERR_ARG, "more patches than pixels");
    if (N - 2 * dx < 2) return fail(JUBIK_ERR_ARG, "kernel crop (N-2dx)//2 must be >= 1");
    jubik_plan* p = new (std::nothrow) jubik_plan();
    if (!p) return fail(JUBIK_ERR_ALLOC, "host allocation failed");
    JbGeom& g = p->g;
    memset(&g, 0, sizeof(g));
    g.N = N; g.n = n; g.m = m; g.dx = dx; g.w = 2 * dx; g.hp = dx / 2; g.P = g.w + 2 * m;
    g.S = d->max_samples; g.T = d->n_modules; g.E = d->n_energy;
    g.e0 = 0; g.Es = g.E;
    p->S_max = d->max_samples;
    // window starts with jax.lax.dynamic_slice clamping (convolve.py:55-57)
    {
        int frame = N + 2 * g.hp;
        for (int j = 0; j < n; ++j) {
            int o = j * dx;
            if (o > frame - g.w) o = frame - g.w;
            if (o < 0) o = 0;
            g.offs[j] = o;
        }
    }
    const size_t nn2 = (size_t)n * n, TE = (size_t)g.T * g.E, STE = TE * p->S_max;
    int rc = 0;
#define JB_TRY(expr) do { rc = (expr); if (rc) { jubik_plan_destroy(p); return rc; } } while (0)
#define JB_TRYCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { jubik_plan_destroy(p); \
        return fail(JUBIK_ERR_CUDA, std::string(#call) + " -> " + cuda_msg(e_)); } ++p->launches; } while (0)

    // ---- PSF sums + support (decides the transform length)
    JbKsrc ks;
    ks.ptr = d->kernel; ks.cropped = (d->kernel_layout == JUBIK_KERNEL_CROPPED); ks.T = g.T; ks.E = g.E;
    double* sums = nullptr;
    int* supp = nullptr;
    JB_TRY(alloc(p, (void**)&sums, TE * nn2 * sizeof(double), "psf sums"));
    if (dev_malloc((void**)&supp, TE * nn2 * sizeof(int))) { dev_free(sums); jubik_plan_destroy(p); return fail(JUBIK_ERR_ALLOC, "psf support"); }
    auto cleanup_tmp = [&]() { dev_free(sums); dev_free(supp); };
    {
        cudaError_t e = jb_launch(k_prep_sum, dim3((unsigned)nn2, (unsigned)TE), dim3(256), 256 * sizeof(double), st, g, ks, sums, supp);
        if (e != cudaSuccess) { cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_CUDA, "k_prep_sum -> " + cuda_msg(e)); }
        ++p->launches;
        std::vector<int> hs(TE * nn2);
        if (d2h(hs.data(), supp, hs.size() * sizeof(int), st) || stream_sync(st)) { cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_CUDA, "reading PSF support failed"); }
        int mx = 0;
        for (int v : hs) mx = v > mx ? v : mx;
        p->support = mx;
    }
    // ---- transform length: exact period P if compiled in; else (compact PSF only) the next larger length
    int L = 0;
    if (d->fft_length > 0) {
        if (d->fft_length < g.P) { cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_ARG, "fft_length < P = 2*int(N/n) + 2*margin"); }
        if (!find_ops(d->fft_length)) { cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_UNSUPPORTED, "fft_length " + std::to_string(d->fft_length) + " is not compiled in"); }
        L = d->fft_length;
    } else if (find_ops(g.P)) {
        L = g.P;
    } else {
        for (int c : k_lengths) if (c >= g.P && (L == 0 || c < L)) L = c;
    }
    if (L == 0) { cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_UNSUPPORTED, "no compiled FFT length >= P = " + std::to_string(g.P)); }
    if (L != g.P && p->support > m) {
        cleanup_tmp(); jubik_plan_destroy(p);
        return fail(JUBIK_ERR_UNSUPPORTED, "P = " + std::to_string(g.P) + " is not a compiled FFT length and the PSF half-support (" +
                    std::to_string(p->support) + ") exceeds the margin (" + std::to_string(m) +
                    "): the reference's period-P wrap-around cannot be reproduced with a padded transform");
    }
    p->exact = (L == g.P);
    g.L = L; g.PH = L / 2 + 1;
    p->ops = find_ops(L);
    const JbFftOps* o = p->ops;
    choose_launch_shapes(p);

    // ---- twiddles and window weights (host, long double), uploaded once
    {
        const int R1 = o->R1, R2 = o->R2;
        std::vector<cd> a((size_t)L), b((size_t)L);
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int k1 = 0; k1 < R1; ++k1)
            for (int n2 = 0; n2 < R2; ++n2) {
                long long e = ((long long)k1 * n2) % L;
                long double ang = -two_pi * (long double)e / (long double)L;
                cd v = cd_make((double)cosl(ang), (double)sinl(ang));
                a[(size_t)k1 * R2 + n2] = v;
                b[(size_t)n2 * R1 + k1] = v;
            }
        JB_TRY(alloc(p, (void**)&p->tw_nk, L * sizeof(cd), "twiddles"));
        JB_TRY(alloc(p, (void**)&p->tw_kn, L * sizeof(cd), "twiddles"));
        // weights: convolve.py:18-25  b = concat(linspace(0,1,w/2), reversed)
        std::vector<double> wv((size_t)g.w);
        int half = g.w / 2;
        for (int i = 0; i < half; ++i) {
            // numpy linspace: start + i*step with step = 1/(half-1), last sample forced to 1; linspace(0,1,1) == [0.]
            double v = 0.0;
            if (half > 1) { double step = 1.0 / (double)(half - 1); v = (i == half - 1) ? 1.0 : (double)i * step; }
            wv[i] = v; wv[g.w - 1 - i] = v;
        }
        JB_TRY(alloc(p, (void**)&p->wgt, g.w * sizeof(double), "weights"));
        if (h2d(p->tw_nk, a.data(), L * sizeof(cd), st) || h2d(p->tw_kn, b.data(), L * sizeof(cd), st) ||
            h2d(p->wgt, wv.data(), g.w * sizeof(double), st) || stream_sync(st)) {
            cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_CUDA, "uploading twiddles failed");
        }
    }
    // ---- workspace
    const size_t nstrips = (size_t)jb_cdiv(g.PH, o->CB);
    p->bytes_khat = TE * nn2 * nstrips * (size_t)L * o->CB * sizeof(cd);
    size_t bA = STE * nn2 * (size_t)g.P * g.PH * sizeof(cd);
    size_t bB = STE * nn2 * (size_t)g.w * g.PH * sizeof(cd);
    size_t bG = STE * (size_t)N * N * sizeof(double);
    p->nblk_out = jb_cdiv(N, p->rpc_out);
    p->split_epi = getenv("JUBIK_B200_FUSED_EPI") ? 0 : 1;
    p->nblk_epi = jb_cdiv(N * N, 256 * 8);
    if (p->nblk_epi > 1024) p->nblk_epi = 1024;
    size_t bP = STE * (size_t)(p->nblk_out > p->nblk_epi ? p->nblk_out : p->nblk_epi) * sizeof(double);
    if (bP < STE * (size_t)N * sizeof(double)) bP = STE * (size_t)N * sizeof(double);   // rows5: one partial per image row
    p->bytes_ws = bA + bB + bG + bP;
    if (alloc(p, (void**)&p->khat, p->bytes_khat, "PSF spectra") || alloc(p, (void**)&p->bufA, bA, "spectrum workspace A") ||
        alloc(p, (void**)&p->bufB, bB, "spectrum workspace B") || alloc(p, (void**)&p->gbuf, bG, "cotangent workspace") ||
        alloc(p, (void**)&p->partial, bP, "energy partials")) {
        std::string msg = g_err; cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_ALLOC, msg);
    }
    // ---- low-overhead row kernels need at most JB_MAXQ row-patches covering any image row
    {
        int mx = 0;
        for (int r = 0; r < N; ++r) {
            int c0 = 0, c1 = 0;
            for (int j = 0; j < n; ++j) {
                int u0 = r + g.hp - g.offs[j], u1 = r + g.hp + m - g.offs[j];
                if (u0 >= 0 && u0 < g.w) ++c0;
                if (u1 >= 0 && u1 < g.P) ++c1;
            }
            mx = c0 > mx ? c0 : mx;
            mx = c1 > mx ? c1 : mx;
        }
        p->fast_rows = (o->R2 > 1) && mx <= JB_MAXQ && !getenv("JUBIK_B200_ROWS_V1");
        p->qmax = mx;
    }
    // ---- pipelined column pass: TMA views of the two spectrum workspaces
    p->use_pipe = getenv("JUBIK_B200_COLS_V1") ? 0 : o->pipe;
    if (p->use_pipe == 2) {
        const int Q1 = o->q1, Q2 = o->q2, Q3 = o->q3, M = Q2 * Q3;
        std::vector<cd> a((size_t)L), b((size_t)M);
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int k1 = 0; k1 < Q1; ++k1)
            for (int mm = 0; mm < M; ++mm) {
                long double ang = -two_pi * (long double)(((long long)k1 * mm) % L) / (long double)L;
                a[(size_t)k1 * M + mm] = cd_make((double)cosl(ang), (double)sinl(ang));
            }
        for (int k2 = 0; k2 < Q2; ++k2)
            for (int n3 = 0; n3 < Q3; ++n3) {
                long double ang = -two_pi * (long double)(((long long)k2 * n3) % M) / (long double)M;
                b[(size_t)k2 * Q3 + n3] = cd_make((double)cosl(ang), (double)sinl(ang));
            }
        if (alloc(p, (void**)&p->tw1, L * sizeof(cd), "twiddles") || alloc(p, (void**)&p->tw2, M * sizeof(cd), "twiddles") ||
            h2d(p->tw1, a.data(), L * sizeof(cd), st) || h2d(p->tw2, b.data(), M * sizeof(cd), st) || stream_sync(st)) {
            cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_CUDA, "uploading three-stage twiddles failed");
        }
    }
    if (p->use_pipe && p->fast_rows && !getenv("JUBIK_B200_ROWS_V2")) {
        const size_t smem_max = 227 * 1024;
        p->rows_pipe_out = o->rows_out_pipe_smem(g, p->qmax) <= smem_max;
        p->rows_pipe_adj = o->rows_adj_out_pipe_smem(g, p->qmax) <= smem_max;
        if (p->use_pipe == 2 && !getenv("JUBIK_B200_ROWS_V3")) {
            // two CTAs per SM: each may use half of the 227 KB (minus the 1 KB per-CTA reservation)
            p->rows_r3_out = o->rows_out_r3_smem(g, p->qmax) <= smem_max / 2 - 1024;
            p->rows_r3_adj = o->rows_adj_out_r3_smem(g, p->qmax) <= smem_max / 2 - 1024;
        }
    }
    if (p->use_pipe == 2 && p->qmax <= JB_MAXQ && !getenv("JUBIK_B200_ROWS_V4"))
        p->rows5 = o->rows5_smem(g) <= (size_t)227 * 1024 - 1024;
    if (p->rows5) {
        std::vector<int> tf((size_t)JB_MAXQ * N), ta((size_t)JB_MAXQ * N);
        p->gq_fwd = o->rows5_gather(g, 0, JB_MAXQ, tf.data());
        p->gq_adj = o->rows5_gather(g, 1, JB_MAXQ, ta.data());
        if (p->gq_fwd < 1 || p->gq_adj < 1) {
            p->rows5 = 0;
        } else if (alloc(p, (void**)&p->gtab_fwd, tf.size() * sizeof(int), "gather table") ||
                   alloc(p, (void**)&p->gtab_adj, ta.size() * sizeof(int), "gather table") ||
                   h2d(p->gtab_fwd, tf.data(), tf.size() * sizeof(int), st) ||
                   h2d(p->gtab_adj, ta.data(), ta.size() * sizeof(int), st) || stream_sync(st)) {
            cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_CUDA, "uploading gather tables failed");
        }
    }
    p->chunk_e = getenv("JUBIK_B200_CHUNK_E") ? 1 : 0;
    p->dev_planes = getenv("JUBIK_B200_DEV_PLANES") ? 1 : 0;
    p->cols_pp = (!p->sweep && getenv("JUBIK_B200_COLS_PP")) ? 1 : 0;
    // warp-private r2c kernels: parity-tested, but measured slower than the two-pass kernels of rows_fast.cuh on B200
    // (forward 69 vs 36 us -- the two-pass kernel skips the zero padding --, adjoint 58 vs 57 us): opt-in
    p->rows6 = (p->rows5 && getenv("JUBIK_B200_ROWS_R2C_V6") && o->rows6_smem(g) <= (size_t)227 * 1024 - 1024) ? 1 : 0;
    // measured on B200 (C2): the read-add-write costs the column pass more (+48 us per direction) than the row pass
    // gains from reading one summed row (-22 / -13 us), so the sweep stays an opt-in experiment
    p->sweep = (p->rows5 && getenv("JUBIK_B200_SWEEP")) ? 1 : 0;
    if (p->rows5 && !getenv("JUBIK_B200_SPLIT_EPI")) { p->split_epi = 0; p->nblk_out = N; }   // fused epilogue: one partial per row
    // fused final reduction (last CTA of the forward c2r kernel sums the partials): measured equal to the separate
    // 8 us k_reduce_energy launch (the lone last CTA adds the same tail), so it stays opt-in
    if (p->rows5 && !p->split_epi && getenv("JUBIK_B200_FUSED_REDUCE")) {
        unsigned zero = 0;
        if (alloc(p, (void**)&p->counter, sizeof(unsigned), "reduction counter") || h2d(p->counter, &zero, sizeof(zero), st) ||
            stream_sync(st)) {
            cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_CUDA, "allocating the reduction counter failed");
        }
    }
    if (p->use_pipe) {
        std::string err;
        int brB = g.w < 64 ? g.w : 64, brA = g.P < 64 ? g.P : 64;
        if (brB > N) brB = N;
        if (brA > N) brA = N;
        // the row-spectrum buffers (slabs = planes*n, N rows, PH) live in the first part of bufB (forward) / bufA (adjoint)
        if (make_tile_src(&p->src_fwd, p->bufB, STE * (size_t)n, N, g.PH, o->CB, brB, &err) ||
            make_tile_src(&p->src_adj, p->bufA, STE * (size_t)n, N, g.PH, o->CB, brA, &err)) {
            cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_CUDA, err);
        }
#ifndef JB_EMU
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        p->n_sm = sms > 0 ? sms : 148;
#else
        p->n_sm = 2;
#endif
    }
    // ---- PSF spectra: rows r2c -> columns FFT -> normalise -> resident in k_cols' register order
    {
        int prpc = 8;
        int items = prpc * (o->R2 > 1 ? (o->R2 > o->R1 / 2 + 1 ? o->R2 : o->R1 / 2 + 1) : 1);
        int thr = ((items + 31) / 32) * 32; if (thr < 32) thr = 32; if (thr > 256) thr = 256;
        cudaError_t e = o->prep_rows(g, ks, p->tw_nk, p->bufB, prpc, thr, st);
        if (e == cudaSuccess) { ++p->launches; e = o->prep_cols(g, p->bufB, sums, p->use_pipe == 2 ? p->tw1 : p->tw_nk, p->tw2, p->khat, p->use_pipe == 2 ? 2 : 0, p->thr_cols, st); }
        if (e == cudaSuccess) ++p->launches;
        if (e != cudaSuccess || stream_sync(st)) { cleanup_tmp(); jubik_plan_destroy(p); return fail(JUBIK_ERR_CUDA, "PSF spectrum kernels -> " + cuda_msg(e)); }
    }
    p->bytes_total -= TE * nn2 * sizeof(double);
    cleanup_tmp();
    // ---- exposure / mask / data
    if (d->exposure) {
        size_t tot = TE * (size_t)N * N;
        JB_TRY(alloc(p, (void**)&p->expo, tot * sizeof(double), "exposure"));
        JB_TRYCK(jb_launch(k_prep_expo, dim3(592), dim3(256), 0, st, d->exposure, d->eff_area, d->exposure_cut,
                           (int)(d->has_exposure_cut != 0), d->pixel_area, (size_t)N * N, tot, p->expo));
        p->has_expo = 1;
    }
    if (d->data) {
        rc = jubik_plan_set_data(p, d->data, d->data_dtype, stream_);
        if (rc) { jubik_plan_destroy(p); return rc; }
    }
    if (stream_sync(st)) { jubik_plan_destroy(p); return fail(JUBIK_ERR_CUDA, "plan creation: stream sync failed"); }
    *out = p;
    return JUBIK_OK;
}

int jubik_plan_set_data(jubik_plan* p, const void* data, int dtype, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (!p || !data) return fail(JUBIK_ERR_ARG, "null plan/data");
    size_t tot = (size_t)p->g.T * p->g.E * p->g.N * p->g.N;
    if (!p->data) { int rc = alloc(p, (void**)&p->data, tot * sizeof(double), "data"); if (rc) return rc; }
    if (dtype == JUBIK_DATA_I64) {
        JB_LAUNCHED(p, jb_launch(k_i64_to_f64, dim3(592), dim3(256), 0, st, (const long long*)data, tot, p->data));
    } else if (dtype == JUBIK_DATA_F64) {
#ifndef JB_EMU
        JB_CK(cudaMemcpyAsync(p->data, data, tot * sizeof(double), cudaMemcpyDeviceToDevice, st));
#else
        memcpy(p->data, data, tot * sizeof(double));
#endif
    } else {
        return fail(JUBIK_ERR_ARG, "unknown data dtype");
    }
    p->has_data = 1;
    return JUBIK_OK;
}

int jubik_plan_get_info(const jubik_plan* p, jubik_plan_info* i) {
    if (!p || !i) return fail(JUBIK_ERR_ARG, "null plan/info");
    const JbGeom& g = p->g;
    i->N = g.N; i->n = g.n; i->m = g.m; i->dx = g.dx; i->w = g.w; i->P = g.P; i->L = g.L; i->PH = g.PH;
    i->S_max = p->S_max; i->T = g.T; i->E = g.E; i->exact_period = p->exact; i->psf_half_support = p->support;
    i->fft_r1 = p->ops->R1; i->fft_r2 = p->ops->R2; i->col_block = p->ops->CB;
    i->bytes_khat = p->bytes_khat; i->bytes_workspace = p->bytes_ws; i->bytes_total = p->bytes_total;
    return JUBIK_OK;
}

long long jubik_plan_launch_count(const jubik_plan* p) { return p ? p->launches : 0; }

static int check_call(jubik_plan* p, const void* a, const void* b, int S) {
    if (!p) return fail(JUBIK_ERR_ARG, "null plan");
    if (!a || !b) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (S < 1 || S > p->S_max) return fail(JUBIK_ERR_ARG, "n_samples must be in [1, max_samples]");
    return 0;
}

// forward chain with the chosen epilogue
// `g` may be a copy of p->g restricted to an energy-bin chunk (e0, Es)
static int run_forward_g(jubik_plan* p, const JbGeom& g, const double* x, int S, const JbEpi& ep, jb_stream_t st) {
    const JbFftOps* o = p->ops;
    if (p->rows6) {
        JB_LAUNCHED(p, o->rows6(g, 0, x, nullptr, p->wgt, p->tw1, p->tw2, p->bufB, S * g.Es, p->n_sm, st));
    } else {
        JB_LAUNCHED(p, o->rows_fwd(g, x, p->wgt, p->tw_nk, p->bufB, S * g.Es, p->rpc_in, p->thr_rows, p->fast_rows, st));
    }
    if (p->use_pipe) {
        JB_LAUNCHED(p, o->cols_pipe(g, false, p->cols_pp ? 2 : p->sweep, p->src_fwd, p->bufA, p->khat, p->tw1, p->tw2, p->wgt, S * g.T * g.Es, p->n_sm, st));
    } else {
        JB_LAUNCHED(p, o->cols(g, false, p->bufB, p->bufA, p->khat, p->tw_nk, p->wgt, S * g.T * g.Es, p->thr_cols, st));
    }
    if (p->rows5) {
        JB_LAUNCHED(p, o->rows5_fwd(g, p->bufA, p->tw1, p->tw2, p->gtab_fwd, p->gq_fwd, ep, p->sweep, S * g.T * g.Es, p->n_sm, st));
    } else if (p->rows_r3_out && ep.mode != JB_EPI_POISSON) {
        JB_LAUNCHED(p, o->rows_out_r3(g, p->bufA, p->tw1, p->tw2, ep, S * g.T * g.Es, p->qmax, p->n_sm, st));
    } else if (p->rows_pipe_out && ep.mode != JB_EPI_POISSON) {
        JB_LAUNCHED(p, o->rows_out_pipe(g, p->bufA, p->tw_kn, ep, S * g.T * g.Es, p->qmax, p->n_sm, st));
    } else {
        JB_LAUNCHED(p, o->rows_out(g, p->bufA, p->tw_kn, ep, S * g.T * g.Es, p->rpc_out, p->thr_rows, p->fast_rows, st));
    }
    return 0;
}
static int run_forward(jubik_plan* p, const double* x, int S, const JbEpi& ep, jb_stream_t st) {
    return run_forward_g(p, p->g, x, S, ep, st);
}

// adjoint chain; expo != nullptr multiplies the incoming cotangent by the effective exposure
static int run_adjoint_g(jubik_plan* p, const JbGeom& g, const double* gin, const double* expo, int S, double* out,
                         jb_stream_t st) {
    const JbFftOps* o = p->ops;
    if (p->rows6) {
        JB_LAUNCHED(p, o->rows6(g, 1, gin, expo, p->wgt, p->tw1, p->tw2, p->bufA, S * g.T * g.Es, p->n_sm, st));
    } else {
        JB_LAUNCHED(p, o->rows_adj_in(g, gin, expo, p->tw_nk, p->bufA, S * g.T * g.Es, p->rpc_in, p->thr_rows, p->fast_rows, st));
    }
    if (p->use_pipe) {
        JB_LAUNCHED(p, o->cols_pipe(g, true, p->cols_pp ? 2 : p->sweep, p->src_adj, p->bufB, p->khat, p->tw1, p->tw2, p->wgt, S * g.T * g.Es, p->n_sm, st));
    } else {
        JB_LAUNCHED(p, o->cols(g, true, p->bufA, p->bufB, p->khat, p->tw_nk, p->wgt, S * g.T * g.Es, p->thr_cols, st));
    }
    if (p->rows5) {
        JB_LAUNCHED(p, o->rows5_adj(g, p->bufB, p->wgt, p->tw1, p->tw2, p->gtab_adj, p->gq_adj, out, p->sweep, S * g.Es, p->n_sm, st));
    } else if (p->rows_r3_adj) {
        JB_LAUNCHED(p, o->rows_adj_out_r3(g, p->bufB, p->wgt, p->tw1, p->tw2, out, S * g.Es, p->qmax, p->n_sm, st));
    } else if (p->rows_pipe_adj) {
        JB_LAUNCHED(p, o->rows_adj_out_pipe(g, p->bufB, p->wgt, p->tw_kn, out, S * g.Es, p->qmax, p->n_sm, st));
    } else {
        JB_LAUNCHED(p, o->rows_adj_out(g, p->bufB, p->wgt, p->tw_kn, out, S * g.Es, p->rpc_out, p->thr_rows, p->fast_rows, st));
    }
    return 0;
}
static int run_adjoint(jubik_plan* p, const double* gin, const double* expo, int S, double* out, jb_stream_t st) {
    return run_adjoint_g(p, p->g, gin, expo, S, out, st);
}

int jubik_psf_apply(jubik_plan* p, const double* x, int S, double* out, void* stream_) {
    if (int rc = check_call(p, x, out, S)) return rc;
    JbEpi ep; memset(&ep, 0, sizeof(ep));
    ep.mode = JB_EPI_PSF; ep.out = out;
    return run_forward(p, x, S, ep, (jb_stream_t)stream_);
}

int jubik_psf_apply_T(jubik_plan* p, const double* gin, int S, double* out, void* stream_) {
    if (int rc = check_call(p, gin, out, S)) return rc;
    return run_adjoint(p, gin, nullptr, S, out, (jb_stream_t)stream_);
}

int jubik_response_apply(jubik_plan* p, const double* x, int S, double* out, void* stream_) {
    if (int rc = check_call(p, x, out, S)) return rc;
    if (!p->has_expo) return fail(JUBIK_ERR_ARG, "plan was created without exposures");
    JbEpi ep; memset(&ep, 0, sizeof(ep));
    ep.mode = JB_EPI_RESP; ep.out = out; ep.expo = p->expo;
    return run_forward(p, x, S, ep, (jb_stream_t)stream_);
}

int jubik_response_apply_T(jubik_plan* p, const double* gin, int S, double* out, void* stream_) {
    if (int rc = check_call(p, gin, out, S)) return rc;
    if (!p->has_expo) return fail(JUBIK_ERR_ARG, "plan was created without exposures");
    return run_adjoint(p, gin, p->expo, S, out, (jb_stream_t)stream_);
}

#ifndef JB_EMU
// streams / events of the plane-pipelined call forms (first use)
static int ensure_plane_streams(jubik_plan* p) {
    if (p->st_in) return 0;
    const int nev = p->S_max * p->g.E;
    JB_CK(cudaStreamCreateWithFlags(&p->st_in, cudaStreamNonBlocking));
    JB_CK(cudaStreamCreateWithFlags(&p->st_out, cudaStreamNonBlocking));
    JB_CK(cudaStreamCreateWithFlags(&p->st_c2, cudaStreamNonBlocking));
    p->ev_in = new cudaEvent_t[nev]; p->ev_c = new cudaEvent_t[nev];
    for (int e = 0; e < nev; ++e) {
        JB_CK(cudaEventCreateWithFlags(&p->ev_in[e], cudaEventDisableTiming));
        JB_CK(cudaEventCreateWithFlags(&p->ev_c[e], cudaEventDisableTiming));
    }
    JB_CK(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
    return 0;
}
// The two compute streams' planes are in different phases at the same time, and the forward / adjoint layouts of one
// spectrum buffer alias across energy bins: the second stream gets its own pair of buffers (one sample's worth).
static int ensure_second_workspace(jubik_plan* p) {
    if (p->ws2_ready) return 0;
    const JbGeom& g = p->g;
    const size_t TE1 = (size_t)g.T * g.E, nn2 = (size_t)g.n * g.n;
    const int brB = std::min(std::min(g.w, 64), g.N), brA = std::min(std::min(g.P, 64), g.N);
    std::string err;
    if ((!p->bufA2 && alloc(p, (void**)&p->bufA2, TE1 * nn2 * (size_t)g.P * g.PH * sizeof(cd), "second spectrum workspace A")) ||
        (!p->bufB2 && alloc(p, (void**)&p->bufB2, TE1 * nn2 * (size_t)g.w * g.PH * sizeof(cd), "second spectrum workspace B")) ||
        (!p->gbuf2 && alloc(p, (void**)&p->gbuf2, TE1 * (size_t)g.N * g.N * sizeof(double), "second cotangent workspace")))
        return JUBIK_ERR_ALLOC;
    if (make_tile_src(&p->src_fwd2, p->bufB2, TE1 * (size_t)g.n, g.N, g.PH, p->ops->CB, brB, &err) ||
        make_tile_src(&p->src_adj2, p->bufA2, TE1 * (size_t)g.n, g.N, g.PH, p->ops->CB, brA, &err))
        return fail(JUBIK_ERR_CUDA, err);
    p->ws2_ready = 1;
    return 0;
}
// kernel arguments are captured at launch, so the second stream's launches simply see the plan with its workspace
// pointers exchanged while they are enqueued
static void swap_workspace(jubik_plan* p) {
    std::swap(p->bufA, p->bufA2); std::swap(p->bufB, p->bufB2);
    std::swap(p->src_fwd, p->src_fwd2); std::swap(p->src_adj, p->src_adj2);
    std::swap(p->gbuf, p->gbuf2);
}
// Fused energy + gradient as (sample, energy bin) planes alternating between the caller's stream and a second one
// (device-resident form of the host entry point's pipeline).  wait_in: per-plane events to wait for (H2D), or null;
// after_plane(c, stream): hook after a plane's kernels were enqueued (the host path issues the D2H there).
static int run_planes_two_streams(jubik_plan* p, const double* x, int S, double* grad, bool two, bool wait_in,
                                  jb_stream_t st, const std::function<int(int)>& after_plane) {
    const JbGeom& g = p->g;
    const size_t npx = (size_t)g.N * g.N;
    if (two) JB_CK(cudaStreamWaitEvent(p->st_c2, p->ev_start, 0));
    // One (sample, energy bin) plane per chunk.  With two streams every stream owns its spectrum buffers AND its
    // cotangent scratch, so chunks on different streams never touch the same memory; chunks on one stream are ordered
    // by the stream.  JUBIK_B200_HOST_PAIR_PLANES=1 pairs up the planes between the first and the last one (fuller
    // waves for the persistent kernels) -- measured slower end to end (1.88 vs 1.61 ms at C2, S = 2: the coarser
    // chunks start later and return their gradients later), kept as a tested option.
    const int nplanes = S * g.E;
    const bool pair_up = two && getenv("JUBIK_B200_HOST_PAIR_PLANES");
    int last_c2 = -1, k = 0;
    for (int c = 0; c < nplanes; ++k) {
        const int si = c / g.E, e = c - si * g.E;
        int len = 1;
        if (pair_up && c > 0 && e + 1 < g.E && c + 1 < nplanes - 1) len = 2;
        jb_stream_t cs = (two && (k & 1)) ? (jb_stream_t)p->st_c2 : st;
        const int cl = c + len - 1;     // last plane of the chunk: its event stands for the chunk
        if (cs != st) last_c2 = cl;
        JbGeom ge = g;
        ge.e0 = e; ge.Es = len;
        if (wait_in) for (int q = c; q <= cl; ++q) JB_CK(cudaStreamWaitEvent(cs, p->ev_in[q], 0));
        JbEpi ep; memset(&ep, 0, sizeof(ep));
        if (cs != st) swap_workspace(p);
        ep.mode = JB_EPI_POISSON; ep.out = p->gbuf; ep.expo = p->expo; ep.data = p->data;
        ep.partial = p->partial + (size_t)si * g.T * g.E * p->nblk_out;
        // the chunk runs as "sample 0" of a one-sample batch on this sample's slices of x / grad / partials
        int rc = run_forward_g(p, ge, x + (size_t)si * g.E * npx, 1, ep, cs);
        if (!rc) rc = run_adjoint_g(p, ge, p->gbuf, nullptr, 1, grad + (size_t)si * g.E * npx, cs);
        if (cs != st) swap_workspace(p);
        if (rc) return rc;
        for (int q = c; q <= cl; ++q) {
            JB_CK(cudaEventRecord(p->ev_c[q], cs));
            if (int rc2 = after_plane(q)) return rc2;
        }
        c += len;
    }
    if (last_c2 >= 0) JB_CK(cudaStreamWaitEvent(st, p->ev_c[last_c2], 0));   // join before the energy reduction
    return 0;
}
#endif

int jubik_poisson_energy_and_grad(jubik_plan* p, const double* x, int S, double* energy, double* grad, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (int rc = check_call(p, x, energy, S)) return rc;
    if (!p->has_expo || !p->has_data) return fail(JUBIK_ERR_ARG, "plan needs exposures and data for the Poisson energy");
    const JbGeom& g = p->g;
    JbEpi ep; memset(&ep, 0, sizeof(ep));
#ifndef JB_EMU
    if (p->dev_planes && grad && !p->split_epi && p->rows5 && p->use_pipe && !p->counter && S * g.E > 1 && !p->prof_on) {
        // JUBIK_B200_DEV_PLANES=1 (experiment): one (sample, energy bin) plane at a time on two alternating streams,
        // so that a plane's spectra are still L2-resident when its next kernel reads them
        if (int rc = ensure_plane_streams(p)) return rc;
        if (int rc = ensure_second_workspace(p)) return rc;
        JB_CK(cudaEventRecord(p->ev_start, st));
        if (int rc = run_planes_two_streams(p, x, S, grad, true, false, st, [](int) { return 0; })) return rc;
        JB_LAUNCHED(p, jb_launch(k_reduce_energy, dim3(S), dim3(256), 256 * sizeof(double), st,
                                 (const double*)p->partial, g.T * g.E * p->nblk_out, energy));
        return JUBIK_OK;
    }
#endif
    if (p->chunk_e && !p->split_epi && grad && g.E > 1) {
        // one energy bin at a time, forward then adjoint: the bin's spectra (column-pass output -> row pass) and its
        // cotangent plane are still L2-resident when the next kernel reads them
        ep.mode = JB_EPI_POISSON; ep.out = p->gbuf; ep.expo = p->expo; ep.data = p->data; ep.partial = p->partial;
        for (int e = 0; e < g.E; ++e) {
            JbGeom ge = g;
            ge.e0 = e; ge.Es = 1;
            if (p->counter && e == g.E - 1) {
                ep.energy = energy; ep.counter = p->counter; ep.n_samples = S; ep.per_sample = g.T * g.E * p->nblk_out;
            }
            if (int rc = run_forward_g(p, ge, x, S, ep, st)) return rc;
            if (int rc = run_adjoint_g(p, ge, p->gbuf, nullptr, S, grad, st)) return rc;
        }
        if (!p->counter)
            JB_LAUNCHED(p, jb_launch(k_reduce_energy, dim3(S), dim3(256), 256 * sizeof(double), st,
                                     (const double*)p->partial, g.T * g.E * p->nblk_out, energy));
        return JUBIK_OK;
    }
    if (p->split_epi) {
        // convolved planes first, then the Poisson epilogue as a separate full-occupancy elementwise pass
        ep.mode = JB_EPI_PSF; ep.out = p->gbuf;
        if (int rc = run_forward(p, x, S, ep, st)) return rc;
        JB_LAUNCHED(p, jb_launch(k_poisson_epi, dim3(p->nblk_epi, S * g.T * g.E), dim3(256), 256 * sizeof(double), st,
                                 g, p->gbuf, (const double*)p->expo, (const double*)p->data, g.N * g.N, p->partial));
        JB_LAUNCHED(p, jb_launch(k_reduce_energy, dim3(S), dim3(256), 256 * sizeof(double), st,
                                 (const double*)p->partial, g.T * g.E * p->nblk_epi, energy));
    } else {
        ep.mode = JB_EPI_POISSON; ep.out = p->gbuf; ep.expo = p->expo; ep.data = p->data; ep.partial = p->partial;
        if (p->counter) {
            ep.energy = energy; ep.counter = p->counter; ep.n_samples = S; ep.per_sample = g.T * g.E * p->nblk_out;
        }
        if (int rc = run_forward(p, x, S, ep, st)) return rc;
        if (!p->counter)
            JB_LAUNCHED(p, jb_launch(k_reduce_energy, dim3(S), dim3(256), 256 * sizeof(double), st,
                                     (const double*)p->partial, g.T * g.E * p->nblk_out, energy));
    }
    if (grad) return run_adjoint(p, p->gbuf, nullptr, S, grad, st);
    return JUBIK_OK;
}

// Same work as jubik_poisson_energy_and_grad with a CUDA event after every kernel; ms_out[0..*n_out) in launch
// order: rows_fwd, cols_fwd, rows_out, [poisson_epi,] reduce, rows_adj_in, cols_adj, rows_adj_out.  Synchronises.
int jubik_poisson_profile(jubik_plan* p, const double* x, int S, double* energy, double* grad, float* ms_out, int* n_out, void* stream_) {
#ifndef JB_EMU
    jb_stream_t st = (jb_stream_t)stream_;
    if (!ms_out || !grad || !n_out) return fail(JUBIK_ERR_ARG, "null pointer");
    for (int i = 0; i < 48; ++i) if (!p->prof_ev[i]) JB_CK(cudaEventCreate(&p->prof_ev[i]));
    p->prof_on = 1; p->prof_n = 0;
    jb_mark(p, st);
    int rc = jubik_poisson_energy_and_grad(p, x, S, energy, grad, stream_);
    p->prof_on = 0;
    if (rc) return rc;
    JB_CK(cudaStreamSynchronize(st));
    for (int i = 0; i + 1 < p->prof_n && i < 47; ++i) JB_CK(cudaEventElapsedTime(&ms_out[i], p->prof_ev[i], p->prof_ev[i + 1]));
    *n_out = p->prof_n - 1;
    return JUBIK_OK;
#else
    (void)p; (void)x; (void)S; (void)energy; (void)grad; (void)ms_out; (void)n_out; (void)stream_;
    return fail(JUBIK_ERR_UNSUPPORTED, "profiling needs CUDA");
#endif
}

// Measured FP64 FMA throughput of this device (DFMA micro-benchmark: 8 independent chains per thread),
// the second roofline denominator of SURVEY 8(d).  Synchronises.
JB_GLOBAL_S void k_dfma_peak(double* out, int iters) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999999, c = 1e-12;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;
}

int jubik_measure_fp64_tflops(double* tflops_out, void* stream_) {
#ifndef JB_EMU
    jb_stream_t st = (jb_stream_t)stream_;
    if (!tflops_out) return fail(JUBIK_ERR_ARG, "null pointer");
    double* d = nullptr;
    JB_CK(cudaMalloc((void**)&d, 8));
    cudaEvent_t a, b;
    JB_CK(cudaEventCreate(&a)); JB_CK(cudaEventCreate(&b));
    const int blocks = 148 * 8, threads = 256, iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        JB_CK(cudaEventRecord(a, st));
        k_dfma_peak<<<blocks, threads, 0, st>>>(d, iters);
        JB_CK(cudaEventRecord(b, st));
        JB_CK(cudaEventSynchronize(b));
        float ms = 0.f;
        JB_CK(cudaEventElapsedTime(&ms, a, b));
        double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    *tflops_out = best;
    return JUBIK_OK;
#else
    (void)tflops_out; (void)stream_;
    return fail(JUBIK_ERR_UNSUPPORTED, "needs CUDA");
#endif
}

int jubik_poisson_metric_apply(jubik_plan* p, const double* x, const double* v, int S, double* out, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (int rc = check_call(p, x, out, S)) return rc;
    if (!v) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (!p->has_expo) return fail(JUBIK_ERR_ARG, "plan was created without exposures");
    const JbGeom& g = p->g;
    if (!p->lbuf) {
        size_t bG = (size_t)p->S_max * g.T * g.E * g.N * g.N * sizeof(double);
        if (int rc = alloc(p, (void**)&p->lbuf, bG, "lambda workspace")) return rc;
    }
    JbEpi ep; memset(&ep, 0, sizeof(ep));
    ep.mode = JB_EPI_RESP; ep.out = p->lbuf; ep.expo = p->expo;
    if (int rc = run_forward(p, x, S, ep, st)) return rc;
    if (p->split_epi) {
        ep.mode = JB_EPI_PSF; ep.out = p->gbuf; ep.expo = nullptr;
        if (int rc = run_forward(p, v, S, ep, st)) return rc;
        JB_LAUNCHED(p, jb_launch(k_metric_epi, dim3(p->nblk_epi, S * g.T * g.E), dim3(256), 0, st, g, p->gbuf,
                                 (const double*)p->expo, (const double*)p->lbuf, g.N * g.N));
    } else {
        ep.mode = JB_EPI_METRIC; ep.out = p->gbuf; ep.lam = p->lbuf;
        if (int rc = run_forward(p, v, S, ep, st)) return rc;
    }
    return run_adjoint(p, p->gbuf, nullptr, S, out, st);
}

int jubik_poisson_left_sqrt_metric(jubik_plan* p, const double* x, const double* t, int S, double* out, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (int rc = check_call(p, x, out, S)) return rc;
    if (!t) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (!p->has_expo) return fail(JUBIK_ERR_ARG, "plan was created without exposures");
    const JbGeom& g = p->g;
    if (!p->lbuf) {
        size_t bG = (size_t)p->S_max * g.T * g.E * g.N * g.N * sizeof(double);
        if (int rc = alloc(p, (void**)&p->lbuf, bG, "lambda workspace")) return rc;
    }
    JbEpi ep; memset(&ep, 0, sizeof(ep));
    ep.mode = JB_EPI_RESP; ep.out = p->lbuf; ep.expo = p->expo;
    if (int rc = run_forward(p, x, S, ep, st)) return rc;
    JB_LAUNCHED(p, jb_launch(k_lsm_epi, dim3(p->nblk_epi, S * g.T * g.E), dim3(256), 0, st, g, t, (const double*)p->expo,
                             (const double*)p->lbuf, p->gbuf, g.N * g.N));
    return run_adjoint(p, p->gbuf, nullptr, S, out, st);
}

int jubik_poisson_energy_and_grad_logsky(jubik_plan* p, const double* log_sky, int S, double* energy, double* grad,
                                         void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (int rc = check_call(p, log_sky, energy, S)) return rc;
    const JbGeom& g = p->g;
    const size_t cnt = (size_t)S * g.E * g.N * g.N;
    if (!p->xexp) {
        if (int rc = alloc(p, (void**)&p->xexp, (size_t)p->S_max * g.E * g.N * g.N * sizeof(double), "exp(log sky) workspace")) return rc;
    }
    JB_LAUNCHED(p, jb_launch(k_exp, dim3(1184), dim3(256), 0, st, log_sky, p->xexp, cnt));
    if (int rc = jubik_poisson_energy_and_grad(p, p->xexp, S, energy, grad, stream_)) return rc;
    if (grad) JB_LAUNCHED(p, jb_launch(k_mul_inplace, dim3(1184), dim3(256), 0, st, grad, (const double*)p->xexp, cnt));
    return JUBIK_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Stand-alone versions of the reference's public helpers (jubik/convolve.py:28-63 slice_patches and its
// jax.linear_transpose, :148-176 convolve).  The fused path above never materialises the patches; these exist so
// that callers of ju.slice_patches / ju.convolve find the same functions.

struct JbSliceGeom { int lead, h, w, n, dr, dx, dy, sx, sy, hx, hy; };

static int slice_geom(JbSliceGeom& q, int lead, int h, int w, int n, int dr) {
    if (lead < 1 || h < 1 || w < 1 || n < 1 || dr < 0) return fail(JUBIK_ERR_ARG, "slice_patches: bad shape");
    q.lead = lead; q.h = h; q.w = w; q.n = n; q.dr = dr;
    q.dx = (h - 2 * dr) / n; q.dy = (w - 2 * dr) / n;          // int() truncation, convolve.py:43-44
    if (q.dx < 1 || q.dy < 1) return fail(JUBIK_ERR_ARG, "slice_patches: more patches than pixels");
    q.sx = 2 * q.dx + 2 * dr; q.sy = 2 * q.dy + 2 * dr;
    q.hx = q.dx / 2; q.hy = q.dy / 2;
    if (q.sx > h + 2 * q.hx || q.sy > w + 2 * q.hy) return fail(JUBIK_ERR_ARG, "slice_patches: window larger than the padded array");
    return 0;
}
// window start with lax.dynamic_slice clamping (convolve.py:55-57)
JB_HD int jb_slice_off(int j, int d, int frame, int size) {
    int o = j * d;
    if (o > frame - size) o = frame - size;
    return o < 0 ? 0 : o;
}

// out (n*n, lead, sx, sy): patch k = i*n + j <- padded x rows offx[j].., columns offy[i]..
JB_GLOBAL_S void k_slice_patches(JbSliceGeom q, const double* __restrict__ x, double* __restrict__ out) {
    const size_t per = (size_t)q.sx * q.sy, tot = (size_t)q.n * q.n * q.lead * per;
    for (size_t idx = (size_t)blockIdx.x * jb_nthr() + jb_tid(); idx < tot; idx += (size_t)gridDim.x * jb_nthr()) {
        size_t r = idx;
        const int c = (int)(r % q.sy); r /= q.sy;
        const int a = (int)(r % q.sx); r /= q.sx;
        const int l = (int)(r % q.lead);
        const int k = (int)(r / q.lead), i = k / q.n, j = k - i * q.n;
        const int row = jb_slice_off(j, q.dx, q.h + 2 * q.hx, q.sx) + a - q.hx;
        const int col = jb_slice_off(i, q.dy, q.w + 2 * q.hy, q.sy) + c - q.hy;
        out[idx] = (row >= 0 && row < q.h && col >= 0 && col < q.w) ? x[((size_t)l * q.h + row) * q.w + col] : 0.0;
    }
}
// transpose as a gather: out (lead, h, w) <- sum over the patches covering the pixel, fixed order k ascending
JB_GLOBAL_S void k_slice_patches_T(JbSliceGeom q, const double* __restrict__ patches, double* __restrict__ out) {
    const size_t tot = (size_t)q.lead * q.h * q.w, per = (size_t)q.sx * q.sy;
    for (size_t idx = (size_t)blockIdx.x * jb_nthr() + jb_tid(); idx < tot; idx += (size_t)gridDim.x * jb_nthr()) {
        size_t r = idx;
        const int col = (int)(r % q.w); r /= q.w;
        const int row = (int)(r % q.h);
        const int l = (int)(r / q.h);
        double acc = 0.0;
        for (int i = 0; i < q.n; ++i) {
            const int c = col + q.hy - jb_slice_off(i, q.dy, q.w + 2 * q.hy, q.sy);
            if (c < 0 || c >= q.sy) continue;
            for (int j = 0; j < q.n; ++j) {
                const int a = row + q.hx - jb_slice_off(j, q.dx, q.h + 2 * q.hx, q.sx);
                if (a < 0 || a >= q.sx) continue;
                acc += patches[((size_t)(i * q.n + j) * q.lead + l) * per + (size_t)a * q.sy + c];
            }
        }
        out[idx] = acc;
    }
}
// dvol * (x circularly convolved with y) over the last two axes by direct summation (== Re ifftn(fftn x * fftn y) for
// real inputs); x: (x_lead, n0, n1) with x_lead in {lead, 1} (1 = the reference's broadcasting branch), y, out: (lead, n0, n1)
JB_GLOBAL_S void k_convolve_direct(const double* __restrict__ x, const double* __restrict__ y, int lead, int x_lead, int n0,
                                   int n1, double dvol, double* __restrict__ out) {
    const size_t tot = (size_t)lead * n0 * n1;
    for (size_t idx = (size_t)blockIdx.x * jb_nthr() + jb_tid(); idx < tot; idx += (size_t)gridDim.x * jb_nthr()) {
        size_t r = idx;
        const int j = (int)(r % n1); r /= n1;
        const int i = (int)(r % n0);
        const int l = (int)(r / n0);
        const double* xp = x + (size_t)(x_lead == 1 ? 0 : l) * n0 * n1;
        const double* yp = y + (size_t)l * n0 * n1;
        double acc = 0.0;
        for (int a = 0; a < n0; ++a) {
            int ia = i - a; if (ia < 0) ia += n0;
            for (int b = 0; b < n1; ++b) {
                int jb = j - b; if (jb < 0) jb += n1;
                acc += xp[(size_t)a * n1 + b] * yp[(size_t)ia * n1 + jb];
            }
        }
        out[idx] = dvol * acc;
    }
}

int jubik_slice_patches(const double* x, int lead, int h, int w, int n_patches, int dr, double* out, void* stream_) {
    if (!x || !out) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (device_count() < 1) return fail(JUBIK_ERR_CUDA, "no CUDA device: jubik_b200 has no CPU fallback");
    JbSliceGeom q;
    if (int rc = slice_geom(q, lead, h, w, n_patches, dr)) return rc;
    JB_CK(jb_launch(k_slice_patches, dim3(1184), dim3(256), 0, (jb_stream_t)stream_, q, x, out));
    return JUBIK_OK;
}
int jubik_slice_patches_T(const double* patches, int lead, int h, int w, int n_patches, int dr, double* out, void* stream_) {
    if (!patches || !out) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (device_count() < 1) return fail(JUBIK_ERR_CUDA, "no CUDA device: jubik_b200 has no CPU fallback");
    JbSliceGeom q;
    if (int rc = slice_geom(q, lead, h, w, n_patches, dr)) return rc;
    JB_CK(jb_launch(k_slice_patches_T, dim3(1184), dim3(256), 0, (jb_stream_t)stream_, q, patches, out));
    return JUBIK_OK;
}
int jubik_convolve(const double* x, const double* y, int lead, int x_lead, int n0, int n1, double dvol, double* out,
                   void* stream_) {
    if (!x || !y || !out) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (device_count() < 1) return fail(JUBIK_ERR_CUDA, "no CUDA device: jubik_b200 has no CPU fallback");
    if (lead < 1 || n0 < 1 || n1 < 1 || (x_lead != lead && x_lead != 1)) return fail(JUBIK_ERR_ARG, "convolve: bad shape");
    if ((size_t)n0 * n1 > (size_t)256 * 256)
        return fail(JUBIK_ERR_UNSUPPORTED, "stand-alone convolve uses direct summation (<= 256x256); the FFT path is "
                                           "jubik_plan_create + jubik_psf_apply");
    JB_CK(jb_launch(k_convolve_direct, dim3(1184), dim3(128), 0, (jb_stream_t)stream_, x, y, lead, x_lead, n0, n1, dvol, out));
    return JUBIK_OK;
}

int jubik_poisson_energy_and_grad_host(jubik_plan* p, const double* xh, int S, double* eh, double* gh, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (int rc = check_call(p, xh, eh, S)) return rc;
    const JbGeom& g = p->g;
    const size_t npx = (size_t)g.N * g.N;
    size_t bx = (size_t)p->S_max * g.E * npx * sizeof(double);
    if (!p->xdev) {
        if (alloc(p, (void**)&p->xdev, bx, "host-call staging x") || alloc(p, (void**)&p->gdev, bx, "host-call staging grad") ||
            alloc(p, (void**)&p->edev, p->S_max * sizeof(double), "host-call staging energy"))
            return JUBIK_ERR_ALLOC;
    }
    size_t nx = (size_t)S * g.E * npx * sizeof(double);
#ifndef JB_EMU
    // Energy bins are independent through the whole fused chain, so the call is pipelined per bin: the H2D copy
    // of bin e+1 and the D2H copy of the gradient of bin e-1 overlap the kernels of bin e (three streams).
    const bool chunked = gh && (g.E > 1 || (S > 1 && !p->split_epi)) && (p->split_epi || p->rows5) && p->has_expo && p->has_data && !getenv("JUBIK_B200_HOST_SERIAL");
    if (chunked) {
        if (int rc = ensure_plane_streams(p)) return rc;
        // the copy streams start after everything already queued on the caller's stream
        JB_CK(cudaEventRecord(p->ev_start, st));
        JB_CK(cudaStreamWaitEvent(p->st_in, p->ev_start, 0));
        JB_CK(cudaStreamWaitEvent(p->st_out, p->ev_start, 0));
        if (!p->split_epi) {
            // one (sample, energy bin) plane per chunk: only the first plane's H2D and the last plane's D2H are exposed
            const size_t plane_b = npx * sizeof(double);
            for (int c = 0; c < S * g.E; ++c) {
                JB_CK(cudaMemcpyAsync(p->xdev + (size_t)c * npx, xh + (size_t)c * npx, plane_b, cudaMemcpyHostToDevice, p->st_in));
                JB_CK(cudaEventRecord(p->ev_in[c], p->st_in));
            }
            // Planes alternate between two compute streams: a one-plane launch leaves the persistent kernels with a
            // ragged last wave (1024 rows on 888 resident CTAs), which the other stream's plane fills.
            const bool two = !getenv("JUBIK_B200_HOST_ONE_STREAM") && S * g.E > 1 && p->use_pipe;
            if (two) { if (int rc = ensure_second_workspace(p)) return rc; }
            auto d2h_plane = [&](int c) -> int {
                JB_CK(cudaStreamWaitEvent(p->st_out, p->ev_c[c], 0));
                JB_CK(cudaMemcpyAsync(gh + (size_t)c * npx, p->gdev + (size_t)c * npx, plane_b, cudaMemcpyDeviceToHost, p->st_out));
                return 0;
            };
            if (int rc = run_planes_two_streams(p, p->xdev, S, p->gdev, two, true, st, d2h_plane)) return rc;
            JB_LAUNCHED(p, jb_launch(k_reduce_energy, dim3(S), dim3(256), 256 * sizeof(double), st,
                                     (const double*)p->partial, g.T * g.E * p->nblk_out, p->edev));
        } else {
            const size_t pitch = (size_t)g.E * npx * sizeof(double), width = npx * sizeof(double);
            for (int e = 0; e < g.E; ++e) {
                JB_CK(cudaMemcpy2DAsync(p->xdev + e * npx, pitch, xh + e * npx, pitch, width, S, cudaMemcpyHostToDevice, p->st_in));
                JB_CK(cudaEventRecord(p->ev_in[e], p->st_in));
            }
            for (int e = 0; e < g.E; ++e) {
                JbGeom ge = g;
                ge.e0 = e; ge.Es = 1;
                JB_CK(cudaStreamWaitEvent(st, p->ev_in[e], 0));
                JbEpi ep; memset(&ep, 0, sizeof(ep));
                ep.mode = JB_EPI_PSF; ep.out = p->gbuf;
                if (int rc = run_forward_g(p, ge, p->xdev, S, ep, st)) return rc;
                JB_LAUNCHED(p, jb_launch(k_poisson_epi, dim3(p->nblk_epi, S * g.T), dim3(256), 256 * sizeof(double), st, ge,
                                         p->gbuf, (const double*)p->expo, (const double*)p->data, (int)npx, p->partial));
                if (int rc = run_adjoint_g(p, ge, p->gbuf, nullptr, S, p->gdev, st)) return rc;
                JB_CK(cudaEventRecord(p->ev_c[e], st));
                JB_CK(cudaStreamWaitEvent(p->st_out, p->ev_c[e], 0));
                JB_CK(cudaMemcpy2DAsync(gh + e * npx, pitch, p->gdev + e * npx, pitch, width, S, cudaMemcpyDeviceToHost, p->st_out));
            }
            JB_LAUNCHED(p, jb_launch(k_reduce_energy, dim3(S), dim3(256), 256 * sizeof(double), st,
                                     (const double*)p->partial, g.T * g.E * p->nblk_epi, p->edev));
        }
        if (d2h(eh, p->edev, S * sizeof(double), st)) return fail(JUBIK_ERR_CUDA, "D2H copy of energy failed");
        if (stream_sync(st) || stream_sync(p->st_out)) return fail(JUBIK_ERR_CUDA, "stream synchronise failed");
        return JUBIK_OK;
    }
#endif
    if (h2d(p->xdev, xh, nx, st)) return fail(JUBIK_ERR_CUDA, "H2D copy of x failed");
    if (int rc = jubik_poisson_energy_and_grad(p, p->xdev, S, p->edev, gh ? p->gdev : nullptr, stream_)) return rc;
    if (d2h(eh, p->edev, S * sizeof(double), st)) return fail(JUBIK_ERR_CUDA, "D2H copy of energy failed");
    if (gh && d2h(gh, p->gdev, nx, st)) return fail(JUBIK_ERR_CUDA, "D2H copy of grad failed");
    if (stream_sync(st)) return fail(JUBIK_ERR_CUDA, "stream synchronise failed");
    return JUBIK_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Multi-GPU variants: the fused evaluation followed by ONE ncclAllReduce of a flat buffer on
// the caller's stream.  NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process -- torch's or
// jax's -- or the system one): the library has no link-time NCCL dependency and single-GPU users never load it.
#ifndef JB_EMU
#include <dlfcn.h>
namespace {
struct JbNcclId { char internal[128]; };          // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* jb_nccl_comm_t;
struct JbNccl {
    void* h = nullptr;
    int (*GetUniqueId)(JbNcclId*) = nullptr;
    int (*CommInitRank)(jb_nccl_comm_t*, int, JbNcclId, int) = nullptr;
    int (*CommDestroy)(jb_nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, jb_nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*CommCount)(jb_nccl_comm_t, int*) = nullptr;
    int (*CommUserRank)(jb_nccl_comm_t, int*) = nullptr;
};
JbNccl g_nccl;
int nccl_load() {
    if (g_nccl.h) return 0;
    const char* names[] = {getenv("JUBIK_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(JUBIK_ERR_UNSUPPORTED, std::string("NCCL not found (libnccl.so.2): ") + (dlerror() ? dlerror() : ""));
    JbNccl t; t.h = h;
    t.GetUniqueId = (decltype(t.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    t.CommInitRank = (decltype(t.CommInitRank))dlsym(h, "ncclCommInitRank");
    t.CommDestroy = (decltype(t.CommDestroy))dlsym(h, "ncclCommDestroy");
    t.AllReduce = (decltype(t.AllReduce))dlsym(h, "ncclAllReduce");
    t.GetErrorString = (decltype(t.GetErrorString))dlsym(h, "ncclGetErrorString");
    t.CommCount = (decltype(t.CommCount))dlsym(h, "ncclCommCount");
    t.CommUserRank = (decltype(t.CommUserRank))dlsym(h, "ncclCommUserRank");
    if (!t.GetUniqueId || !t.CommInitRank || !t.CommDestroy || !t.AllReduce || !t.GetErrorString || !t.CommCount || !t.CommUserRank)
        return fail(JUBIK_ERR_UNSUPPORTED, "libnccl lacks a required symbol");
    g_nccl = t;
    return 0;
}
int nccl_fail(const char* what, int rc) {
    return fail(JUBIK_ERR_CUDA, std::string(what) + " -> NCCL error " + std::to_string(rc) + ": " + g_nccl.GetErrorString(rc));
}
}  // namespace
#endif

struct jubik_comm {
    void* nccl = nullptr;   // ncclComm_t
    int rank = 0, n_ranks = 1, owned = 0;
};

extern "C" {

int jubik_comm_unique_id(void* id_out) {
#ifndef JB_EMU
    if (!id_out) return fail(JUBIK_ERR_ARG, "null pointer");
    if (int rc = nccl_load()) return rc;
    if (int rc = g_nccl.GetUniqueId((JbNcclId*)id_out)) return nccl_fail("ncclGetUniqueId", rc);
    return JUBIK_OK;
#else
    (void)id_out;
    return fail(JUBIK_ERR_UNSUPPORTED, "no NCCL in the emulation build");
#endif
}

int jubik_comm_create(jubik_comm** out, const void* id, int rank, int n_ranks) {
#ifndef JB_EMU
    if (!out || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(JUBIK_ERR_ARG, "bad communicator arguments");
    if (int rc = nccl_load()) return rc;
    JbNcclId nid;
    memcpy(&nid, id, sizeof(nid));
    jb_nccl_comm_t c = nullptr;
    if (int rc = g_nccl.CommInitRank(&c, n_ranks, nid, rank)) return nccl_fail("ncclCommInitRank", rc);
    jubik_comm* m = new (std::nothrow) jubik_comm();
    if (!m) return fail(JUBIK_ERR_ALLOC, "out of host memory");
    m->nccl = c; m->rank = rank; m->n_ranks = n_ranks; m->owned = 1;
    *out = m;
    return JUBIK_OK;
#else
    (void)out; (void)id; (void)rank; (void)n_ranks;
    return fail(JUBIK_ERR_UNSUPPORTED, "no NCCL in the emulation build");
#endif
}

int jubik_comm_adopt(jubik_comm** out, void* nccl_comm) {
#ifndef JB_EMU
    if (!out || !nccl_comm) return fail(JUBIK_ERR_ARG, "null pointer");
    if (int rc = nccl_load()) return rc;
    jubik_comm* m = new (std::nothrow) jubik_comm();
    if (!m) return fail(JUBIK_ERR_ALLOC, "out of host memory");
    m->nccl = nccl_comm; m->owned = 0;
    int rc = g_nccl.CommCount(nccl_comm, &m->n_ranks);
    if (!rc) rc = g_nccl.CommUserRank(nccl_comm, &m->rank);
    if (rc) { delete m; return nccl_fail("ncclCommCount/ncclCommUserRank", rc); }
    *out = m;
    return JUBIK_OK;
#else
    (void)out; (void)nccl_comm;
    return fail(JUBIK_ERR_UNSUPPORTED, "no NCCL in the emulation build");
#endif
}

int jubik_comm_destroy(jubik_comm* c) {
    if (!c) return JUBIK_OK;
#ifndef JB_EMU
    if (c->owned && c->nccl && g_nccl.h) g_nccl.CommDestroy(c->nccl);
#endif
    delete c;
    return JUBIK_OK;
}

int jubik_comm_rank(const jubik_comm* c) { return c ? c->rank : 0; }
int jubik_comm_size(const jubik_comm* c) { return c ? c->n_ranks : 1; }

long long jubik_sharded_flat_size(const jubik_plan* p, const jubik_comm* c, int shard_mode, int S) {
    if (!p || S < 1) return -1;
    const long long npx = (long long)p->g.E * p->g.N * p->g.N;
    const int R = c ? c->n_ranks : 1;
    if (shard_mode == JUBIK_SHARD_SAMPLES) return npx + (long long)R * S;
    if (shard_mode == JUBIK_SHARD_MODULES) return (long long)S * npx + S;
    return -1;
}

int jubik_poisson_energy_and_grad_sharded(jubik_plan* p, jubik_comm* c, int shard_mode, const double* x, int S,
                                          double* grad_local, double* flat, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (int rc = check_call(p, x, flat, S)) return rc;
    const JbGeom& g = p->g;
    const size_t npx = (size_t)g.E * g.N * g.N;
    const int R = c ? c->n_ranks : 1, rank = c ? c->rank : 0;
    size_t count = 0;
    if (shard_mode == JUBIK_SHARD_MODULES) {
        // this rank's modules only: partial energies (S) and partial gradients (S,E,N,N) land in flat and are summed
        count = (size_t)S * npx + S;
        if (int rc = jubik_poisson_energy_and_grad(p, x, S, flat + (size_t)S * npx, flat, stream_)) return rc;
        if (grad_local) return fail(JUBIK_ERR_ARG, "module sharding has no per-rank gradient output (pass NULL)");
    } else if (shard_mode == JUBIK_SHARD_SAMPLES) {
        count = npx + (size_t)R * S;
        double* gl = grad_local;
        if (!gl && S > 1) {
            if (!p->gdev) {
                if (int rc = alloc(p, (void**)&p->gdev, (size_t)p->S_max * npx * sizeof(double), "per-sample gradient staging")) return rc;
            }
            gl = p->gdev;
        }
        if (!gl) gl = flat;   // one local sample and no per-sample output wanted: the gradient is written in place
        if (int rc = jubik_poisson_energy_and_grad(p, x, S, flat + npx + (size_t)rank * S, gl, stream_)) return rc;
        JB_LAUNCHED(p, jb_launch(k_fold_samples, dim3(1184), dim3(256), 0, st, (const double*)(gl == flat ? nullptr : gl), S, npx,
                                 flat, rank * S, R * S));
    } else {
        return fail(JUBIK_ERR_ARG, "unknown shard mode");
    }
#ifndef JB_EMU
    if (c && c->nccl) {   // also for a one-rank communicator (NCCL treats it as an in-place copy)
        if (int rc = g_nccl.AllReduce(flat, flat, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->nccl, st))
            return nccl_fail("ncclAllReduce", rc);
    }
#endif
    (void)count;
    return JUBIK_OK;
}

}  // extern "C"
