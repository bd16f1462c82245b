// capi.cu -- the C ABI declared in include/jubik_b200.h: plan creation (PSF spectra precomputed and kept
// resident, exposure/mask/data folded into dense device planes) and the per-call kernel sequences.
#include "ops.cuh"
#include "fft_list.h"
#include "../../include/jubik_b200.h"
#include <algorithm>
#include <functional>
#include <map>
#include <mutex>
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

// Rows per TMA box of a column-pass input tile of `rows` rows: as few boxes as the 256-row limit of a box allows, equal in
// size (the last box overlaps the one before it by at most one row), even (box starts stay 128-byte aligned with 64-byte
// tile rows).  Thread 0 issues these copies between two stages of its own butterflies: 3 boxes instead of the 11 of
// 64-row boxes at P = 664 (JB_TILE_BOX_ROWS = 64 restores the old size).
#ifndef JB_TILE_BOX_ROWS
#define JB_TILE_BOX_ROWS 256
#endif
static int tile_box_rows(int rows) {
    const int nbox = (rows + JB_TILE_BOX_ROWS - 1) / JB_TILE_BOX_ROWS;
    int br = (rows + nbox - 1) / nbox;
    br += br & 1;
    return br > rows ? rows : br;
}

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

#define JB_MAX_SLOTS 4096   // energies (n_ranks x samples) behind the gradient in the host-call flat buffer

struct jubik_plan {
    JbGeom g;
    const JbFftOps* ops = nullptr;
    int S_max = 0, flags = 0;
    int exact = 0, support = 0;
    int has_expo = 0, has_data = 0;
    int point_S = 0;           // samples of the cached linearisation point (0: none)
    int rpc = 1, thr_rows = 128, thr_cols = 128;
    long long launches = 0;
    // device memory
    cd* khat = nullptr;
    cd* tw_nk = nullptr;
    cd* tw_kn = nullptr;
    cd* tw1 = nullptr;         // three-stage column pass: W_L^{m k1}  [k1*M + m]
    cd* tw2 = nullptr;         //                          W_M^{n3 k2} [k2*R3 + n3]
    cd* tw1r = nullptr;        // the same two tables in the factor order of the c2r row kernels (rows5.cuh)
    cd* tw2r = nullptr;
    double* wgt = nullptr;
    double* expo = nullptr;    // (T,E,N,N) effective exposure, 0 == masked
    double* data = nullptr;    // (T,E,N,N)
    double* psums = nullptr;   // (T*E*n*n) PSF sums and half-supports: plan creation and jubik_plan_set_kernel
    int* psupp = nullptr;
    unsigned long long* fpdev = nullptr;   // per-CTA fingerprint pairs (jubik_plan_ensure_kernel)
    unsigned long long fp[2] = {0, 0};     // fingerprint of the kernels the resident spectra were built from
    int fp_layout = -1;
    // bufA: forward  column-pass output (STE, n^2, P or L, PH) -> rows_out;  adjoint row spectra rowsA (STE, n, N, PH)
    // bufB: adjoint  column-pass output (STE, n^2, w, PH) -> rows_adj_out; forward row spectra rowsF (S*E, n, N, PH)
    // (each direction uses one buffer as its de-duplicated input and the other as its per-patch output)
    cd* bufA = nullptr;
    cd* bufB = nullptr;
    double* gbuf = nullptr;    // (S_max,T,E,N,N) cotangent scratch
    double* fac = nullptr;     // (S_max,T,E,N,N) expo^2 / lambda at the linearisation point (JUBIK_PLAN_METRIC)
    double* partial = nullptr; // energy partials
    double* xdev = nullptr;    // staging of the host-buffer entry point (JUBIK_PLAN_HOST_CALL)
    double* gdev = nullptr;
    double* edev = nullptr;
    double* gstage = nullptr;  // per-sample gradients of the sharded call without grad_local (JUBIK_PLAN_SHARDED)
    double* flatdev = nullptr; // device image of `flat` for the host-buffer sharded call (HOST_CALL | SHARDED)
    cd* cols_acc = nullptr;    // T > 1: per-CTA accumulator tiles of the module-summing adjoint column pass (stays in L2)
    cd* cols_acc2 = nullptr;   // ... of the second compute stream (JUBIK_PLAN_HOST_CALL)
    size_t bytes_khat = 0, bytes_ws = 0, bytes_total = 0;
    int nblk_out = 0;
    int use_pipe = 0, n_sm = 1;
    int rows5 = 0;                         // warp-private c2r kernels with the fused Poisson epilogue (rows5.cuh)
    int* gtab_fwd = nullptr;               // rows5 overlap-add gather tables [gq][N]
    int* gtab_adj = nullptr;
    int gq_fwd = 0, gq_adj = 0;
    int qmax = 0;
    int fast_rows = 0;         // rows_fast.cuh kernels (two-pass length, <= JB_MAXQ row-patches per image row, no alias fold)
    JbTileSrc src_fwd, src_adj;  // TMA views of bufB / bufA for the pipelined column pass
    // second spectrum workspace (one sample) of the host-buffer entry point's second compute stream
    cd* bufA2 = nullptr;
    cd* bufB2 = nullptr;
    double* gbuf2 = nullptr;
    int ws2_ready = 0;
    JbTileSrc src_fwd2, src_adj2;
    // per-kernel timing (jubik_poisson_profile): events recorded after every launch while enabled
    int prof_on = 0, prof_n = 0;
#ifndef JB_EMU
    cudaEvent_t prof_ev[48] = {};
    // host-buffer entry point: copy streams and per-plane events; sharded call: the collective's stream
    cudaStream_t st_in = nullptr, st_out = nullptr, st_c2 = nullptr, st_comm = nullptr;
    cudaEvent_t *ev_in = nullptr, *ev_c = nullptr, ev_start = nullptr, *ev_bin = nullptr, ev_comm = nullptr;
    cudaEvent_t *ev_ring = nullptr, ev_free[2] = {nullptr, nullptr};   // host-buffer sharded call: x ring (2 slots x E bins)
    int n_ev = 0;
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
int jubik_version(void) { return 200; }
int jubik_fft_length_supported(int length) { return find_ops(length) != nullptr; }
int jubik_device_count(void) { return device_count(); }

int jubik_plan_destroy(jubik_plan* p) {
    if (!p) return JUBIK_OK;
    void* bufs[] = {p->fpdev, p->khat, p->tw_nk, p->tw_kn, p->tw1, p->tw2, p->tw1r, p->tw2r, p->wgt, p->expo, p->data, p->psums, p->psupp, p->bufA, p->bufB,
                    p->bufA2, p->bufB2, p->gbuf2, p->gbuf, p->fac, p->partial, p->gtab_fwd, p->gtab_adj, p->xdev, p->gdev,
                    p->edev, p->gstage, p->flatdev, p->cols_acc, p->cols_acc2};
    for (void* b : bufs) dev_free(b);
#ifndef JB_EMU
    for (int e = 0; e < p->n_ev; ++e) {
        if (p->ev_in && p->ev_in[e]) cudaEventDestroy(p->ev_in[e]);
        if (p->ev_c && p->ev_c[e]) cudaEventDestroy(p->ev_c[e]);
    }
    delete[] p->ev_in; delete[] p->ev_c;
    if (p->ev_bin) { for (int e = 0; e < p->g.E; ++e) if (p->ev_bin[e]) cudaEventDestroy(p->ev_bin[e]); delete[] p->ev_bin; }
    if (p->ev_ring) { for (int e = 0; e < 2 * p->g.E; ++e) if (p->ev_ring[e]) cudaEventDestroy(p->ev_ring[e]); delete[] p->ev_ring; }
    for (int i = 0; i < 2; ++i) if (p->ev_free[i]) cudaEventDestroy(p->ev_free[i]);
    if (p->ev_start) cudaEventDestroy(p->ev_start);
    if (p->ev_comm) cudaEventDestroy(p->ev_comm);
    if (p->st_in) cudaStreamDestroy(p->st_in);
    if (p->st_out) cudaStreamDestroy(p->st_out);
    if (p->st_c2) cudaStreamDestroy(p->st_c2);
    if (p->st_comm) cudaStreamDestroy(p->st_comm);
    for (int i = 0; i < 48; ++i) if (p->prof_ev[i]) cudaEventDestroy(p->prof_ev[i]);
#endif
    delete p;
    return JUBIK_OK;
}

static int alloc(jubik_plan* p, void** ptr, size_t bytes, const char* what) {
    if (dev_malloc(ptr, bytes)) {
        *ptr = nullptr;
        return fail(JUBIK_ERR_ALLOC, std::string("device allocation failed for ") + what + " (" + std::to_string(bytes) + " bytes)");
    }
    p->bytes_total += bytes;
    return 0;
}

static void choose_launch_shapes(jubik_plan* p) {
    const JbGeom& g = p->g;
    const JbFftOps* o = p->ops;
    // rows kernels: aim for >= 8 sequences per CTA, <= 256 threads; items per pass ~ B*R2 and B*(R1/2+1)
    int rpc = 1;
    while (rpc * g.n < 8 && rpc < g.N) rpc *= 2;
    p->rpc = rpc;
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

// PSF sums and half-supports of every (module, energy, patch) -> p->psums / p->psupp; p->support = the largest
static int psf_stats(jubik_plan* p, const JbKsrc& ks, jb_stream_t st) {
    const JbGeom& g = p->g;
    const size_t nn2 = (size_t)g.n * g.n, TE = (size_t)g.T * g.E;
    JB_CK(jb_launch(k_prep_sum, dim3((unsigned)nn2, (unsigned)TE), dim3(256), 256 * sizeof(double), st, g, ks, p->psums, p->psupp));
    ++p->launches;
    std::vector<int> hs(TE * nn2);
    if (d2h(hs.data(), p->psupp, hs.size() * sizeof(int), st) || stream_sync(st)) return fail(JUBIK_ERR_CUDA, "reading PSF support failed");
    int mx = 0;
    for (int v : hs) mx = v > mx ? v : mx;
    p->support = mx;
    return 0;
}

#define JB_FP_CTAS 296
static size_t kernel_elems(const jubik_plan* p, int layout) {
    const JbGeom& g = p->g;
    const size_t side = layout == JUBIK_KERNEL_CROPPED ? (size_t)g.w : (size_t)g.N;
    return (size_t)g.T * g.n * g.n * g.E * side * side;
}
// fingerprint of a kernel array -> fp[2]; synchronises the stream
static int kernel_fingerprint(jubik_plan* p, const double* kernel, int layout, jb_stream_t st, unsigned long long* fp) {
    if (!p->fpdev && alloc(p, (void**)&p->fpdev, 2 * JB_FP_CTAS * sizeof(unsigned long long), "kernel fingerprint")) return JUBIK_ERR_ALLOC;
    JB_CK(jb_launch(k_fingerprint, dim3(JB_FP_CTAS), dim3(256), 2 * 256 * sizeof(unsigned long long), st, kernel,
                    kernel_elems(p, layout), p->fpdev));
    ++p->launches;
    unsigned long long h[2 * JB_FP_CTAS];
    if (d2h(h, p->fpdev, sizeof(h), st) || stream_sync(st)) return fail(JUBIK_ERR_CUDA, "reading the kernel fingerprint failed");
    fp[0] = fp[1] = 0;
    for (int i = 0; i < JB_FP_CTAS; ++i) { fp[0] += h[2 * i]; fp[1] += h[2 * i + 1]; }
    return 0;
}

// PSF spectra: rows r2c -> columns FFT -> normalise -> resident in the column pass's register order (convolve.py:104-131
// + the kernel FFT of :165, hoisted out of the per-call path).  bufB is the scratch for the row spectra.
static int psf_spectra(jubik_plan* p, const JbKsrc& ks, jb_stream_t st) {
    const JbGeom& g = p->g;
    const JbFftOps* o = p->ops;
    const int prpc = 8;
    int items = prpc * (o->R2 > 1 ? (o->R2 > o->R1 / 2 + 1 ? o->R2 : o->R1 / 2 + 1) : 1);
    int thr = ((items + 31) / 32) * 32; if (thr < 32) thr = 32; if (thr > 256) thr = 256;
    JB_CK(o->prep_rows(g, ks, p->tw_nk, p->bufB, prpc, thr, st));
    ++p->launches;
    JB_CK(o->prep_cols(g, p->bufB, p->psums, p->use_pipe == 2 ? p->tw1 : p->tw_nk, p->tw2, p->khat, p->use_pipe == 2 ? 2 : 0, p->thr_cols, st));
    ++p->launches;
    return 0;
}

#ifndef JB_EMU
// streams / events of the plane-pipelined host call and of the overlapped collective
static int ensure_plane_streams(jubik_plan* p) {
    if (p->st_in) return 0;
    const int nev = p->S_max * p->g.E;
    JB_CK(cudaStreamCreateWithFlags(&p->st_in, cudaStreamNonBlocking));
    JB_CK(cudaStreamCreateWithFlags(&p->st_out, cudaStreamNonBlocking));
    JB_CK(cudaStreamCreateWithFlags(&p->st_c2, cudaStreamNonBlocking));
    p->ev_in = new cudaEvent_t[nev](); p->ev_c = new cudaEvent_t[nev]();
    p->n_ev = nev;
    for (int e = 0; e < nev; ++e) {
        JB_CK(cudaEventCreateWithFlags(&p->ev_in[e], cudaEventDisableTiming));
        JB_CK(cudaEventCreateWithFlags(&p->ev_c[e], cudaEventDisableTiming));
    }
    if (!p->ev_start) JB_CK(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
    return 0;
}
static int ensure_comm_stream(jubik_plan* p) {
    if (p->st_comm) return 0;
    JB_CK(cudaStreamCreateWithFlags(&p->st_comm, cudaStreamNonBlocking));
    p->ev_bin = new cudaEvent_t[p->g.E]();
    for (int e = 0; e < p->g.E; ++e) JB_CK(cudaEventCreateWithFlags(&p->ev_bin[e], cudaEventDisableTiming));
    p->ev_ring = new cudaEvent_t[2 * p->g.E]();
    for (int e = 0; e < 2 * p->g.E; ++e) JB_CK(cudaEventCreateWithFlags(&p->ev_ring[e], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) JB_CK(cudaEventCreateWithFlags(&p->ev_free[i], cudaEventDisableTiming));
    JB_CK(cudaEventCreateWithFlags(&p->ev_comm, cudaEventDisableTiming));
    if (!p->ev_start) JB_CK(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
    return 0;
}
// The two compute streams' planes are in different phases at the same time, and the forward / adjoint layouts of one
// spectrum buffer alias across energy bins: the second stream gets its own pair of buffers (one sample's worth).
static int ensure_second_workspace(jubik_plan* p) {
    if (p->ws2_ready) return 0;
    const JbGeom& g = p->g;
    const size_t TE1 = (size_t)g.T * g.E, nn2 = (size_t)g.n * g.n;
    const int brB = std::min(tile_box_rows(g.w), g.N), brA = std::min(tile_box_rows(g.P), g.N);
    const size_t orows = g.fold ? g.L : g.P;
    std::string err;
    if ((!p->bufA2 && alloc(p, (void**)&p->bufA2, TE1 * nn2 * orows * g.PH * sizeof(cd), "second spectrum workspace A")) ||
        (!p->bufB2 && alloc(p, (void**)&p->bufB2, TE1 * nn2 * (size_t)g.w * g.PH * sizeof(cd), "second spectrum workspace B")) ||
        (!p->gbuf2 && alloc(p, (void**)&p->gbuf2, TE1 * (size_t)g.N * g.N * sizeof(double), "second cotangent workspace")) ||
        (p->cols_acc && !p->cols_acc2 &&
         alloc(p, (void**)&p->cols_acc2, (size_t)p->n_sm * p->ops->cols_cps * g.L * p->ops->CB * sizeof(cd), "second adjoint column accumulators")))
        return JUBIK_ERR_ALLOC;
    if (p->use_pipe && (make_tile_src(&p->src_fwd2, p->bufB2, TE1 * (size_t)g.n, g.N, g.PH, p->ops->CB, brB, &err) ||
                        make_tile_src(&p->src_adj2, p->bufA2, TE1 * (size_t)g.n, g.N, g.PH, p->ops->CB, brA, &err)))
        return fail(JUBIK_ERR_CUDA, err);
    p->ws2_ready = 1;
    return 0;
}
#endif

// Transform length.  P compiled in: the period is exactly P (reference semantics for any PSF).  Otherwise a larger
// compiled length: the padded linear convolution when the PSF half-support s <= m, else (L >= w + 2s) the linear
// convolution folded modulo P (alias fold, kernels.cuh).
static int choose_length(const JbGeom& g, int support, int forced, int* L_out, int* fold_out, int* ext_out) {
    const int need_fold = g.w + 2 * support;
    int L = 0, fold = 0;
    if (forced > 0) {
        if (forced < g.P) return fail(JUBIK_ERR_ARG, "fft_length < P = 2*int(N/n) + 2*margin");
        if (!find_ops(forced)) return fail(JUBIK_ERR_UNSUPPORTED, "fft_length " + std::to_string(forced) + " is not compiled in");
        L = forced;
        if (L != g.P && support > g.m) {
            if (L < need_fold)
                return fail(JUBIK_ERR_ARG, "fft_length " + std::to_string(L) + " is too short for the alias fold: the PSF half-support (" +
                            std::to_string(support) + ") exceeds the margin, which needs L >= w + 2*support = " + std::to_string(need_fold));
            fold = 1;
        }
    } else if (find_ops(g.P)) {
        L = g.P;
    } else if (support <= g.m) {
        for (int c : k_lengths) if (c >= g.P && (L == 0 || c < L)) L = c;
        if (!L) return fail(JUBIK_ERR_UNSUPPORTED, "no compiled FFT length >= P = " + std::to_string(g.P));
    } else {
        for (int c : k_lengths) if (c >= need_fold && c >= g.P && (L == 0 || c < L)) L = c;
        if (!L)
            return fail(JUBIK_ERR_UNSUPPORTED, "P = " + std::to_string(g.P) + " is not a compiled FFT length, the PSF half-support (" +
                        std::to_string(support) + ") exceeds the margin (" + std::to_string(g.m) + "), and no compiled length >= w + 2*support = " +
                        std::to_string(need_fold) + " exists for the period-P alias fold");
        fold = 1;
    }
    *L_out = L; *fold_out = fold; *ext_out = fold ? support - g.m : 0;
    return 0;
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
    // every early return below goes through this guard: the plan owns all its temporaries, so nothing leaks
    struct Guard { jubik_plan* p; ~Guard() { if (p) jubik_plan_destroy(p); } } guard{p};
    JbGeom& g = p->g;
    memset(&g, 0, sizeof(g));
    g.N = N; g.n = n; g.m = m; g.dx = dx; g.w = 2 * dx; g.hp = dx / 2; g.P = g.w + 2 * m;
    g.S = d->max_samples; g.T = d->n_modules; g.E = d->n_energy;
    g.e0 = 0; g.Es = g.E;
    p->S_max = d->max_samples;
    p->flags = d->flags;
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
#define JB_TRY(expr) do { rc = (expr); if (rc) return rc; } while (0)

    // ---- PSF sums + support (decides the transform length)
    JbKsrc ks;
    ks.ptr = d->kernel; ks.cropped = (d->kernel_layout == JUBIK_KERNEL_CROPPED); ks.T = g.T; ks.E = g.E;
    JB_TRY(alloc(p, (void**)&p->psums, TE * nn2 * sizeof(double), "psf sums"));
    JB_TRY(alloc(p, (void**)&p->psupp, TE * nn2 * sizeof(int), "psf support"));
    JB_TRY(psf_stats(p, ks, st));
    int L = 0;
    JB_TRY(choose_length(g, p->support, d->fft_length, &L, &g.fold, &g.ext));
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
            h2d(p->wgt, wv.data(), g.w * sizeof(double), st) || stream_sync(st))
            return fail(JUBIK_ERR_CUDA, "uploading twiddles failed");
    }
    // ---- workspace
    const size_t nstrips = (size_t)jb_cdiv(g.PH, o->CB);
    const size_t orows = g.fold ? (size_t)L : (size_t)g.P;    // rows per patch of the forward column-pass output
    p->bytes_khat = TE * nn2 * nstrips * (size_t)L * o->CB * sizeof(cd);
    size_t bA = STE * nn2 * orows * g.PH * sizeof(cd);
    size_t bB = STE * nn2 * (size_t)g.w * g.PH * sizeof(cd);
    size_t bG = STE * (size_t)N * N * sizeof(double);
    size_t bP = STE * (size_t)N * sizeof(double);   // one energy partial per image row at most
    p->bytes_ws = bA + bB + bG + bP;
    JB_TRY(alloc(p, (void**)&p->khat, p->bytes_khat, "PSF spectra"));
    JB_TRY(alloc(p, (void**)&p->bufA, bA, "spectrum workspace A"));
    JB_TRY(alloc(p, (void**)&p->bufB, bB, "spectrum workspace B"));
    JB_TRY(alloc(p, (void**)&p->gbuf, bG, "cotangent workspace"));
    JB_TRY(alloc(p, (void**)&p->partial, bP, "energy partials"));
    // ---- which kernels: the fast path (TMA column pipeline + warp-private c2r rows + prefetching r2c rows) needs a
    // three-stage length, at most JB_MAXQ row-patches covering any image row and no alias fold; everything else runs
    // the generic kernels of kernels.cuh (same maths, any geometry).  No environment switches: the plan is a pure
    // function of its descriptor.
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
        p->qmax = mx;
        p->fast_rows = (o->R2 > 1) && mx <= JB_MAXQ && !g.fold;
    }
    p->use_pipe = g.fold ? 0 : o->pipe;
    if (p->use_pipe == 2) {
        // two table pairs: the column pass's factor order (tw1/tw2) and the c2r row kernels' (tw1r/tw2r, fft_list.h)
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int which = 0; which < 2; ++which) {
            const int Q1 = which ? o->rq1 : o->q1, Q2 = which ? o->rq2 : o->q2, Q3 = which ? o->rq3 : o->q3, M = Q2 * Q3;
            std::vector<cd> a((size_t)L), b((size_t)M);
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
            cd** d1 = which ? &p->tw1r : &p->tw1;
            cd** d2 = which ? &p->tw2r : &p->tw2;
            JB_TRY(alloc(p, (void**)d1, L * sizeof(cd), "twiddles"));
            JB_TRY(alloc(p, (void**)d2, M * sizeof(cd), "twiddles"));
            if (h2d(*d1, a.data(), L * sizeof(cd), st) || h2d(*d2, b.data(), M * sizeof(cd), st) || stream_sync(st))
                return fail(JUBIK_ERR_CUDA, "uploading three-stage twiddles failed");
        }
    }
    if (p->use_pipe == 2 && p->qmax <= JB_MAXQ)
        p->rows5 = o->rows5_smem(g) <= (size_t)227 * 1024 - 1024;
    if (p->rows5) {
        std::vector<int> tf((size_t)JB_MAXQ * N), ta((size_t)JB_MAXQ * N);
        p->gq_fwd = o->rows5_gather(g, 0, JB_MAXQ, tf.data());
        p->gq_adj = o->rows5_gather(g, 1, JB_MAXQ, ta.data());
        if (p->gq_fwd < 1 || p->gq_adj < 1) {
            p->rows5 = 0;
        } else {
            JB_TRY(alloc(p, (void**)&p->gtab_fwd, tf.size() * sizeof(int), "gather table"));
            JB_TRY(alloc(p, (void**)&p->gtab_adj, ta.size() * sizeof(int), "gather table"));
            if (h2d(p->gtab_fwd, tf.data(), tf.size() * sizeof(int), st) ||
                h2d(p->gtab_adj, ta.data(), ta.size() * sizeof(int), st) || stream_sync(st))
                return fail(JUBIK_ERR_CUDA, "uploading gather tables failed");
        }
    }
    p->nblk_out = p->rows5 ? N : jb_cdiv(N, p->rpc);
    if (!p->rows5 && o->rows_generic_smem(g, p->rpc, p->thr_rows) > (size_t)227 * 1024)
        return fail(JUBIK_ERR_UNSUPPORTED, "this geometry needs the generic row kernels (alias fold, or a transform length without the "
                                           "three-stage pipeline) and their shared-memory tile exceeds 227 KB: use fewer patches per axis");
    if (p->use_pipe) {
        std::string err;
        int brB = tile_box_rows(g.w), brA = tile_box_rows(g.P);
        if (brB > N) brB = N;
        if (brA > N) brA = N;
        // the row-spectrum buffers (slabs = planes*n, N rows, PH) live in the first part of bufB (forward) / bufA (adjoint)
        if (make_tile_src(&p->src_fwd, p->bufB, STE * (size_t)n, N, g.PH, o->CB, brB, &err) ||
            make_tile_src(&p->src_adj, p->bufA, STE * (size_t)n, N, g.PH, o->CB, brA, &err))
            return fail(JUBIK_ERR_CUDA, err);
    }
#ifndef JB_EMU
    {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        p->n_sm = sms > 0 ? sms : 148;
    }
#else
    p->n_sm = 2;
#endif
    if (p->use_pipe == 2 && p->rows5 && o->cols_mt && g.T > 1)
        JB_TRY(alloc(p, (void**)&p->cols_acc, (size_t)p->n_sm * o->cols_cps * L * o->CB * sizeof(cd), "adjoint column accumulators"));
    JB_TRY(psf_spectra(p, ks, st));
    if (stream_sync(st)) return fail(JUBIK_ERR_CUDA, "PSF spectrum kernels failed");
    JB_TRY(kernel_fingerprint(p, d->kernel, d->kernel_layout, st, p->fp));
    p->fp_layout = d->kernel_layout;
    // ---- optional staging, all of it allocated HERE (no allocation after plan creation)
    if (p->flags & JUBIK_PLAN_METRIC) JB_TRY(alloc(p, (void**)&p->fac, bG, "metric linearisation point"));
    if (p->flags & JUBIK_PLAN_SHARDED)
        JB_TRY(alloc(p, (void**)&p->gstage, (size_t)p->S_max * g.E * N * N * sizeof(double), "per-sample gradient staging"));
    if (p->flags & JUBIK_PLAN_HOST_CALL) {
        const size_t bx = (size_t)p->S_max * g.E * N * N * sizeof(double);
        JB_TRY(alloc(p, (void**)&p->xdev, bx, "host-call staging x"));
        JB_TRY(alloc(p, (void**)&p->gdev, bx, "host-call staging grad"));
        JB_TRY(alloc(p, (void**)&p->edev, p->S_max * sizeof(double), "host-call staging energy"));
#ifndef JB_EMU
        JB_TRY(ensure_plane_streams(p));
        if (p->S_max * g.E > 1) JB_TRY(ensure_second_workspace(p));
#endif
    }
    if ((p->flags & JUBIK_PLAN_SHARDED) && (p->flags & JUBIK_PLAN_HOST_CALL))
        JB_TRY(alloc(p, (void**)&p->flatdev, ((size_t)g.E * N * N + JB_MAX_SLOTS) * sizeof(double), "host-call flat buffer"));
#ifndef JB_EMU
    if (p->flags & JUBIK_PLAN_SHARDED) JB_TRY(ensure_comm_stream(p));
#endif
    // ---- exposure / mask / data
    if (d->exposure) {
        size_t tot = TE * (size_t)N * N;
        JB_TRY(alloc(p, (void**)&p->expo, tot * sizeof(double), "exposure"));
        cudaError_t e = jb_launch(k_prep_expo, dim3(592), dim3(256), 0, st, d->exposure, d->eff_area, d->exposure_cut,
                                  (int)(d->has_exposure_cut != 0), d->pixel_area, (size_t)N * N, tot, p->expo);
        if (e != cudaSuccess) return fail(JUBIK_ERR_CUDA, "k_prep_expo -> " + cuda_msg(e));
        ++p->launches;
        p->has_expo = 1;
    }
    if (d->data) JB_TRY(jubik_plan_set_data(p, d->data, d->data_dtype, stream_));
    if (stream_sync(st)) return fail(JUBIK_ERR_CUDA, "plan creation: stream sync failed");
    guard.p = nullptr;
    *out = p;
    return JUBIK_OK;
}

int jubik_plan_set_kernel(jubik_plan* p, const double* kernel, int kernel_layout, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (!p || !kernel) return fail(JUBIK_ERR_ARG, "null plan/kernel");
    JbGeom& g = p->g;
    JbKsrc ks;
    ks.ptr = kernel; ks.cropped = (kernel_layout == JUBIK_KERNEL_CROPPED); ks.T = g.T; ks.E = g.E;
    const int old_support = p->support;
    if (int rc = psf_stats(p, ks, st)) return rc;
    if (!p->exact) {
        // a padded transform stays exact only while the new PSF fits the plan's choice
        const bool ok = g.fold ? (g.L >= g.w + 2 * p->support) : (p->support <= g.m);
        if (!ok) {
            const int s_new = p->support;
            p->support = old_support;
            return fail(JUBIK_ERR_UNSUPPORTED, "the new PSF's half-support (" + std::to_string(s_new) + ") does not fit this plan's "
                        "transform length " + std::to_string(g.L) + " (P = " + std::to_string(g.P) + "): create a new plan");
        }
        if (g.fold) g.ext = p->support > g.m ? p->support - g.m : 0;
    }
    p->point_S = 0;    // a cached linearisation point belongs to the old response
    p->fp_layout = -1;
    if (int rc = psf_spectra(p, ks, st)) return rc;
    if (int rc = kernel_fingerprint(p, kernel, kernel_layout, st, p->fp)) return rc;
    p->fp_layout = kernel_layout;
    return JUBIK_OK;
}

int jubik_plan_ensure_kernel(jubik_plan* p, const double* kernel, int kernel_layout, int* replanned, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (!p || !kernel) return fail(JUBIK_ERR_ARG, "null plan/kernel");
    if (replanned) *replanned = 0;
    unsigned long long fp[2];
    if (int rc = kernel_fingerprint(p, kernel, kernel_layout, st, fp)) return rc;
    if (kernel_layout == p->fp_layout && fp[0] == p->fp[0] && fp[1] == p->fp[1]) return JUBIK_OK;
    if (replanned) *replanned = 1;
    return jubik_plan_set_kernel(p, kernel, kernel_layout, stream_);
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
    i->alias_fold = g.fold; i->alias_ext = g.ext;
    i->fast_path = (p->use_pipe == 2 && p->rows5 && p->fast_rows) ? 1 : 0;
    i->flags = p->flags;
    i->bytes_khat = p->bytes_khat; i->bytes_workspace = p->bytes_ws; i->bytes_total = p->bytes_total;
    return JUBIK_OK;
}

long long jubik_plan_launch_count(const jubik_plan* p) { return p ? p->launches : 0; }

int jubik_stream_synchronize(void* stream_) {
    if (stream_sync((jb_stream_t)stream_)) return fail(JUBIK_ERR_CUDA, "stream synchronise failed");
    return JUBIK_OK;
}

static int check_call(jubik_plan* p, const void* a, const void* b, int S) {
    if (!p) return fail(JUBIK_ERR_ARG, "null plan");
    if (!a || !b) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (S < 1 || S > p->S_max) return fail(JUBIK_ERR_ARG, "n_samples must be in [1, max_samples]");
    return 0;
}

// forward chain with the chosen epilogue; `g` may be a copy of p->g restricted to an energy-bin chunk (e0, Es);
// exp_in: x holds log_sky and the r2c row kernel forms exp(log_sky) on the fly
static int run_forward_g(jubik_plan* p, const JbGeom& g, const double* x, int S, const JbEpi& ep, int exp_in, jb_stream_t st) {
    const JbFftOps* o = p->ops;
    JB_LAUNCHED(p, o->rows_fwd(g, x, p->wgt, p->tw_nk, p->bufB, S * g.Es, p->rpc, p->thr_rows, p->fast_rows, exp_in, st));
    if (p->use_pipe) {
        JB_LAUNCHED(p, o->cols_pipe(g, false, p->src_fwd, p->bufA, p->khat, p->tw1, p->tw2, p->wgt, S * g.T * g.Es, p->n_sm,
                                    (int)(o->cols_mt && g.T > 1), nullptr, st));
    } else {
        JB_LAUNCHED(p, o->cols(g, false, p->bufB, p->bufA, p->khat, p->tw_nk, p->wgt, S * g.T * g.Es, p->thr_cols, st));
    }
    if (p->rows5) {
        JB_LAUNCHED(p, o->rows5_fwd(g, p->bufA, p->tw1r, p->tw2r, p->gtab_fwd, p->gq_fwd, ep, S * g.T * g.Es, p->n_sm, st));
    } else {
        JB_LAUNCHED(p, o->rows_out(g, p->bufA, p->tw_kn, ep, S * g.T * g.Es, p->rpc, p->thr_rows, p->fast_rows, st));
    }
    return 0;
}
static int run_forward(jubik_plan* p, const double* x, int S, const JbEpi& ep, jb_stream_t st) {
    return run_forward_g(p, p->g, x, S, ep, 0, st);
}

// adjoint chain; expo != nullptr multiplies the incoming cotangent by the effective exposure; logsky != nullptr
// multiplies the outgoing gradient by exp(log_sky) (VJP of the sky-model tail)
static int run_adjoint_g(jubik_plan* p, const JbGeom& g, const double* gin, const double* expo, int S, double* out,
                         const double* logsky, jb_stream_t st) {
    const JbFftOps* o = p->ops;
    JB_LAUNCHED(p, o->rows_adj_in(g, gin, expo, p->tw_nk, p->bufA, S * g.T * g.Es, p->rpc, p->thr_rows, p->fast_rows, st));
    // several modules on the fast path: the column pass sums them in the spectral domain (one inverse transform per
    // patch) and the c2r row pass reads the module-summed spectra as if there were one module
    const int mt = (p->use_pipe == 2 && p->rows5 && o->cols_mt && g.T > 1 && p->cols_acc) ? 1 : 0;
    if (p->use_pipe) {
        JB_LAUNCHED(p, o->cols_pipe(g, true, p->src_adj, p->bufB, p->khat, p->tw1, p->tw2, p->wgt, S * g.T * g.Es, p->n_sm, mt,
                                    p->cols_acc, st));
    } else {
        JB_LAUNCHED(p, o->cols(g, true, p->bufA, p->bufB, p->khat, p->tw_nk, p->wgt, S * g.T * g.Es, p->thr_cols, st));
    }
    if (p->rows5) {
        JbGeom gr = g;
        if (mt) gr.T = 1;
        JB_LAUNCHED(p, o->rows5_adj(gr, p->bufB, p->wgt, p->tw1r, p->tw2r, p->gtab_adj, p->gq_adj, out, logsky, S * g.Es, p->n_sm, st));
    } else {
        JB_LAUNCHED(p, o->rows_adj_out(g, p->bufB, p->wgt, p->tw_kn, out, logsky, S * g.Es, p->rpc, p->thr_rows, p->fast_rows, st));
    }
    return 0;
}
static int run_adjoint(jubik_plan* p, const double* gin, const double* expo, int S, double* out, jb_stream_t st) {
    return run_adjoint_g(p, p->g, gin, expo, S, out, nullptr, st);
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

static int reduce_energy(jubik_plan* p, int S, double* energy, jb_stream_t st) {
    JB_LAUNCHED(p, jb_launch(k_reduce_energy, dim3(S), dim3(256), 256 * sizeof(double), st, (const double*)p->partial,
                             p->g.T * p->g.E * p->nblk_out, energy));
    return 0;
}

// fused energy + gradient of sum(lambda) - d log(lambda); logsky: x holds log_sky (exp on load, * exp(log_sky) on store);
// gs: geometry with the sky-space strides of a padded log-sky (or p->g)
static int energy_and_grad_geom(jubik_plan* p, const JbGeom& gs, const double* x, int S, double* energy, double* grad, bool logsky,
                                jb_stream_t st) {
    if (!p->has_expo || !p->has_data) return fail(JUBIK_ERR_ARG, "plan needs exposures and data for the Poisson energy");
    JbEpi ep; memset(&ep, 0, sizeof(ep));
    ep.mode = JB_EPI_POISSON; ep.out = p->gbuf; ep.expo = p->expo; ep.data = p->data; ep.partial = p->partial;
    if (int rc = run_forward_g(p, gs, x, S, ep, logsky ? 1 : 0, st)) return rc;
    if (int rc = reduce_energy(p, S, energy, st)) return rc;
    if (grad) return run_adjoint_g(p, gs, p->gbuf, nullptr, S, grad, logsky ? x : nullptr, st);
    return JUBIK_OK;
}
static int energy_and_grad_impl(jubik_plan* p, const double* x, int S, double* energy, double* grad, bool logsky, jb_stream_t st) {
    return energy_and_grad_geom(p, p->g, x, S, energy, grad, logsky, st);
}

int jubik_poisson_energy_and_grad(jubik_plan* p, const double* x, int S, double* energy, double* grad, void* stream_) {
    if (int rc = check_call(p, x, energy, S)) return rc;
    return energy_and_grad_impl(p, x, S, energy, grad, false, (jb_stream_t)stream_);
}

int jubik_poisson_energy_and_grad_logsky(jubik_plan* p, const double* log_sky, int S, double* energy, double* grad,
                                         void* stream_) {
    if (int rc = check_call(p, log_sky, energy, S)) return rc;
    return energy_and_grad_impl(p, log_sky, S, energy, grad, true, (jb_stream_t)stream_);
}

int jubik_poisson_energy_and_grad_logsky_padded(jubik_plan* p, const double* log_sky, int S, int e_pad, int n_pad, double* energy,
                                                double* grad, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (int rc = check_call(p, log_sky, energy, S)) return rc;
    const JbGeom& g = p->g;
    if (e_pad < g.E || n_pad < g.N) return fail(JUBIK_ERR_ARG, "the padded log-sky must be at least (E, N, N)");
    if (n_pad & 1) return fail(JUBIK_ERR_UNSUPPORTED, "the padded spatial size must be even (16-byte aligned rows)");
    JbGeom gs = g;
    gs.sky_row = n_pad;
    gs.sky_pln = (long long)n_pad * n_pad;
    gs.sky_smp = (long long)e_pad * n_pad * n_pad;
    if (grad) {   // zero outside the [:E, :N, :N] crop (the kernels write the crop only)
#ifndef JB_EMU
        JB_CK(cudaMemsetAsync(grad, 0, (size_t)S * e_pad * n_pad * n_pad * sizeof(double), st));
#else
        memset(grad, 0, (size_t)S * e_pad * n_pad * n_pad * sizeof(double));
#endif
    }
    return energy_and_grad_geom(p, gs, log_sky, S, energy, grad, true, st);
}

#ifndef JB_EMU
// kernel arguments are captured at launch, so the second stream's launches simply see the plan with its workspace
// pointers exchanged while they are enqueued
static void swap_workspace(jubik_plan* p) {
    std::swap(p->bufA, p->bufA2); std::swap(p->bufB, p->bufB2);
    std::swap(p->src_fwd, p->src_fwd2); std::swap(p->src_adj, p->src_adj2);
    std::swap(p->gbuf, p->gbuf2);
    if (p->cols_acc2) std::swap(p->cols_acc, p->cols_acc2);
}
// Fused energy + gradient as (sample, energy bin) planes alternating between the caller's stream and a second one.
// wait_in: wait for the per-plane H2D events; after_plane(c): hook after a plane's kernels were enqueued (D2H).
static int run_planes_two_streams(jubik_plan* p, const double* x, int S, double* grad, bool two, bool wait_in,
                                  jb_stream_t st, const std::function<int(int)>& after_plane) {
    const JbGeom& g = p->g;
    const size_t npx = (size_t)g.N * g.N;
    if (two) JB_CK(cudaStreamWaitEvent(p->st_c2, p->ev_start, 0));
    // One (sample, energy bin) plane per chunk.  With two streams every stream owns its spectrum buffers AND its
    // cotangent scratch, so chunks on different streams never touch the same memory; chunks on one stream are ordered
    // by the stream.
    const int nplanes = S * g.E;
    int last_c2 = -1;
    for (int c = 0; c < nplanes; ++c) {
        const int si = c / g.E, e = c - si * g.E;
        jb_stream_t cs = (two && (c & 1)) ? (jb_stream_t)p->st_c2 : st;
        if (cs != st) last_c2 = c;
        JbGeom ge = g;
        ge.e0 = e; ge.Es = 1;
        if (wait_in) JB_CK(cudaStreamWaitEvent(cs, p->ev_in[c], 0));
        JbEpi ep; memset(&ep, 0, sizeof(ep));
        if (cs != st) swap_workspace(p);
        ep.mode = JB_EPI_POISSON; ep.out = p->gbuf; ep.expo = p->expo; ep.data = p->data;
        ep.partial = p->partial + (size_t)si * g.T * g.E * p->nblk_out;
        // the chunk runs as "sample 0" of a one-sample batch on this sample's slices of x / grad / partials
        int rc = run_forward_g(p, ge, x + (size_t)si * g.E * npx, 1, ep, 0, cs);
        if (!rc) rc = run_adjoint_g(p, ge, p->gbuf, nullptr, 1, grad + (size_t)si * g.E * npx, nullptr, cs);
        if (cs != st) swap_workspace(p);
        if (rc) return rc;
        JB_CK(cudaEventRecord(p->ev_c[c], cs));
        if (int rc2 = after_plane(c)) return rc2;
    }
    if (last_c2 >= 0) JB_CK(cudaStreamWaitEvent(st, p->ev_c[last_c2], 0));   // join before the energy reduction
    return 0;
}
#endif

// Same work as jubik_poisson_energy_and_grad with a CUDA event after every kernel; ms_out[0..*n_out) in launch
// order: rows_fwd, cols_fwd, rows_out, reduce, rows_adj_in, cols_adj, rows_adj_out.  Synchronises.
int jubik_poisson_profile(jubik_plan* p, const double* x, int S, double* energy, double* grad, float* ms_out, int* n_out, void* stream_) {
#ifndef JB_EMU
    jb_stream_t st = (jb_stream_t)stream_;
    if (!p || !ms_out || !grad || !n_out) return fail(JUBIK_ERR_ARG, "null pointer");
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

// ---- Fisher metric of the Poisson likelihood pulled back through R, with a cached linearisation point.
// NIFTy's CG holds the position fixed for 60-70 metric applies per sample (eROSITA_demo.yaml:161): set_point runs ONE
// forward pass that stores fac = expo^2 / lambda(x0) (0 where masked); every apply is then one forward (epilogue
// out = fac * conv(v)) and one adjoint.
int jubik_poisson_set_point(jubik_plan* p, const double* x, int S, void* stream_) {
    if (int rc = check_call(p, x, x, S)) return rc;
    if (!p->has_expo) return fail(JUBIK_ERR_ARG, "plan was created without exposures");
    if (!p->fac) return fail(JUBIK_ERR_ARG, "plan was created without JUBIK_PLAN_METRIC (no linearisation-point buffer)");
    JbEpi ep; memset(&ep, 0, sizeof(ep));
    ep.mode = JB_EPI_MFAC; ep.out = p->fac; ep.expo = p->expo;
    if (int rc = run_forward(p, x, S, ep, (jb_stream_t)stream_)) return rc;
    p->point_S = S;
    return JUBIK_OK;
}

int jubik_poisson_metric_apply_at_point(jubik_plan* p, const double* v, int S, double* out, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (int rc = check_call(p, v, out, S)) return rc;
    if (!p->fac || p->point_S < 1) return fail(JUBIK_ERR_ARG, "no linearisation point: call jubik_poisson_set_point first");
    if (S != p->point_S) return fail(JUBIK_ERR_ARG, "n_samples differs from the linearisation point's");
    JbEpi ep; memset(&ep, 0, sizeof(ep));
    ep.mode = JB_EPI_METRIC; ep.out = p->gbuf; ep.fac = p->fac;
    if (int rc = run_forward(p, v, S, ep, st)) return rc;
    return run_adjoint(p, p->gbuf, nullptr, S, out, st);
}

int jubik_poisson_left_sqrt_metric_at_point(jubik_plan* p, const double* t, int S, double* out, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (int rc = check_call(p, t, out, S)) return rc;
    if (!p->fac || p->point_S < 1) return fail(JUBIK_ERR_ARG, "no linearisation point: call jubik_poisson_set_point first");
    if (S != p->point_S) return fail(JUBIK_ERR_ARG, "n_samples differs from the linearisation point's");
    const size_t cnt = (size_t)S * p->g.T * p->g.E * p->g.N * p->g.N;
    JB_LAUNCHED(p, jb_launch(k_lsm_epi, dim3(1184), dim3(256), 0, st, t, (const double*)p->fac, p->gbuf, cnt));
    return run_adjoint(p, p->gbuf, nullptr, S, out, st);
}

int jubik_poisson_metric_apply(jubik_plan* p, const double* x, const double* v, int S, double* out, void* stream_) {
    if (!v) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (int rc = jubik_poisson_set_point(p, x, S, stream_)) return rc;
    return jubik_poisson_metric_apply_at_point(p, v, S, out, stream_);
}

int jubik_poisson_left_sqrt_metric(jubik_plan* p, const double* x, const double* t, int S, double* out, void* stream_) {
    if (!t) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (int rc = jubik_poisson_set_point(p, x, S, stream_)) return rc;
    return jubik_poisson_left_sqrt_metric_at_point(p, t, S, out, stream_);
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
// twiddle tables of a compiled length for the stand-alone convolution: device-resident, built once per (device, length)
static int conv_twiddles(int L, const JbFftOps* o, const cd** tw_nk, const cd** tw_kn, jb_stream_t st) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, std::pair<cd*, cd*>> cache;
    int dev = 0;
#ifndef JB_EMU
    cudaGetDevice(&dev);
#endif
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, L});
    if (it == cache.end()) {
        const int R1 = o->R1, R2 = o->R2;
        std::vector<cd> a((size_t)L), b((size_t)L);
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int k1 = 0; k1 < R1; ++k1)
            for (int n2 = 0; n2 < R2; ++n2) {
                long double ang = -two_pi * (long double)(((long long)k1 * n2) % L) / (long double)L;
                cd v = cd_make((double)cosl(ang), (double)sinl(ang));
                a[(size_t)k1 * R2 + n2] = v;
                b[(size_t)n2 * R1 + k1] = v;
            }
        cd *da = nullptr, *db = nullptr;
        if (dev_malloc((void**)&da, L * sizeof(cd)) || dev_malloc((void**)&db, L * sizeof(cd)))
            return fail(JUBIK_ERR_ALLOC, "device allocation failed for convolution twiddles");
        if (h2d(da, a.data(), L * sizeof(cd), st) || h2d(db, b.data(), L * sizeof(cd), st) || stream_sync(st))
            return fail(JUBIK_ERR_CUDA, "uploading convolution twiddles failed");
        it = cache.emplace(std::make_pair(dev, L), std::make_pair(da, db)).first;
    }
    *tw_nk = it->second.first; *tw_kn = it->second.second;
    return 0;
}

int jubik_convolve(const double* x, const double* y, int lead, int x_lead, int n0, int n1, double dvol, double* out,
                   void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (!x || !y || !out) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (device_count() < 1) return fail(JUBIK_ERR_CUDA, "no CUDA device: jubik_b200 has no CPU fallback");
    if (lead < 1 || n0 < 1 || n1 < 1 || (x_lead != lead && x_lead != 1)) return fail(JUBIK_ERR_ARG, "convolve: bad shape");
    const JbFftOps *orow = find_ops(n1), *ocol = find_ops(n0);
    const bool small = (size_t)n0 * n1 < (size_t)32 * 32;
    if (small || !orow || !ocol) {
        if ((size_t)n0 * n1 > (size_t)256 * 256)
            return fail(JUBIK_ERR_UNSUPPORTED, "stand-alone convolve: " + std::to_string(n0) + " x " + std::to_string(n1) +
                        " has a side that is not a compiled transform length (jubik_fft_length_supported) and is too large "
                        "for the direct sum (<= 256 x 256)");
        JB_CK(jb_launch(k_convolve_direct, dim3(1184), dim3(128), 0, st, x, y, lead, x_lead, n0, n1, dvol, out));
        return JUBIK_OK;
    }
    // FFT path: rows r2c of both operands, column pass (FFT(y) kept in a scratch slice, FFT(x) * FFT(y), inverse), rows c2r
    const cd *twr_nk, *twr_kn, *twc_nk, *twc_kn;
    if (int rc = conv_twiddles(n1, orow, &twr_nk, &twr_kn, st)) return rc;
    if (int rc = conv_twiddles(n0, ocol, &twc_nk, &twc_kn, st)) return rc;
    const int PH = n1 / 2 + 1, nstrips = jb_cdiv(PH, ocol->CB);
    const size_t bx = (size_t)x_lead * n0 * PH, by = (size_t)lead * n0 * PH, bs = (size_t)lead * nstrips * n0 * ocol->CB;
    cd* ws = nullptr;
#ifndef JB_EMU
    JB_CK(cudaMallocAsync((void**)&ws, (bx + by + bs) * sizeof(cd), st));
#else
    if (dev_malloc((void**)&ws, (bx + by + bs) * sizeof(cd))) return fail(JUBIK_ERR_ALLOC, "convolve workspace");
#endif
    cd *xs = ws, *ys = ws + bx, *sc = ys + by;
    cudaError_t e = orow->conv_rows_r2c(x, xs, twr_nk, (long)x_lead * n0, st);
    if (e == cudaSuccess) e = orow->conv_rows_r2c(y, ys, twr_nk, (long)lead * n0, st);
    if (e == cudaSuccess) e = ocol->conv_cols(xs, ys, sc, twc_nk, PH, lead, x_lead, st);
    if (e == cudaSuccess) e = orow->conv_rows_c2r(ys, out, twr_kn, (long)lead * n0, dvol / ((double)n0 * (double)n1), st);
#ifndef JB_EMU
    cudaFreeAsync(ws, st);
#else
    dev_free(ws);
#endif
    if (e != cudaSuccess) return fail(JUBIK_ERR_CUDA, "convolve kernels -> " + cuda_msg(e));
    return JUBIK_OK;
}

int jubik_psf_interpolate(const jubik_psf_tables* t, int N, int n_patches, double d0, double d1, int kernel_layout,
                          double* out, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (!t || !t->psfs || !t->rs || !t->center_ids || !out) return fail(JUBIK_ERR_ARG, "null pointer");
    if (device_count() < 1) return fail(JUBIK_ERR_CUDA, "no CUDA device: jubik_b200 has no CPU fallback");
    const int E = t->n_energy, nrad = t->n_rad, n = n_patches;
    if (E < 1 || nrad < 1 || t->px < 1 || t->py < 1 || n < 1) return fail(JUBIK_ERR_ARG, "psf tables: bad shape");
    // erosita_psf_utils.py:209-213: the patch grid must divide the (even) image
    if (N < 2 || N % n != 0 || N % 2 != 0) return fail(JUBIK_ERR_ARG, "psf_interpolator: the image size must be even and divisible by npatch");
    const int nn2 = n * n, ps = N / n;
    std::vector<double> wts((size_t)E * nn2 * nrad, 0.0), cid((size_t)E * nrad * 2), rot((size_t)nn2 * 2);
    std::vector<int> perm((size_t)E * nrad);
    // patch centres: k = a*n + b has ra = cen1[n-1-a], dec = cen0[b] (meshgrid 'xy' + flipped ra list, :236-246)
    std::vector<double> ra(nn2), dec(nn2);
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            ra[a * n + b] = ((double)((n - 1 - a) * ps) + (double)ps / 2.0) * d1;
            dec[a * n + b] = ((double)(b * ps) + (double)ps / 2.0) * d0;
        }
    const double half_pi = 1.57079632679489661923;
    for (int k = 0; k < nn2; ++k) {
        const double c0 = ra[k] - t->pointing0, c1 = dec[k] - t->pointing1;
        const double phi = atan2(c1, c0) - half_pi;          // to_r_phi, :17-29
        rot[2 * k] = cos(phi); rot[2 * k + 1] = sin(phi);
    }
    for (int e = 0; e < E; ++e) {
        const double* rs_in = t->rs + (size_t)e * nrad;
        std::vector<int> order(nrad);
        for (int i = 0; i < nrad; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rs_in[a] < rs_in[b]; });   // np.argsort
        std::vector<double> rs(nrad);
        for (int i = 0; i < nrad; ++i) {
            rs[i] = rs_in[order[i]];
            perm[(size_t)e * nrad + i] = order[i];
            const double* c = t->center_ids + ((size_t)e * nrad + order[i]) * 2;
            if (c[0] < 0 || c[1] < 0 || c[0] > t->px || c[1] > t->py) return fail(JUBIK_ERR_ARG, "psf tables: centre index outside the table");
            cid[((size_t)e * nrad + i) * 2] = c[0]; cid[((size_t)e * nrad + i) * 2 + 1] = c[1];
        }
        for (int k = 0; k < nn2; ++k) {
            const double c0 = ra[k] - t->pointing0, c1 = dec[k] - t->pointing1;
            const double r = sqrt(c0 * c0 + c1 * c1);
            double* w = wts.data() + ((size_t)e * nn2 + k) * nrad;      // get_interpolation_weights, :41-72 (overlaps kept)
            if (r <= rs[0]) w[0] += 1.0;
            if (r >= rs[nrad - 1]) w[nrad - 1] += 1.0;
            for (int i = 0; i + 1 < nrad; ++i)
                if (rs[i] < r && rs[i + 1] >= r) {
                    const double wg = (r - rs[i]) / (rs[i + 1] - rs[i]);
                    w[i] += 1.0 - wg; w[i + 1] += wg;
                }
        }
    }
    const size_t bw = wts.size() * sizeof(double), bc = cid.size() * sizeof(double), br = rot.size() * sizeof(double),
                 bp = (perm.size() * sizeof(int) + 15) & ~(size_t)15;
    unsigned char* ws = nullptr;
    if (dev_malloc((void**)&ws, bw + bc + br + bp)) return fail(JUBIK_ERR_ALLOC, "psf_interpolator workspace");
    JbPsfInterp q;
    q.E = E; q.nrad = nrad; q.px = t->px; q.py = t->py; q.N = N; q.nn2 = nn2; q.dx = N / n;
    q.cropped = kernel_layout == JUBIK_KERNEL_CROPPED; q.side = q.cropped ? 2 * q.dx : N;
    q.d0 = d0; q.d1 = d1; q.delta0 = t->delta0; q.delta1 = t->delta1;
    q.psfs = t->psfs;
    q.wts = (const double*)ws; q.cid = (const double*)(ws + bw); q.rot = (const double*)(ws + bw + bc);
    q.perm = (const int*)(ws + bw + bc + br);
    int rc = JUBIK_OK;
    if (h2d(ws, wts.data(), bw, st) || h2d(ws + bw, cid.data(), bc, st) || h2d(ws + bw + bc, rot.data(), br, st) ||
        h2d(ws + bw + bc + br, perm.data(), perm.size() * sizeof(int), st))
        rc = fail(JUBIK_ERR_CUDA, "uploading the interpolation tables failed");
    if (!rc) {
        const size_t npix = (size_t)q.side * q.side;
        const unsigned gx = (unsigned)std::min<size_t>((npix + 255) / 256, 296);
        cudaError_t e = jb_launch(k_psf_interp, dim3(gx, (unsigned)nn2, (unsigned)E), dim3(256), 0, st, q, out);
        if (e != cudaSuccess) rc = fail(JUBIK_ERR_CUDA, "k_psf_interp -> " + cuda_msg(e));
    }
    if (stream_sync(st) && !rc) rc = fail(JUBIK_ERR_CUDA, "stream synchronise failed");   // the host tables die with this frame
    dev_free(ws);
    return rc;
}

int jubik_poisson_energy_and_grad_host(jubik_plan* p, const double* xh, int S, double* eh, double* gh, void* stream_) {
    jb_stream_t st = (jb_stream_t)stream_;
    if (int rc = check_call(p, xh, eh, S)) return rc;
    if (!p->xdev) return fail(JUBIK_ERR_ARG, "plan was created without JUBIK_PLAN_HOST_CALL (no staging buffers for host arrays)");
    if (!p->has_expo || !p->has_data) return fail(JUBIK_ERR_ARG, "plan needs exposures and data for the Poisson energy");
    const JbGeom& g = p->g;
    const size_t npx = (size_t)g.N * g.N;
    size_t nx = (size_t)S * g.E * npx * sizeof(double);
#ifndef JB_EMU
    // (sample, energy bin) planes are independent through the whole fused chain, so the call is pipelined per plane:
    // the H2D copy of plane c+1 and the D2H copy of the gradient of plane c-1 overlap the kernels of plane c.
    if (gh && S * g.E > 1) {
        // the copy streams start after everything already queued on the caller's stream
        JB_CK(cudaEventRecord(p->ev_start, st));
        JB_CK(cudaStreamWaitEvent(p->st_in, p->ev_start, 0));
        JB_CK(cudaStreamWaitEvent(p->st_out, p->ev_start, 0));
        const size_t plane_b = npx * sizeof(double);
        for (int c = 0; c < S * g.E; ++c) {
            JB_CK(cudaMemcpyAsync(p->xdev + (size_t)c * npx, xh + (size_t)c * npx, plane_b, cudaMemcpyHostToDevice, p->st_in));
            JB_CK(cudaEventRecord(p->ev_in[c], p->st_in));
        }
        // Planes alternate between two compute streams: a one-plane launch leaves the persistent kernels with a
        // ragged last wave (1024 rows on 888 resident CTAs), which the other stream's plane fills.
        auto d2h_plane = [&](int c) -> int {
            JB_CK(cudaStreamWaitEvent(p->st_out, p->ev_c[c], 0));
            JB_CK(cudaMemcpyAsync(gh + (size_t)c * npx, p->gdev + (size_t)c * npx, plane_b, cudaMemcpyDeviceToHost, p->st_out));
            return 0;
        };
        if (int rc = run_planes_two_streams(p, p->xdev, S, p->gdev, p->ws2_ready != 0, true, st, d2h_plane)) return rc;
        if (int rc = reduce_energy(p, S, p->edev, st)) return rc;
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

// Sample- / module-sharded fused evaluation.  S may exceed the plan's max_samples: the samples run in chunks of
// max_samples (the workspace holds one chunk) and their gradients are folded into `flat` chunk after chunk, in a
// fixed order.  xh/flat_h != NULL: the host-buffer form (x arrives chunk by chunk through a two-slot device ring on a
// copy stream, one energy bin at a time; the reduced segments of `flat` leave on another copy stream).
static int sharded_impl(jubik_plan* p, jubik_comm* c, int shard_mode, const double* x, const double* xh, int S,
                        double* grad_local, double* flat, double* flat_h, jb_stream_t st) {
    if (!p) return fail(JUBIK_ERR_ARG, "null plan");
    if ((!x && !xh) || !flat) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (S < 1) return fail(JUBIK_ERR_ARG, "n_samples must be >= 1");
    if (!p->has_expo || !p->has_data) return fail(JUBIK_ERR_ARG, "plan needs exposures and data for the Poisson energy");
    const JbGeom& g = p->g;
    const size_t npl = (size_t)g.N * g.N, npx = (size_t)g.E * npl;
    const int R = c ? c->n_ranks : 1, rank = c ? c->rank : 0;
    const int SM = p->S_max, nchunks = (S + SM - 1) / SM;
    if (shard_mode == JUBIK_SHARD_MODULES) {
        // this rank's modules only: partial energies (S) and partial gradients (S,E,N,N) land in flat and are summed
        if (grad_local) return fail(JUBIK_ERR_ARG, "module sharding has no per-rank gradient output (pass NULL)");
        if (xh) return fail(JUBIK_ERR_ARG, "the host-buffer form shards samples only");
        for (int c0 = 0; c0 < S; c0 += SM) {
            const int Sc = std::min(SM, S - c0);
            if (int rc = energy_and_grad_impl(p, x + (size_t)c0 * npx, Sc, flat + (size_t)S * npx + c0, flat + (size_t)c0 * npx, false, st))
                return rc;
        }
#ifndef JB_EMU
        if (c && c->nccl)
            if (int rc = g_nccl.AllReduce(flat, flat, (size_t)S * npx + S, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->nccl, st))
                return nccl_fail("ncclAllReduce", rc);
#endif
        return JUBIK_OK;
    }
    if (shard_mode != JUBIK_SHARD_SAMPLES) return fail(JUBIK_ERR_ARG, "unknown shard mode");
    if (!grad_local && S > 1 && !p->gstage)
        return fail(JUBIK_ERR_ARG, "plan was created without JUBIK_PLAN_SHARDED: pass grad_local (S,E,N,N) or create the plan with the flag");
    double* eslot = flat + npx + (size_t)rank * S;
    // chunk -> where its per-sample gradients go (S == 1 without grad_local: straight into flat, nothing to fold)
    auto chunk_grad = [&](int c0) -> double* {
        if (grad_local) return grad_local + (size_t)c0 * npx;
        return S == 1 ? flat : p->gstage;
    };
#ifndef JB_EMU
    const bool nccl = c && c->nccl;
    // large planes (BASELINE configs[4]: 2 samples x 7 modules at 2048^2 per bin) also run one bin at a time without a
    // communicator: the per-bin launch order is ~2 % faster there (r2e: 33.7 vs 34.3 ms per chunk); small ones (C2) lose
    // more to ragged waves than they gain and keep the batched launches
    const bool big = (size_t)std::min(SM, S) * g.T * npl >= ((size_t)32 << 20);
    const bool per_bin = p->st_comm && g.E > 1 && ((nccl && R > 1) || xh || big);
    if (xh && !(p->st_comm && p->st_in && p->xdev && p->gdev))
        return fail(JUBIK_ERR_ARG, "the host-buffer sharded call needs a plan created with JUBIK_PLAN_HOST_CALL | JUBIK_PLAN_SHARDED");
    if (per_bin || xh) {
        // One energy bin at a time: as soon as the sample-summed gradient of bin e of the LAST chunk is folded into flat,
        // its all-reduce is issued on the collective's stream (and, host form, its D2H on the copy stream) while the
        // kernels of bin e+1 run.  Only the last bin's all-reduce (which also carries the energies, contiguous behind
        // it) is exposed.
        JbEpi ep; memset(&ep, 0, sizeof(ep));
        ep.mode = JB_EPI_POISSON; ep.out = p->gbuf; ep.expo = p->expo; ep.data = p->data; ep.partial = p->partial;
        JB_CK(cudaEventRecord(p->ev_start, st));
        JB_CK(cudaStreamWaitEvent(p->st_comm, p->ev_start, 0));
        double* ring[2] = {p->xdev, p->gdev};
        auto h2d_chunk = [&](int ci) -> int {      // host form: planes in (bin, sample) order, one event per bin
            const int c0 = ci * SM, Sc = std::min(SM, S - c0), slot = ci & 1;
            if (ci >= 2) JB_CK(cudaStreamWaitEvent(p->st_in, p->ev_free[slot], 0));
            for (int e = 0; e < g.E; ++e) {
                for (int s = 0; s < Sc; ++s)
                    JB_CK(cudaMemcpyAsync(ring[slot] + ((size_t)s * g.E + e) * npl, xh + ((size_t)(c0 + s) * g.E + e) * npl,
                                          npl * sizeof(double), cudaMemcpyHostToDevice, p->st_in));
                JB_CK(cudaEventRecord(p->ev_ring[slot * g.E + e], p->st_in));
            }
            return 0;
        };
        if (xh) {
            JB_CK(cudaStreamWaitEvent(p->st_in, p->ev_start, 0));
            JB_CK(cudaStreamWaitEvent(p->st_out, p->ev_start, 0));
            if (int rc = h2d_chunk(0)) return rc;
        }
        for (int ci = 0; ci < nchunks; ++ci) {
            const int c0 = ci * SM, Sc = std::min(SM, S - c0), slot = ci & 1;
            const bool last_chunk = ci == nchunks - 1;
            if (xh && !last_chunk) if (int rc = h2d_chunk(ci + 1)) return rc;
            const double* xc = xh ? ring[slot] : x + (size_t)c0 * npx;
            double* gl = chunk_grad(c0);
            for (int e = 0; e < g.E; ++e) {
                JbGeom ge = g;
                ge.e0 = e; ge.Es = 1;
                const bool last = e == g.E - 1;
                if (xh) JB_CK(cudaStreamWaitEvent(st, p->ev_ring[slot * g.E + e], 0));
                if (int rc = run_forward_g(p, ge, xc, Sc, ep, 0, st)) return rc;
                if (last) { if (int rc = reduce_energy(p, Sc, eslot + c0, st)) return rc; }
                if (int rc = run_adjoint_g(p, ge, p->gbuf, nullptr, Sc, gl, nullptr, st)) return rc;
                if (last && xh) JB_CK(cudaEventRecord(p->ev_free[slot], st));
                JB_LAUNCHED(p, jb_launch(k_fold_samples, dim3(592), dim3(256), 0, st,
                                         (const double*)(gl == flat ? nullptr : gl + (size_t)e * npl), Sc, npl, npx,
                                         flat + (size_t)e * npl, (int)(ci > 0),
                                         (last && last_chunk) ? flat + npx : nullptr, rank * S, S, R * S));
                if (!last_chunk) continue;
                JB_CK(cudaEventRecord(p->ev_bin[e], st));
                double* seg = flat + (size_t)e * npl;
                const size_t cnt = last ? npl + (size_t)R * S : npl;
                if (nccl) {
                    JB_CK(cudaStreamWaitEvent(p->st_comm, p->ev_bin[e], 0));
                    if (int rc = g_nccl.AllReduce(seg, seg, cnt, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->nccl, p->st_comm))
                        return nccl_fail("ncclAllReduce", rc);
                    if (flat_h) JB_CK(cudaEventRecord(p->ev_bin[e], p->st_comm));
                }
                if (flat_h) {
                    JB_CK(cudaStreamWaitEvent(p->st_out, p->ev_bin[e], 0));
                    JB_CK(cudaMemcpyAsync(flat_h + (size_t)e * npl, seg, cnt * sizeof(double), cudaMemcpyDeviceToHost, p->st_out));
                }
            }
        }
        JB_CK(cudaEventRecord(p->ev_comm, p->st_comm));
        JB_CK(cudaStreamWaitEvent(st, p->ev_comm, 0));
        if (flat_h) {
            JB_CK(cudaEventRecord(p->ev_comm, p->st_out));
            JB_CK(cudaStreamWaitEvent(st, p->ev_comm, 0));
        }
        return JUBIK_OK;
    }
#else
    (void)xh; (void)flat_h;
#endif
    for (int ci = 0; ci < nchunks; ++ci) {
        const int c0 = ci * SM, Sc = std::min(SM, S - c0);
        double* gl = chunk_grad(c0);
        if (int rc = energy_and_grad_impl(p, x + (size_t)c0 * npx, Sc, eslot + c0, gl, false, st)) return rc;
        JB_LAUNCHED(p, jb_launch(k_fold_samples, dim3(1184), dim3(256), 0, st, (const double*)(gl == flat ? nullptr : gl), Sc, npx,
                                 npx, flat, (int)(ci > 0), ci == nchunks - 1 ? flat + npx : nullptr, rank * S, S, R * S));
    }
#ifndef JB_EMU
    if (c && c->nccl) {   // also for a one-rank communicator (NCCL treats it as an in-place copy)
        if (int rc = g_nccl.AllReduce(flat, flat, npx + (size_t)R * S, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->nccl, st))
            return nccl_fail("ncclAllReduce", rc);
    }
#endif
    return JUBIK_OK;
}

int jubik_poisson_energy_and_grad_sharded(jubik_plan* p, jubik_comm* c, int shard_mode, const double* x, int S,
                                          double* grad_local, double* flat, void* stream_) {
    if (!x) return fail(JUBIK_ERR_ARG, "null array pointer");
    return sharded_impl(p, c, shard_mode, x, nullptr, S, grad_local, flat, nullptr, (jb_stream_t)stream_);
}

int jubik_poisson_energy_and_grad_sharded_host(jubik_plan* p, jubik_comm* c, const double* x_host, int S, double* flat_host,
                                               void* stream_) {
#ifndef JB_EMU
    jb_stream_t st = (jb_stream_t)stream_;
    if (!x_host || !flat_host) return fail(JUBIK_ERR_ARG, "null array pointer");
    if (!p || !p->flatdev) return fail(JUBIK_ERR_ARG, "the host-buffer sharded call needs a plan created with JUBIK_PLAN_HOST_CALL | JUBIK_PLAN_SHARDED");
    const int R = c ? c->n_ranks : 1;
    if ((size_t)R * S > (size_t)JB_MAX_SLOTS) return fail(JUBIK_ERR_ARG, "more than 4096 samples over all ranks");
    if (int rc = sharded_impl(p, c, JUBIK_SHARD_SAMPLES, nullptr, x_host, S, nullptr, p->flatdev, flat_host, st)) return rc;
    if (stream_sync(st)) return fail(JUBIK_ERR_CUDA, "stream synchronise failed");
    return JUBIK_OK;
#else
    (void)p; (void)c; (void)x_host; (void)S; (void)flat_host; (void)stream_;
    return fail(JUBIK_ERR_UNSUPPORTED, "host-buffer calls need CUDA");
#endif
}

}  // extern "C"
