/* jubik_b200.h -- C ABI of the B200-native J-UBIK instrument-response hot path.
 *
 * The reference (NIFTy-PPL/J-UBIK) has no FFI of its own: the hot path is a set of Python
 * callables built by its response/likelihood builders and lowered by JAX/XLA.  Each entry point
 * below replaces the XLA lowering of one of those callables; a jax.ffi / ctypes / cffi shim binds
 * them 1:1.  Reference interfaces replaced (paths under the reference repo):
 *
 *   jubik_plan_create          jubik/convolve.py:104-131 (PSF crop/pad/shift/normalise + its FFT, hoisted
 *                              out of the per-call path), jubik/response.py:12-45,48-112 (exposure cut,
 *                              readout mask), instruments/erosita/erosita_response.py:307,327-328
 *                              (pixel area, effective area), instruments/*_likelihood.py (data binding)
 *   jubik_psf_apply            vmap(linpatch_convolve) -- erosita_response.py:209-212,
 *                              chandra_response.py:127-130, jubik/convolve.py:66-145
 *   jubik_psf_apply_T          jax.linear_transpose / VJP of the above
 *   jubik_response_apply       R = mask(exposure(psf(x * pixel_area))) -- erosita_response.py:349-350,
 *                              chandra_response.py:138-142 (dense layout, masked pixels = 0)
 *   jubik_response_apply_T     its transpose
 *   jubik_poisson_energy_and_grad   jft.Poissonian(d).amend(R): value_and_grad of
 *                              sum(lambda) - d.log(lambda) -- erosita_likelihood.py:85-89,
 *                              chandra_likelihood.py:37, formula demos/x-ray-imaging.py:313
 *   jubik_poisson_metric_apply R^T diag(1/lambda) R v  (Fisher metric NIFTy builds by linearize+transpose)
 *   jubik_poisson_left_sqrt_metric   R^T (t / sqrt(lambda)): Poissonian left_sqrt_metric pulled back through R,
 *                              what NIFTy's geoVI/MGVI sample draws apply
 *   jubik_poisson_energy_and_grad_logsky   the same value_and_grad with the sky-model tail x = exp(log_sky)
 *                              (jubik/sky_models.py:340-341) and its VJP grad * x folded in
 *
 * Conventions: every array pointer is a DEVICE pointer to contiguous C-order memory (float64 unless
 * stated); `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous
 * on that stream; return value 0 = ok, otherwise an error code whose text is jubik_last_error().
 * No allocation happens after plan creation: the staging of the optional entry points (host-buffer call, metric,
 * sharded call) is requested with jubik_plan_desc.flags and allocated by jubik_plan_create (the one exception is
 * jubik_plan_set_data on a plan created without data, which allocates the counts plane once).  Kernel selection is a
 * pure function of the descriptor -- no environment variables.  A plan owns one workspace: use one plan per stream.
 * There is no CPU fallback: without a CUDA device every entry point fails.
 */
#ifndef JUBIK_B200_H
#define JUBIK_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jubik_plan jubik_plan;

enum { JUBIK_KERNEL_FULL = 0,    /* (T, n*n, E, N, N) origin-centred, as psf_interpolator returns   */
       JUBIK_KERNEL_CROPPED = 1  /* (T, n*n, E, 2dx, 2dx) centred = what convolve.py:107-111 extracts */ };
enum { JUBIK_DATA_F64 = 0, JUBIK_DATA_I64 = 1 };

/* Optional staging, allocated at plan creation (no allocation happens afterwards).  An entry point whose staging was
 * not requested fails with JUBIK_ERR_ARG. */
enum { JUBIK_PLAN_HOST_CALL = 1,  /* jubik_poisson_energy_and_grad_host: device copies of x / grad / energy, copy streams,
                                     and the second compute stream's workspace (one sample's worth of spectra)          */
       JUBIK_PLAN_METRIC = 2,     /* jubik_poisson_set_point / metric / left-sqrt-metric: (S,T,E,N,N) linearisation point */
       JUBIK_PLAN_SHARDED = 4     /* jubik_poisson_energy_and_grad_sharded: per-sample gradient staging (used when
                                     grad_local is NULL) and the collective's stream                                     */ };

enum { JUBIK_OK = 0, JUBIK_ERR_ARG = 1, JUBIK_ERR_UNSUPPORTED = 2, JUBIK_ERR_CUDA = 3, JUBIK_ERR_ALLOC = 4 };

typedef struct jubik_plan_desc {
    int32_t n0, n1;            /* spatial shape of the sky (square, even)                       */
    int32_t n_energy;          /* E                                                             */
    int32_t n_modules;         /* T: telescope modules / observations (vmap axis of the PSF)    */
    int32_t max_samples;       /* S_max: leading batch axis the workspace is sized for          */
    int32_t n_patches;         /* n_patches_per_axis (>= 3, convolve.py:107-111)                */
    int32_t margin;            /* m                                                             */
    int32_t kernel_layout;     /* JUBIK_KERNEL_*                                                */
    int32_t data_dtype;        /* JUBIK_DATA_*                                                  */
    int32_t has_exposure_cut;  /* 0: exposure_cut=None                                          */
    int32_t fft_length;        /* 0 = choose; otherwise force this transform length (>= P)      */
    int32_t flags;             /* JUBIK_PLAN_*: which optional staging the plan allocates        */
    double d0, d1;             /* pixel distances (domain.distances[-2:]); cancel analytically   */
    double exposure_cut;
    double pixel_area;         /* x * pixel_area in R (1.0 if unused)                            */
    const double* kernel;      /* device, layout above                                          */
    const double* exposure;    /* device (T,E,N,N) raw exposures, or NULL (psf-only plan)       */
    const double* eff_area;    /* device (T,E) or NULL                                          */
    const void* data;          /* device (T,E,N,N) counts, dense, masked entries ignored; or NULL*/
} jubik_plan_desc;

typedef struct jubik_plan_info {
    int32_t N, n, m, dx, w, P, L, PH, S_max, T, E;
    int32_t exact_period;      /* 1: circular period == P (reference semantics for any PSF support) */
    int32_t psf_half_support;  /* max |offset| with a non-zero PSF value                            */
    int32_t fft_r1, fft_r2, col_block;
    int32_t alias_fold;        /* 1: L > P with a PSF wider than the margin: linear convolution folded modulo P     */
    int32_t alias_ext;         /* rows / columns at either end of the P-window that alias (half-support - margin)   */
    int32_t fast_path;         /* 1: TMA column pipeline + warp-private c2r rows; 0: generic kernels                */
    int32_t flags;             /* the descriptor's JUBIK_PLAN_* flags                                               */
    size_t bytes_khat, bytes_workspace, bytes_total;
} jubik_plan_info;

const char* jubik_last_error(void);
int jubik_version(void);
/* 1 if an FFT of this length is compiled in */
int jubik_fft_length_supported(int length);
/* number of CUDA devices visible (0 => every other call fails) */
int jubik_device_count(void);

int jubik_plan_create(jubik_plan** plan, const jubik_plan_desc* desc, void* stream);
int jubik_plan_destroy(jubik_plan* plan);
int jubik_plan_get_info(const jubik_plan* plan, jubik_plan_info* info);
/* (re)bind the observed counts, dense (T,E,N,N) */
int jubik_plan_set_data(jubik_plan* plan, const void* data, int data_dtype, void* stream);
/* Replace the PSF kernels of an existing plan (same geometry): redoes convolve.py:104-131 + the kernel FFT into the
 * resident spectra without reallocating.  This is the eROSITA call form R(x, k) with the kernel as a dynamic argument
 * (erosita_response.py:349-350, erosita_likelihood.py:73-83): the binding calls it only when k actually changed.
 * Fails with JUBIK_ERR_UNSUPPORTED if the new PSF's support no longer fits the plan's transform length. Synchronises. */
int jubik_plan_set_kernel(jubik_plan* plan, const double* kernel, int kernel_layout, void* stream);
/* The same, but only if `kernel` differs from the kernels the resident spectra were built from: a 128-bit content
 * fingerprint (two wrap-around 64-bit sums over the bit patterns, one position-weighted) is taken on the device -- one
 * streaming read of the array, 0.1 ms for the 700 MB of cropped kernels of BASELINE configs[2] -- and compared with the
 * plan's.  This is what R(x, k) costs per call when k is a traced argument whose identity the binding cannot see
 * (FullModel.__call__, erosita_likelihood.py:73-83).  *replanned (may be NULL) <- 1 if the spectra were rebuilt.
 * Synchronises. */
int jubik_plan_ensure_kernel(jubik_plan* plan, const double* kernel, int kernel_layout, int* replanned, void* stream);
/* cudaStreamSynchronize for bindings that have no CUDA runtime of their own at hand (the jax binding synchronises the
 * stream it created the plan / bound the data on before XLA's streams use them) */
int jubik_stream_synchronize(void* stream);
/* kernels launched through this plan so far (bench.py's gpu_launches) */
long long jubik_plan_launch_count(const jubik_plan* plan);

/* x:(S,E,N,N) -> out:(S,T,E,N,N) */
int jubik_psf_apply(jubik_plan* plan, const double* x, int n_samples, double* out, void* stream);
/* g:(S,T,E,N,N) -> out:(S,E,N,N) */
int jubik_psf_apply_T(jubik_plan* plan, const double* g, int n_samples, double* out, void* stream);
int jubik_response_apply(jubik_plan* plan, const double* x, int n_samples, double* out, void* stream);
int jubik_response_apply_T(jubik_plan* plan, const double* g, int n_samples, double* out, void* stream);
/* x:(S,E,N,N) -> energy:(S,), grad:(S,E,N,N); grad may be NULL (energy only) */
int jubik_poisson_energy_and_grad(jubik_plan* plan, const double* x, int n_samples,
                                  double* energy, double* grad, void* stream);
/* out = R^T diag(1/(R x)) R v ; x,v,out:(S,E,N,N) */
int jubik_poisson_metric_apply(jubik_plan* plan, const double* x, const double* v, int n_samples,
                               double* out, void* stream);
/* out = R^T (t / sqrt(R x)) ; x,out:(S,E,N,N), t:(S,T,E,N,N) (masked pixels of t are ignored) */
int jubik_poisson_left_sqrt_metric(jubik_plan* plan, const double* x, const double* t, int n_samples,
                                   double* out, void* stream);
/* The same two operators with a CACHED linearisation point: NIFTy's CG applies the metric 60-70 times per sample at a
 * fixed position (lin_maxiter, demos/configs/eROSITA_demo.yaml:161).  jubik_poisson_set_point runs one forward pass
 * and stores expo^2 / lambda(x) (S,T,E,N,N); each *_at_point call is then one forward + one adjoint pass (metric) or one
 * adjoint pass (left-sqrt-metric).  The two calls above are set_point followed by the matching *_at_point call.
 * Needs JUBIK_PLAN_METRIC. */
int jubik_poisson_set_point(jubik_plan* plan, const double* x, int n_samples, void* stream);
int jubik_poisson_metric_apply_at_point(jubik_plan* plan, const double* v, int n_samples, double* out, void* stream);
int jubik_poisson_left_sqrt_metric_at_point(jubik_plan* plan, const double* t, int n_samples, double* out, void* stream);
/* log_sky:(S,E,N,N) -> energy of x = exp(log_sky) and d energy / d log_sky = grad_x * x.  The exp is taken by the r2c
 * row kernel on load and the product by the c2r row kernel on store: no extra pass over the sky. */
int jubik_poisson_energy_and_grad_logsky(jubik_plan* plan, const double* log_sky, int n_samples,
                                         double* energy, double* grad, void* stream);
/* The same on the sky model's PADDED log-sky: sky = exp(log_sky[:E, :N, :N]) (jubik/sky_models.py:340-341, the padding of
 * the correlated-field model).  log_sky, grad: (S, e_pad, n_pad, n_pad) with e_pad >= E, n_pad >= N (even); the r2c row
 * kernel reads the crop in place with the padded strides, the c2r row kernel writes d energy / d log_sky into the crop
 * of grad and the rest of grad is zeroed -- no cropped copy, no padded copy. */
int jubik_poisson_energy_and_grad_logsky_padded(jubik_plan* plan, const double* log_sky, int n_samples, int e_pad, int n_pad,
                                                double* energy, double* grad, void* stream);
/* Stand-alone helpers of the reference's public API (no plan needed; device pointers, fp64):
 *   jubik_slice_patches    jubik/convolve.py:28-63: x (lead,h,w) -> out (n*n, lead, 2dx+2dr, 2dy+2dr), patch k = i*n+j at
 *                          row offset j*dx, column offset i*dy of the dx//2-padded array (dynamic_slice clamping)
 *   jubik_slice_patches_T  its exact transpose (jax.linear_transpose): patches -> (lead,h,w), deterministic gather
 *   jubik_convolve         jubik/convolve.py:148-176: dvol * Re ifftn(fftn(x) * fftn(y)) over the last two axes of x
 *                          (x_lead,n0,n1) and y (lead,n0,n1); x_lead == 1 is the broadcasting branch :167-171.  When both
 *                          sides are compiled transform lengths (jubik_fft_length_supported; any of them, rectangular
 *                          too) it runs as r2c rows -> column FFTs * product * inverse -> c2r rows on the library's FFT
 *                          kernels (workspace: stream-ordered cudaMallocAsync); other shapes, and anything below
 *                          32 x 32, as a direct circular sum (n0*n1 <= 256*256). */
int jubik_slice_patches(const double* x, int lead, int h, int w, int n_patches, int dr, double* out, void* stream);
int jubik_slice_patches_T(const double* patches, int lead, int h, int w, int n_patches, int dr, double* out, void* stream);
int jubik_convolve(const double* x, const double* y, int lead, int x_lead, int n0, int n1, double dvol, double* out,
                   void* stream);
/* Device PSF builder: jubik/instruments/erosita/erosita_psf_utils.py:198-264 psf_interpolator with
 * :94-171 get_psf, for ONE telescope module.  The PSF tables of the off-axis radii bracketing every patch centre are
 * rotated to the patch's position angle and bilinearly resampled (map_coordinates order=1, mode='nearest') onto the
 * sky grid.  out: DEVICE (n*n, E, side, side) -- JUBIK_KERNEL_FULL: side = N, the reference's origin-centred layout;
 * JUBIK_KERNEL_CROPPED: side = 2*int(N/n), the centred crop that convolve.py:107-111 extracts, i.e. exactly what
 * jubik_plan_create consumes, so the (T, n*n, E, N, N) array (17-70 GB at the BASELINE configurations) never exists.
 * d0, d1 = domain.distances[-2:].  Fails like the reference (ValueError -> JUBIK_ERR_ARG) when N is odd or not divisible
 * by n_patches.  Synchronises. */
typedef struct jubik_psf_tables {
    int32_t n_energy, n_rad, px, py;
    const double* psfs;            /* DEVICE (n_energy, n_rad, px, py): 'psfs' of every energy's psf_infos, any radius order */
    const double* rs;              /* HOST   (n_energy, n_rad): 'rs', the off-axis radius of each table                      */
    const double* center_ids;      /* HOST   (n_energy, n_rad, 2): 'patch_center_ids', index of the PSF centre in each table */
    double delta0, delta1;         /* 'patch_deltas': bin size of the tables                                                */
    double pointing0, pointing1;   /* 'pointing_center' of this module                                                      */
} jubik_psf_tables;
int jubik_psf_interpolate(const jubik_psf_tables* tables, int N, int n_patches, double d0, double d1, int kernel_layout,
                          double* out, void* stream);
/* Same as jubik_poisson_energy_and_grad with HOST buffers (pinned or pageable): H2D of x, the kernels,
 * D2H of energy and grad, all on `stream`, then a stream synchronise.  This is the end-to-end call. */
int jubik_poisson_energy_and_grad_host(jubik_plan* plan, const double* x_host, int n_samples,
                                       double* energy_host, double* grad_host, void* stream);

/* Multi-GPU variants (one process per GPU).  The reference has no multi-device path: its VI driver maps the
 * likelihood over posterior samples on one device (jft.optimize_kl, erosita_likelihood.py:85-89 under jax.vmap) and sums
 * the telescope modules inside one callable (erosita_response.py:209-212).  These entry points shard exactly those two
 * axes and combine the partial log-likelihoods and gradients with ONE ncclAllReduce(sum, fp64) of `flat` on `stream`:
 *   JUBIK_SHARD_SAMPLES  every rank evaluates its own S samples of a replicated plan.
 *       flat = [ sum over ALL samples of d energy / d sky  (E,N,N) | energies (n_ranks, S) ]
 *       grad_local (S,E,N,N) receives this rank's per-sample gradients (NULL: not wanted).
 *   JUBIK_SHARD_MODULES  every rank holds a plan of its own modules (T_local) and evaluates the same S samples.
 *       flat = [ d energy / d sky (S,E,N,N) | energies (S) ], both summed over ranks; grad_local must be NULL.
 * jubik_sharded_flat_size gives the element count of `flat`.  comm == NULL runs the single-rank path (no collective).
 * n_samples may EXCEED the plan's max_samples: the samples are then evaluated in chunks of max_samples (the workspace
 * holds one chunk) and folded into `flat` chunk after chunk in a fixed order -- BASELINE configs[4] on fewer than 8 GPUs
 * (16 samples: 8 chunks of 2 on one GPU).  With more than one rank and more than one energy bin the all-reduce of bin e
 * overlaps the kernels of bin e+1 on a second stream (needs JUBIK_PLAN_SHARDED).
 * jubik_poisson_energy_and_grad_sharded_host is the same call (sample sharding) on HOST buffers -- the end-to-end form:
 * x_host (n_samples,E,N,N) is copied chunk by chunk, one energy bin at a time, through a two-slot device ring while the
 * previous chunk computes; flat_host receives [sum-gradient | energies], each bin as soon as it is reduced; then a
 * stream synchronise.  Needs JUBIK_PLAN_HOST_CALL | JUBIK_PLAN_SHARDED and n_ranks * n_samples <= 4096.
 * NCCL is bound with dlopen("libnccl.so.2") at the first jubik_comm_* call; jubik_comm_create wraps ncclCommInitRank on
 * the current device with a 128-byte ncclUniqueId from jubik_comm_unique_id (rank 0) that the caller distributes;
 * jubik_comm_adopt wraps an ncclComm_t the caller already owns (it is not destroyed by jubik_comm_destroy). */
typedef struct jubik_comm jubik_comm;
enum { JUBIK_SHARD_SAMPLES = 0, JUBIK_SHARD_MODULES = 1 };
int jubik_comm_unique_id(void* id128_out);
int jubik_comm_create(jubik_comm** comm, const void* id128, int rank, int n_ranks);
int jubik_comm_adopt(jubik_comm** comm, void* nccl_comm);
int jubik_comm_destroy(jubik_comm* comm);
int jubik_comm_rank(const jubik_comm* comm);
int jubik_comm_size(const jubik_comm* comm);
long long jubik_sharded_flat_size(const jubik_plan* plan, const jubik_comm* comm, int shard_mode, int n_samples);
int jubik_poisson_energy_and_grad_sharded(jubik_plan* plan, jubik_comm* comm, int shard_mode, const double* x,
                                          int n_samples, double* grad_local, double* flat, void* stream);
int jubik_poisson_energy_and_grad_sharded_host(jubik_plan* plan, jubik_comm* comm, const double* x_host, int n_samples,
                                               double* flat_host, void* stream);

/* Measurement helpers (bench.py): the energy+grad sequence with a CUDA event after every kernel,
 * ms_out[0..*n_out) (room for 48) in launch order = rows_fwd, cols_fwd, rows_out, [poisson_epi,] reduce,
 * rows_adj_in, cols_adj, rows_adj_out (synchronises); and a DFMA micro-benchmark giving this device's FP64 FMA
 * throughput in TFLOP/s. */
int jubik_poisson_profile(jubik_plan* plan, const double* x, int n_samples, double* energy, double* grad,
                          float* ms_out, int* n_out, void* stream);
int jubik_measure_fp64_tflops(double* tflops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
