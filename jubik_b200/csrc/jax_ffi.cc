// jax_ffi.cc -- XLA FFI (jax.ffi) custom-call handlers in front of the C ABI of include/jubik_b200.h.
//
// This is the binding layer BASELINE.json's north_star names: "each call goes to CUDA through a thin C-ABI registered
// as jax.ffi custom calls with a custom_vjp".  The Python half is jubik_b200/jax_binding.py (register_ffi_target,
// custom_vjp on the energy, linear_call on R, vmap through a leading sample axis).  What it replaces in the reference:
//   psf_func / psf_op            erosita_response.py:209-212, chandra_response.py:127-130   -> JubikPsfApply[_T]
//   response_func R(x) / R(x,k)  erosita_response.py:349-350, chandra_response.py:138-142   -> JubikResponseApply[T|K|TK]
//   jft.Poissonian(d).amend(R)   erosita_likelihood.py:85-89, chandra_likelihood.py:37      -> JubikEnergyAndGrad,
//                                                                          JubikMetricApply, JubikLeftSqrtMetric
//   sky = exp(log_sky) tail      sky_models.py:340-341                                      -> JubikEnergyAndGradLogsky
//
// Built only where jaxlib's headers exist (jubik_b200/build.py: `jax.ffi.include_dir()`); this image has no jax, so the
// file is compiled and exercised on a host that has it (tests/test_jax_binding.py, skipped without jax).
//
// Conventions: every handler takes the plan as the int64 attribute "plan" (the jubik_plan* the Python side created with
// jubik_plan_create and keeps alive), runs on XLA's stream, and treats ALL leading axes of its first operand as the
// sample axis S of the C ABI (that is what jax.vmap adds with vmap_method="expand_dims" / "broadcast_all").  S larger
// than the plan's max_samples is served in chunks.  No allocation, no synchronisation (except JubikResponseApplyK, whose
// kernel fingerprint is read back -- see jubik_plan_ensure_kernel).
#include <cuda_runtime_api.h>

#include <cstdint>
#include <string>

#include "../../include/jubik_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

struct Shape {
    jubik_plan* plan;
    jubik_plan_info info;
    int64_t sky;   // E*N*N
    int64_t det;   // T*E*N*N
};

ffi::Error err(const std::string& what) { return ffi::Error(ffi::ErrorCode::kInternal, what + ": " + jubik_last_error()); }
ffi::Error bad(const std::string& what) { return ffi::Error(ffi::ErrorCode::kInvalidArgument, what); }

ffi::Error get_shape(int64_t handle, Shape* s) {
    s->plan = reinterpret_cast<jubik_plan*>(static_cast<intptr_t>(handle));
    if (!s->plan) return bad("jubik_b200: null plan attribute");
    if (jubik_plan_get_info(s->plan, &s->info)) return err("jubik_plan_get_info");
    s->sky = (int64_t)s->info.E * s->info.N * s->info.N;
    s->det = s->sky * s->info.T;
    return ffi::Error::Success();
}

// samples = element count / per-sample size; the trailing axes must be the plan's
ffi::Error samples_of(const ffi::Buffer<ffi::F64>& b, int64_t per_sample, const char* what, int64_t* S) {
    const int64_t n = (int64_t)b.element_count();
    if (per_sample <= 0 || n % per_sample || n == 0)
        return bad(std::string("jubik_b200: ") + what + " does not have the plan's trailing shape");
    *S = n / per_sample;
    return ffi::Error::Success();
}

// run f(offset_in_samples, n_samples) over chunks of at most max_samples
template <class F>
ffi::Error chunked(const Shape& s, int64_t S, const char* name, F&& f) {
    for (int64_t c0 = 0; c0 < S; c0 += s.info.S_max) {
        const int Sc = (int)((S - c0) < s.info.S_max ? (S - c0) : s.info.S_max);
        if (f(c0, Sc)) return err(name);
    }
    return ffi::Error::Success();
}

// ---- linear operators: x ([...],E,N,N) -> ([...],T,E,N,N) and back
using Fwd = int (*)(jubik_plan*, const double*, int, double*, void*);

ffi::Error apply(cudaStream_t st, int64_t handle, ffi::Buffer<ffi::F64> in, ffi::ResultBuffer<ffi::F64> out, Fwd fn,
                 bool to_detector, const char* name) {
    Shape s;
    if (auto e = get_shape(handle, &s); e.failure()) return e;
    int64_t S = 0;
    const int64_t isz = to_detector ? s.sky : s.det, osz = to_detector ? s.det : s.sky;
    if (auto e = samples_of(in, isz, name, &S); e.failure()) return e;
    if ((int64_t)out->element_count() != S * osz) return bad(std::string("jubik_b200: bad output shape for ") + name);
    const double* ip = in.typed_data();
    double* op = out->typed_data();
    return chunked(s, S, name, [&](int64_t c0, int Sc) { return fn(s.plan, ip + c0 * isz, Sc, op + c0 * osz, st); });
}

ffi::Error PsfApplyImpl(cudaStream_t st, int64_t plan, ffi::Buffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> y) {
    return apply(st, plan, x, y, jubik_psf_apply, true, "jubik_psf_apply");
}
ffi::Error PsfApplyTImpl(cudaStream_t st, int64_t plan, ffi::Buffer<ffi::F64> g, ffi::ResultBuffer<ffi::F64> y) {
    return apply(st, plan, g, y, jubik_psf_apply_T, false, "jubik_psf_apply_T");
}
ffi::Error ResponseApplyImpl(cudaStream_t st, int64_t plan, ffi::Buffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> y) {
    return apply(st, plan, x, y, jubik_response_apply, true, "jubik_response_apply");
}
ffi::Error ResponseApplyTImpl(cudaStream_t st, int64_t plan, ffi::Buffer<ffi::F64> g, ffi::ResultBuffer<ffi::F64> y) {
    return apply(st, plan, g, y, jubik_response_apply_T, false, "jubik_response_apply_T");
}

// R(x, k) / psf(x, k): the eROSITA call form with the kernel as an operand (erosita_response.py:209-212,349-350).  The
// resident spectra are rebuilt only when k's content differs from what they were built from.
ffi::Error with_kernel(cudaStream_t st, int64_t handle, int64_t layout, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> k,
                       ffi::ResultBuffer<ffi::F64> y, Fwd fn, bool to_detector, const char* name) {
    Shape s;
    if (auto e = get_shape(handle, &s); e.failure()) return e;
    const int64_t side = layout == JUBIK_KERNEL_CROPPED ? s.info.w : s.info.N;
    const int64_t want = (int64_t)s.info.T * s.info.n * s.info.n * s.info.E * side * side;
    if ((int64_t)k.element_count() != want) return bad("jubik_b200: kernel operand does not match the plan's geometry");
    if (jubik_plan_ensure_kernel(s.plan, k.typed_data(), (int)layout, nullptr, st)) return err("jubik_plan_ensure_kernel");
    return apply(st, handle, x, y, fn, to_detector, name);
}
ffi::Error ResponseApplyKImpl(cudaStream_t st, int64_t plan, int64_t layout, ffi::Buffer<ffi::F64> x,
                              ffi::Buffer<ffi::F64> k, ffi::ResultBuffer<ffi::F64> y) {
    return with_kernel(st, plan, layout, x, k, y, jubik_response_apply, true, "jubik_response_apply(x, k)");
}
ffi::Error ResponseApplyTKImpl(cudaStream_t st, int64_t plan, int64_t layout, ffi::Buffer<ffi::F64> g,
                               ffi::Buffer<ffi::F64> k, ffi::ResultBuffer<ffi::F64> y) {
    return with_kernel(st, plan, layout, g, k, y, jubik_response_apply_T, false, "jubik_response_apply_T(g, k)");
}
ffi::Error PsfApplyTKImpl(cudaStream_t st, int64_t plan, int64_t layout, ffi::Buffer<ffi::F64> g, ffi::Buffer<ffi::F64> k,
                          ffi::ResultBuffer<ffi::F64> y) {
    return with_kernel(st, plan, layout, g, k, y, jubik_psf_apply_T, false, "jubik_psf_apply_T(g, k)");
}
ffi::Error PsfApplyKImpl(cudaStream_t st, int64_t plan, int64_t layout, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> k,
                         ffi::ResultBuffer<ffi::F64> y) {
    return with_kernel(st, plan, layout, x, k, y, jubik_psf_apply, true, "jubik_psf_apply(x, k)");
}

// ---- fused Poisson energy + gradient (the custom_vjp's forward pass returns both; the backward pass is a multiply)
ffi::Error energy_grad(cudaStream_t st, int64_t handle, ffi::Buffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> energy,
                       ffi::ResultBuffer<ffi::F64> grad, bool logsky) {
    Shape s;
    if (auto e = get_shape(handle, &s); e.failure()) return e;
    int64_t S = 0;
    if (auto e = samples_of(x, s.sky, "sky", &S); e.failure()) return e;
    if ((int64_t)energy->element_count() != S || (int64_t)grad->element_count() != S * s.sky)
        return bad("jubik_b200: bad output shapes for energy_and_grad");
    const double* xp = x.typed_data();
    double *ep = energy->typed_data(), *gp = grad->typed_data();
    return chunked(s, S, "jubik_poisson_energy_and_grad", [&](int64_t c0, int Sc) {
        return logsky ? jubik_poisson_energy_and_grad_logsky(s.plan, xp + c0 * s.sky, Sc, ep + c0, gp + c0 * s.sky, st)
                      : jubik_poisson_energy_and_grad(s.plan, xp + c0 * s.sky, Sc, ep + c0, gp + c0 * s.sky, st);
    });
}
ffi::Error EnergyAndGradImpl(cudaStream_t st, int64_t plan, ffi::Buffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> e,
                             ffi::ResultBuffer<ffi::F64> g) {
    return energy_grad(st, plan, x, e, g, false);
}
ffi::Error EnergyAndGradLogskyImpl(cudaStream_t st, int64_t plan, ffi::Buffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> e,
                                   ffi::ResultBuffer<ffi::F64> g) {
    return energy_grad(st, plan, x, e, g, true);
}
// the sky model's padded log-sky: sky = exp(log_sky[:E, :N, :N]) read in place (sky_models.py:340-341); x, grad:
// ([...], e_pad, n_pad, n_pad)
ffi::Error EnergyAndGradLogskyPaddedImpl(cudaStream_t st, int64_t handle, int64_t e_pad, int64_t n_pad, ffi::Buffer<ffi::F64> x,
                                         ffi::ResultBuffer<ffi::F64> energy, ffi::ResultBuffer<ffi::F64> grad) {
    Shape s;
    if (auto e = get_shape(handle, &s); e.failure()) return e;
    const int64_t per = e_pad * n_pad * n_pad;
    int64_t S = 0;
    if (auto e = samples_of(x, per, "padded log_sky", &S); e.failure()) return e;
    if ((int64_t)energy->element_count() != S || (int64_t)grad->element_count() != S * per)
        return bad("jubik_b200: bad output shapes for energy_and_grad_logsky_padded");
    const double* xp = x.typed_data();
    double *ep = energy->typed_data(), *gp = grad->typed_data();
    return chunked(s, S, "jubik_poisson_energy_and_grad_logsky_padded", [&](int64_t c0, int Sc) {
        return jubik_poisson_energy_and_grad_logsky_padded(s.plan, xp + c0 * per, Sc, (int)e_pad, (int)n_pad, ep + c0,
                                                           gp + c0 * per, st);
    });
}
// energy only (no adjoint chain): the primal of the custom_vjp when no gradient is requested
ffi::Error EnergyImpl(cudaStream_t st, int64_t handle, ffi::Buffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> energy) {
    Shape s;
    if (auto e = get_shape(handle, &s); e.failure()) return e;
    int64_t S = 0;
    if (auto e = samples_of(x, s.sky, "sky", &S); e.failure()) return e;
    if ((int64_t)energy->element_count() != S) return bad("jubik_b200: bad output shape for energy");
    const double* xp = x.typed_data();
    double* ep = energy->typed_data();
    return chunked(s, S, "jubik_poisson_energy_and_grad", [&](int64_t c0, int Sc) {
        return jubik_poisson_energy_and_grad(s.plan, xp + c0 * s.sky, Sc, ep + c0, nullptr, st);
    });
}

// ---- Fisher metric R^T diag(1/lambda) R v and left-sqrt-metric R^T (t / sqrt(lambda)) at x
ffi::Error MetricApplyImpl(cudaStream_t st, int64_t handle, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> v,
                           ffi::ResultBuffer<ffi::F64> out) {
    Shape s;
    if (auto e = get_shape(handle, &s); e.failure()) return e;
    int64_t S = 0, Sv = 0;
    if (auto e = samples_of(x, s.sky, "x", &S); e.failure()) return e;
    if (auto e = samples_of(v, s.sky, "v", &Sv); e.failure()) return e;
    if (S != Sv || (int64_t)out->element_count() != S * s.sky) return bad("jubik_b200: metric operands disagree");
    const double *xp = x.typed_data(), *vp = v.typed_data();
    double* op = out->typed_data();
    return chunked(s, S, "jubik_poisson_metric_apply", [&](int64_t c0, int Sc) {
        return jubik_poisson_metric_apply(s.plan, xp + c0 * s.sky, vp + c0 * s.sky, Sc, op + c0 * s.sky, st);
    });
}
ffi::Error LeftSqrtMetricImpl(cudaStream_t st, int64_t handle, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> t,
                              ffi::ResultBuffer<ffi::F64> out) {
    Shape s;
    if (auto e = get_shape(handle, &s); e.failure()) return e;
    int64_t S = 0, St = 0;
    if (auto e = samples_of(x, s.sky, "x", &S); e.failure()) return e;
    if (auto e = samples_of(t, s.det, "t", &St); e.failure()) return e;
    if (S != St || (int64_t)out->element_count() != S * s.sky) return bad("jubik_b200: left_sqrt_metric operands disagree");
    const double *xp = x.typed_data(), *tp = t.typed_data();
    double* op = out->typed_data();
    return chunked(s, S, "jubik_poisson_left_sqrt_metric", [&](int64_t c0, int Sc) {
        return jubik_poisson_left_sqrt_metric(s.plan, xp + c0 * s.sky, tp + c0 * s.det, Sc, op + c0 * s.sky, st);
    });
}
// Cached linearisation point (NIFTy's CG: one position, 60-70 applies).  set_point returns a token (one double) that the
// *_at_point calls take as an operand, so that XLA orders them by data dependence.  One chunk only: S <= max_samples.
ffi::Error SetPointImpl(cudaStream_t st, int64_t handle, ffi::Buffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> token) {
    Shape s;
    if (auto e = get_shape(handle, &s); e.failure()) return e;
    int64_t S = 0;
    if (auto e = samples_of(x, s.sky, "x", &S); e.failure()) return e;
    if (S > s.info.S_max) return bad("jubik_b200: set_point takes at most max_samples samples");
    if (jubik_poisson_set_point(s.plan, x.typed_data(), (int)S, st)) return err("jubik_poisson_set_point");
    if (cudaMemsetAsync(token->typed_data(), 0, sizeof(double) * token->element_count(), st) != cudaSuccess)
        return ffi::Error(ffi::ErrorCode::kInternal, "jubik_b200: cudaMemsetAsync failed");
    return ffi::Error::Success();
}
ffi::Error MetricAtPointImpl(cudaStream_t st, int64_t handle, ffi::Buffer<ffi::F64> v, ffi::Buffer<ffi::F64> /*token*/,
                             ffi::ResultBuffer<ffi::F64> out) {
    Shape s;
    if (auto e = get_shape(handle, &s); e.failure()) return e;
    int64_t S = 0;
    if (auto e = samples_of(v, s.sky, "v", &S); e.failure()) return e;
    if ((int64_t)out->element_count() != S * s.sky) return bad("jubik_b200: bad output shape for metric_at_point");
    if (jubik_poisson_metric_apply_at_point(s.plan, v.typed_data(), (int)S, out->typed_data(), st))
        return err("jubik_poisson_metric_apply_at_point");
    return ffi::Error::Success();
}
ffi::Error LeftSqrtMetricAtPointImpl(cudaStream_t st, int64_t handle, ffi::Buffer<ffi::F64> t,
                                     ffi::Buffer<ffi::F64> /*token*/, ffi::ResultBuffer<ffi::F64> out) {
    Shape s;
    if (auto e = get_shape(handle, &s); e.failure()) return e;
    int64_t S = 0;
    if (auto e = samples_of(t, s.det, "t", &S); e.failure()) return e;
    if ((int64_t)out->element_count() != S * s.sky) return bad("jubik_b200: bad output shape for left_sqrt_metric_at_point");
    if (jubik_poisson_left_sqrt_metric_at_point(s.plan, t.typed_data(), (int)S, out->typed_data(), st))
        return err("jubik_poisson_left_sqrt_metric_at_point");
    return ffi::Error::Success();
}

}  // namespace

#define JB_BIND_1_1() ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("plan") \
                          .Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
#define JB_BIND_2_1() ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("plan") \
                          .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
#define JB_BIND_1_2() ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("plan") \
                          .Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
#define JB_BIND_K() ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("plan").Attr<int64_t>("layout") \
                        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikPsfApply, PsfApplyImpl, JB_BIND_1_1());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikPsfApplyT, PsfApplyTImpl, JB_BIND_1_1());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikResponseApply, ResponseApplyImpl, JB_BIND_1_1());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikResponseApplyT, ResponseApplyTImpl, JB_BIND_1_1());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikResponseApplyK, ResponseApplyKImpl, JB_BIND_K());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikPsfApplyK, PsfApplyKImpl, JB_BIND_K());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikResponseApplyTK, ResponseApplyTKImpl, JB_BIND_K());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikPsfApplyTK, PsfApplyTKImpl, JB_BIND_K());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikEnergy, EnergyImpl, JB_BIND_1_1());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikEnergyAndGrad, EnergyAndGradImpl, JB_BIND_1_2());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikEnergyAndGradLogsky, EnergyAndGradLogskyImpl, JB_BIND_1_2());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikEnergyAndGradLogskyPadded, EnergyAndGradLogskyPaddedImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("plan").Attr<int64_t>("e_pad")
                                  .Attr<int64_t>("n_pad").Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikMetricApply, MetricApplyImpl, JB_BIND_2_1());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikLeftSqrtMetric, LeftSqrtMetricImpl, JB_BIND_2_1());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikSetPoint, SetPointImpl, JB_BIND_1_1());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikMetricAtPoint, MetricAtPointImpl, JB_BIND_2_1());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JubikLeftSqrtMetricAtPoint, LeftSqrtMetricAtPointImpl, JB_BIND_2_1());
