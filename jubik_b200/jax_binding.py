"""jax.ffi binding of libjubik_b200.so: the drop-in for jubik's response / likelihood callables under jax.jit, jax.grad,
jax.vmap and jax.linear_transpose (BASELINE.json north_star: "registered as jax.ffi custom calls with a custom_vjp").

    import jubik_b200.jax_binding as jbx
    rd = jbx.build_response(kernel, exposures, domain, npatch, margin, pixel_area, exposure_cut=cut,
                            effective_area=ea, keys=tm_ids, dynamic_kernel=True)        # erosita_response.py:307-358
    lam = rd['R'](sky, rd['kernel'])                                                     # jft.Vector-like dict of 1-D arrays
    lh  = jft.Poissonian(masked_data).amend(FullModel(...))                              # unchanged reference code, or
    H   = jbx.poisson_energy(rd, masked_data)                                            # fused: one custom call + custom_vjp
    jax.value_and_grad(H)(sky)

What is bound (C handlers in csrc/jax_ffi.cc, C ABI in include/jubik_b200.h):
    psf / psf(x, k)        erosita_response.py:209-212, chandra_response.py:127-130      linear_call (jvp + transpose)
    R_dense / R / R(x, k)  erosita_response.py:349-350, chandra_response.py:138-142      linear_call; readout in jnp
    poisson_energy         jft.Poissonian(d).amend(R): erosita_likelihood.py:85-89       custom_vjp (fused value + grad)
    metric / left_sqrt_metric / set_point + *_at_point                                   NIFTy's CG inside optimize_kl
    poisson_energy_logsky  sky = exp(log_sky), sky_models.py:340-341                     custom_vjp

jax is NOT installable in the image this repository is built in (no wheel, no network): this module is written against
the public jax.ffi API (jax >= 0.4.38: jax.ffi.ffi_call / register_ffi_target / pycapsule / include_dir) and is exercised
by tests/test_jax_binding.py wherever jax with CUDA support exists (skipped otherwise).  It never imports torch.
Importing it without jax raises ImportError; nothing else in the package depends on it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import _lib
from .plan import RawPlan, make_desc

import jax  # noqa: E402  (ImportError here == "no jax": the caller falls back to the torch-carried API)
import jax.numpy as jnp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FFI_LIB = os.path.join(HERE, "libjubik_b200_jax.so")

TARGETS = {
    "jubik_psf_apply": "JubikPsfApply", "jubik_psf_apply_T": "JubikPsfApplyT",
    "jubik_response_apply": "JubikResponseApply", "jubik_response_apply_T": "JubikResponseApplyT",
    "jubik_response_apply_k": "JubikResponseApplyK", "jubik_psf_apply_k": "JubikPsfApplyK",
    "jubik_response_apply_tk": "JubikResponseApplyTK", "jubik_psf_apply_tk": "JubikPsfApplyTK",
    "jubik_energy": "JubikEnergy", "jubik_energy_and_grad": "JubikEnergyAndGrad",
    "jubik_energy_and_grad_logsky": "JubikEnergyAndGradLogsky",
    "jubik_energy_and_grad_logsky_padded": "JubikEnergyAndGradLogskyPadded",
    "jubik_metric_apply": "JubikMetricApply", "jubik_left_sqrt_metric": "JubikLeftSqrtMetric",
    "jubik_set_point": "JubikSetPoint", "jubik_metric_at_point": "JubikMetricAtPoint",
    "jubik_left_sqrt_metric_at_point": "JubikLeftSqrtMetricAtPoint",
}
_registered = False


def build_ffi_library(force=False):
    """Compile csrc/jax_ffi.cc against jaxlib's XLA FFI headers and link it to libjubik_b200.so (in-tree)."""
    src = os.path.join(HERE, "csrc", "jax_ffi.cc")
    if not force and os.path.exists(FFI_LIB) and os.path.getmtime(FFI_LIB) >= os.path.getmtime(src):
        return FFI_LIB
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", src, "-I", jax.ffi.include_dir(), "-I", os.path.join(cuda, "include"),
           "-L", HERE, "-ljubik_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath,$ORIGIN", "-o", FFI_LIB]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError("building the jax.ffi shim failed:\n" + r.stdout[-4000:])
    return FFI_LIB


def register():
    """dlopen the shim and register every handler as a CUDA FFI target (idempotent)."""
    global _registered
    if _registered:
        return
    _lib.load()                                    # libjubik_b200.so first: the shim resolves its symbols against it
    lib = C.CDLL(build_ffi_library())
    for target, symbol in TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")
    _registered = True


def _ptr(a):
    """Device pointer of a committed jax.Array (None passes through)."""
    if a is None:
        return None
    a.block_until_ready()
    return a.unsafe_buffer_pointer()


class JaxPlan:
    """jubik_plan* owned from JAX: PSF spectra / exposure / counts resident on the device of `kernel`.  The arrays are
    only read during construction (the plan keeps its own prepared copies)."""

    def __init__(self, kernel, N, n_patches, margin, exposure=None, exposure_cut=None, eff_area=None, pixel_area=1.0,
                 data=None, max_samples=1, cropped=False, distances=(1.0, 1.0), metric=True):
        register()
        if not jax.config.read("jax_enable_x64"):
            raise RuntimeError("jubik_b200 computes in float64: enable it with jax.config.update('jax_enable_x64', True)")
        f64 = jnp.float64
        kernel = jnp.asarray(kernel, f64)
        self.T, nn2, self.E = int(kernel.shape[0]), int(kernel.shape[1]), int(kernel.shape[2])
        if nn2 != n_patches * n_patches:
            raise ValueError("kernel.shape[1] must be n_patches_per_axis**2")
        self.N, self.n_patches, self.margin, self.S_max = int(N), int(n_patches), int(margin), int(max_samples)
        self.cropped = bool(cropped)
        exposure = None if exposure is None else jnp.asarray(exposure, f64)
        eff_area = None if eff_area is None else jnp.asarray(eff_area, f64)
        data_is_int = False
        if data is not None:
            data = jnp.asarray(data)
            data_is_int = jnp.issubdtype(data.dtype, jnp.integer)
            data = data.astype(jnp.int64 if data_is_int else f64)
        flags = _lib.PLAN_METRIC if (metric and exposure is not None) else 0
        desc = make_desc(N, self.E, self.T, n_patches, margin, _ptr(kernel), cropped, max_samples, _ptr(exposure),
                         exposure_cut, _ptr(eff_area), pixel_area, _ptr(data), data_is_int, distances, flags=flags)
        self.raw = RawPlan(_lib.load(), desc, 0)
        self.info = self.raw.info
        self.handle = np.int64(self.raw.handle.value)

    def set_data(self, data):
        data = jnp.asarray(data)
        is_int = jnp.issubdtype(data.dtype, jnp.integer)
        data = data.astype(jnp.int64 if is_int else jnp.float64)
        if tuple(data.shape) != (self.T, self.E, self.N, self.N):
            raise ValueError("data should have shape (T, E, N, N)")
        self.raw.call("jubik_plan_set_data", C.c_void_p(_ptr(data)), _lib.DATA_I64 if is_int else _lib.DATA_F64, None)
        _lib.check(self.raw.lib, self.raw.lib.jubik_stream_synchronize(None))   # bound before XLA's streams read it

    def close(self):
        self.raw.close()

    # ---- shapes
    def _sky_lead(self, x):
        if x.ndim < 3 or tuple(x.shape[-3:]) != (self.E, self.N, self.N):
            raise ValueError("sky should have shape (..., E, N, N)")
        return tuple(x.shape[:-3])

    def _det_lead(self, g):
        if g.ndim < 4 or tuple(g.shape[-4:]) != (self.T, self.E, self.N, self.N):
            raise ValueError("detector-space input should have shape (..., T, E, N, N)")
        return tuple(g.shape[:-4])

    def _call(self, target, out_shapes, *operands, **attrs):
        outs = [jax.ShapeDtypeStruct(s, jnp.float64) for s in out_shapes]
        fn = jax.ffi.ffi_call(target, outs if len(outs) > 1 else outs[0], vmap_method="expand_dims")
        return fn(*[jnp.asarray(o, jnp.float64) for o in operands], plan=self.handle, **attrs)

    # ---- linear operators (linear_call: jvp, transpose and therefore jax.grad / jax.linear_transpose all work)
    def _linear(self, fwd_target, adj_target):
        det = (self.T, self.E, self.N, self.N)
        sky = (self.E, self.N, self.N)

        def fwd(_, x):
            return self._call(fwd_target, [self._sky_lead(x) + det], x)

        def adj(_, g):
            return self._call(adj_target, [self._det_lead(g) + sky], g)

        def op(x):
            return jax.custom_derivatives.linear_call(fwd, adj, None, jnp.asarray(x, jnp.float64))
        return op

    @property
    def psf(self):
        """vmap(linpatch_convolve) over the modules: (..., E, N, N) -> (..., T, E, N, N)."""
        return self._linear("jubik_psf_apply", "jubik_psf_apply_T")

    @property
    def response(self):
        """Dense R(x) = exposure * psf(x * pixel_area), masked pixels 0."""
        return self._linear("jubik_response_apply", "jubik_response_apply_T")

    def _linear_k(self, fwd_target, adj_target):
        det = (self.T, self.E, self.N, self.N)
        sky = (self.E, self.N, self.N)

        def fwd(k, x):
            layout = _lib.KERNEL_CROPPED if tuple(k.shape[-2:]) != (self.N, self.N) else _lib.KERNEL_FULL
            return self._call(fwd_target, [self._sky_lead(x) + det], x, k, layout=np.int64(layout))

        def adj(k, g):
            layout = _lib.KERNEL_CROPPED if tuple(k.shape[-2:]) != (self.N, self.N) else _lib.KERNEL_FULL
            return self._call(adj_target, [self._det_lead(g) + sky], g, k, layout=np.int64(layout))

        def op(x, k):
            return jax.custom_derivatives.linear_call(fwd, adj, jnp.asarray(k, jnp.float64), jnp.asarray(x, jnp.float64))
        return op

    @property
    def psf_k(self):
        return self._linear_k("jubik_psf_apply_k", "jubik_psf_apply_tk")

    @property
    def response_k(self):
        return self._linear_k("jubik_response_apply_k", "jubik_response_apply_tk")

    # ---- Poisson likelihood
    def _energy_fn(self, target):
        def lead_of(x):
            return self._sky_lead(x)

        @jax.custom_vjp
        def energy(x):
            if target == "jubik_energy_and_grad":
                return self._call("jubik_energy", [lead_of(x)], x)
            return self._call(target, [lead_of(x), x.shape], x)[0]

        def fwd(x):
            e, g = self._call(target, [lead_of(x), x.shape], x)
            return e, g

        def bwd(g, ct):
            return (g * jnp.reshape(ct, ct.shape + (1, 1, 1)),)

        energy.defvjp(fwd, bwd)
        return energy

    @property
    def energy(self):
        """x (..., E, N, N) -> H(x) = sum(lambda) - d . log(lambda) per leading index; jax.grad costs ONE fused call."""
        return self._energy_fn("jubik_energy_and_grad")

    @property
    def energy_logsky(self):
        """The same with the sky-model tail fused: log_sky -> H(exp(log_sky)) (sky_models.py:340-341)."""
        return self._energy_fn("jubik_energy_and_grad_logsky")

    def energy_logsky_padded(self, e_pad, n_pad):
        """The same on the sky model's PADDED log-sky (..., e_pad, n_pad, n_pad): exp(log_sky[:E, :N, :N]) is read in
        place, the gradient comes back padded with zeros outside the crop (sky_models.py:340-341)."""
        trail = (int(e_pad), int(n_pad), int(n_pad))

        def call(x):
            lead = tuple(x.shape[:-3])
            if tuple(x.shape[-3:]) != trail:
                raise ValueError("padded log_sky should have trailing shape %s" % (trail,))
            return self._call("jubik_energy_and_grad_logsky_padded", [lead, tuple(x.shape)], x,
                              e_pad=np.int64(e_pad), n_pad=np.int64(n_pad))

        @jax.custom_vjp
        def energy(x):
            return call(x)[0]

        energy.defvjp(lambda x: call(x), lambda g, ct: (g * jnp.reshape(ct, ct.shape + (1, 1, 1)),))
        return energy

    def metric(self, x, v):
        """R^T diag(1 / R x) R v."""
        return self._call("jubik_metric_apply", [tuple(jnp.shape(x))], x, v)

    def left_sqrt_metric(self, x, t):
        """R^T (t / sqrt(R x)), t in data space (..., T, E, N, N)."""
        return self._call("jubik_left_sqrt_metric", [tuple(jnp.shape(x))], x, t)

    def set_point(self, x):
        """Cache expo^2 / lambda(x); returns the token the *_at_point calls take (orders them after this call)."""
        return self._call("jubik_set_point", [(1,)], x)

    def metric_at_point(self, v, token):
        return self._call("jubik_metric_at_point", [tuple(jnp.shape(v))], v, token)

    def left_sqrt_metric_at_point(self, t, token):
        lead = self._det_lead(jnp.asarray(t))
        return self._call("jubik_left_sqrt_metric_at_point", [lead + (self.E, self.N, self.N)], t, token)


class Vector(dict):
    """Stand-in for jft.Vector where NIFTy is absent: a pytree (dict) with a `.tree` view of itself."""

    @property
    def tree(self):
        return self


def _vector(tree):
    try:
        import nifty.re as jft
        return jft.Vector(tree)
    except Exception:
        return Vector(tree)


def build_response(kernel, exposures, domain, n_patches_per_axis, margin, pixel_area, exposure_cut=None,
                   effective_area=None, keys=None, max_samples=1, cropped=False, dynamic_kernel=False):
    """response_dict of erosita_response.py:307-358 (dynamic_kernel=True: 'kernel' + R(x, k)) and of
    chandra_response.py:127-149 (dynamic_kernel=False: R(x)), every entry a jit-/grad-/vmap-able JAX callable on the
    CUDA path, plus 'mask_adj', 'R_dense' and 'plan'."""
    exposures = np.array(exposures, dtype=np.float64, copy=True)
    if exposure_cut is not None:
        if exposure_cut < 0:
            raise ValueError("exposure_cut should be non-negative or None!")
    kernel = jnp.asarray(kernel, jnp.float64)
    T = int(kernel.shape[0])
    plan = JaxPlan(kernel, domain.shape[-1], n_patches_per_axis, margin, exposure=exposures, exposure_cut=exposure_cut,
                   eff_area=effective_area, pixel_area=pixel_area, max_samples=max_samples, cropped=cropped,
                   distances=tuple(domain.distances[-2:]))
    thr = exposures.copy()
    if exposure_cut is not None:
        thr[thr < exposure_cut] = 0
    mask = thr == 0                                              # static (NumPy): the compaction has static shapes
    thr_j = jnp.asarray(thr)
    eff_j = None if effective_area is None else jnp.asarray(effective_area, jnp.float64)[:, :, None, None]
    key_list = list(keys) if keys is not None else ['masked input'] if T == 1 else list(range(T))
    if len(key_list) != T:
        raise ValueError("length of keys should match the number of flag maps.")
    idx = [np.flatnonzero(~mask[i].ravel()) for i in range(T)]   # readout order of x[i][~mask[i]] (response.py:110)

    def exposure_func(x):
        y = thr_j * x
        return y if eff_j is None else y * eff_j

    def mask_func(x):
        if x.ndim != 4:
            raise ValueError("input should have shape (n, m, q, l)!")
        return _vector({k: x[i].reshape(-1)[idx[i]] for i, k in enumerate(key_list)})

    def mask_adj(v):
        tree = v.tree if hasattr(v, "tree") else v
        planes = [jnp.zeros(mask[i].size, jnp.float64).at[idx[i]].set(jnp.asarray(tree[k], jnp.float64)).reshape(mask[i].shape)
                  for i, k in enumerate(key_list)]
        return (jnp.stack(planes),)

    out = {'pix_area': pixel_area, 'exposure': exposure_func, 'mask': mask_func, 'mask_adj': mask_adj,
           'R_dense': plan.response, 'plan': plan}
    if dynamic_kernel:
        psf_k, resp_k = plan.psf_k, plan.response_k
        out.update({'psf': psf_k, 'R': lambda x, k: mask_func(resp_k(x, k)), 'kernel': kernel})
    else:
        psf, resp = plan.psf, plan.response
        out.update({'psf': psf, 'R': lambda x: mask_func(resp(x))})
    return out


def poisson_energy(response_dict, masked_data, logsky=False):
    """jft.Poissonian(masked_data).amend(R) as ONE fused custom call with a custom_vjp: returns H(sky) (or
    H(log_sky) with logsky=True), differentiable with jax.grad / jax.value_and_grad, batched with jax.vmap."""
    plan = response_dict['plan']
    tree = masked_data.tree if hasattr(masked_data, "tree") else masked_data
    dense = response_dict['mask_adj'](tree)[0]
    plan.set_data(dense)
    return plan.energy_logsky if logsky else plan.energy
