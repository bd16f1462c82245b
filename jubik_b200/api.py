"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as the reference so that callers (and tests) read like
the reference's own:

    Domain                      jubik/data.py:19-29
    linpatch_convolve           jubik/convolve.py:66-145
    slice_patches / _T          jubik/convolve.py:28-63 (+ its jax.linear_transpose, :141-143)
    convolve                    jubik/convolve.py:148-176
    integrate                   jubik/convolve.py:179-184 (host code at set-up time in the reference too)
    build_exposure_function     jubik/response.py:12-45
    build_readout_function      jubik/response.py:48-112
    build_psf_operator          instruments/erosita/erosita_response.py:206-212 (psf_op), chandra :127-130
    build_response              erosita_response.py:307-358 / chandra_response.py:127-149 -> response_dict
                                (dynamic_kernel=True: the eROSITA call form R(x, k) with 'kernel' in the dict)
    psf_interpolator            instruments/erosita/erosita_psf_utils.py:198-264 (PSF tables -> patch kernels, on the device)
    build_mask_adjoint          jax.linear_transpose(response_dict['mask'], ...)  (data.py:156-160) -> 'mask_adj'
    create_mock_data            data.py:138-146 (R(sky[, kernel]) + Poisson draw per key)
    calculate_nwr               diagnostics.py:62-125 (batched R over the samples)
    PoissonLikelihood           jft.Poissonian(masked_data).amend(R)  (erosita_likelihood.py:85-89,
                                chandra_likelihood.py:37)

Arrays are torch float64 CUDA tensors.  The compute runs in libjubik_b200.so through ResponsePlan; torch is
only the memory/stream carrier (plus the trivial un-fused elementwise exposure/readout callables that the
reference also exposes separately).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

from .plan import ResponsePlan


class Domain(NamedTuple):
    """jubik/data.py:19-29 -- only shape[-2:] and distances[-2:] are consumed by the hot path."""
    shape: tuple
    distances: tuple


class Vector:
    """Minimal stand-in for nifty.re.Vector (pytree wrapper around a dict of arrays)."""

    def __init__(self, tree):
        self.tree = tree

    def __getitem__(self, k):
        return self.tree[k]

    def keys(self):
        return self.tree.keys()


def _torch():
    import torch
    return torch


def _as_cuda(a, device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: jubik_b200 has no CPU fallback")
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.as_tensor(a)
    return t.to(device or "cuda", torch.float64)


_PLAN_CACHE = {}


class _CachedPsf:
    """One cached linpatch plan: the plan, the device kernel it was built from (kept alive, so its address cannot be
    recycled for a different PSF) and the identity/version of the caller's object for the cheap hit test."""

    def __init__(self, plan, kern_dev, src):
        self.plan, self.kern_dev = plan, kern_dev
        self.remember(src)

    def remember(self, src):
        torch = _torch()
        self.src_id = id(src)
        self.src_version = src._version if isinstance(src, torch.Tensor) else None
        self.src_ref = src if isinstance(src, torch.Tensor) else None   # pins id(src) for tensors

    def same_object(self, src):
        return self.src_ref is not None and src is self.src_ref and src._version == self.src_version


def linpatch_convolve(x, domain, kernel, n_patches_per_axis, margin):
    """Drop-in for jubik.linpatch_convolve (convolve.py:66-145).

    x: (..., N, N); kernel: (n*n, ..., N, N) origin-centred, same leading axes as x.  One plan (PSF spectra) is cached
    per geometry.  A call with the same kernel tensor (same object, unmodified) or a kernel of equal CONTENT reuses the
    resident spectra; any other kernel re-runs the PSF preparation on the device (jubik_plan_set_kernel), so two
    different PSFs of the same shape can never be confused.  Hot loops should hold a plan via build_psf_operator.
    """
    if not isinstance(domain, Domain):
        raise ValueError("domain has to be an instance of jubik.Domain")
    torch = _torch()
    from ._lib import JubikError
    x = _as_cuda(x)
    kern_dev = _as_cuda(kernel, x.device)
    N = domain.shape[-1]
    lead = tuple(x.shape[:-2])
    E = 1
    for s in lead:
        E *= s
    key = (tuple(kern_dev.shape), N, n_patches_per_axis, margin, x.device.index, tuple(domain.distances[-2:]))
    k5 = kern_dev.reshape((1, kern_dev.shape[0], E) + tuple(kern_dev.shape[-2:])).contiguous()

    def private(t):
        # the cache's copy must not alias the caller's tensor: an in-place update of that tensor would otherwise change
        # both sides of the content comparison below
        shares = isinstance(kernel, torch.Tensor) and t.untyped_storage().data_ptr() == kernel.untyped_storage().data_ptr()
        return t.clone() if shares else t
    ent = _PLAN_CACHE.get(key)
    if ent is not None and not ent.same_object(kernel):
        if not torch.equal(ent.kern_dev, k5):
            try:
                ent.plan.set_kernel(k5, cropped=False)
                ent.kern_dev = private(k5)
            except JubikError:      # the new PSF's support does not fit the cached transform length
                ent.plan.close()
                del _PLAN_CACHE[key]
                ent = None
        if ent is not None:
            ent.remember(kernel)
    if ent is None:
        plan = ResponsePlan(k5, N, n_patches_per_axis, margin, distances=tuple(domain.distances[-2:]))
        if len(_PLAN_CACHE) >= 4:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE))).plan.close()
        ent = _PLAN_CACHE[key] = _CachedPsf(plan, private(k5), kernel)
    y = ent.plan.psf(x.reshape((E, N, N)))[0]
    return y.reshape(lead + (N, N))


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def slice_patches(x, shape, n_patches_per_axis, additional_margin):
    """Drop-in for jubik.convolve.slice_patches (convolve.py:28-63): (..., H, W) -> (n*n, ..., 2dx+2dr, 2dy+2dr)."""
    torch = _torch()
    from ._lib import load, check
    lib = load()
    x = _as_cuda(x)
    h, w = int(shape[-2]), int(shape[-1])
    n, dr = int(n_patches_per_axis), int(additional_margin)
    lead = tuple(x.shape[:-2])
    nl = 1
    for v in lead:
        nl *= v
    dx, dy = int((h - 2 * dr) / n), int((w - 2 * dr) / n)
    out = torch.empty((n * n,) + lead + (2 * dx + 2 * dr, 2 * dy + 2 * dr), dtype=torch.float64, device=x.device)
    check(lib, lib.jubik_slice_patches(C.c_void_p(x.data_ptr()), max(nl, 1), h, w, n, dr, C.c_void_p(out.data_ptr()),
                                       _stream_ptr(torch)))
    return out


def slice_patches_T(patches, shape, n_patches_per_axis, additional_margin):
    """Exact transpose of slice_patches (what jax.linear_transpose yields, convolve.py:141-143)."""
    torch = _torch()
    from ._lib import load, check
    lib = load()
    patches = _as_cuda(patches)
    h, w = int(shape[-2]), int(shape[-1])
    n, dr = int(n_patches_per_axis), int(additional_margin)
    lead = tuple(patches.shape[1:-2])
    nl = 1
    for v in lead:
        nl *= v
    out = torch.empty(lead + (h, w), dtype=torch.float64, device=patches.device)
    check(lib, lib.jubik_slice_patches_T(C.c_void_p(patches.data_ptr()), max(nl, 1), h, w, n, dr,
                                         C.c_void_p(out.data_ptr()), _stream_ptr(torch)))
    return out


def convolve(x, y, domain, axes):
    """Drop-in for jubik.convolve (convolve.py:148-176): dvol * Re ifftn(fftn(x) * fftn(y)) over the last two axes, any
    leading shape, including the broadcasting branch (y.ndim > x.ndim: hx[..., newaxis, :, :] * hy) and its console
    message.  Arrays whose sides are compiled transform lengths run on the library's FFT kernels; anything else (up to
    256 x 256) as a direct circular sum."""
    torch = _torch()
    from ._lib import load, check
    lib = load()
    x = _as_cuda(x)
    y = _as_cuda(y, x.device)
    nd = max(x.dim(), y.dim())
    if sorted(a % nd if a >= 0 else a + nd for a in axes) != [nd - 2, nd - 1] and sorted(axes) != [-2, -1]:
        raise NotImplementedError("convolve: only the last two axes are supported")
    dvol = 1.0
    for a in axes:
        dvol *= float(domain.distances[a])
    n0, n1 = int(y.shape[-2]), int(y.shape[-1])
    if tuple(x.shape[-2:]) != (n0, n1):
        raise ValueError("convolve: the last two axes of both arrays must agree")
    if y.dim() > x.dim():
        print("kernel_shape:", tuple(x.shape))
        print("signal_shape:", tuple(y.shape))
        print("Dimension Inconsistency. Broadcasting PSFs")
        x = x[..., None, :, :]
    out_shape = tuple(torch.broadcast_shapes(tuple(x.shape), tuple(y.shape)))
    lead = 1
    for v in out_shape[:-2]:
        lead *= v
    x_lead1 = x.numel() == n0 * n1                 # one kernel for every signal plane: no expansion needed
    xe = x.reshape(1, n0, n1) if x_lead1 else x.expand(out_shape)
    xe = xe.contiguous()
    ye = y.expand(out_shape).contiguous()
    out = torch.empty(out_shape, dtype=torch.float64, device=x.device)
    if lead == 0:
        return out
    check(lib, lib.jubik_convolve(C.c_void_p(xe.data_ptr()), C.c_void_p(ye.data_ptr()), lead, 1 if x_lead1 else lead,
                                  n0, n1, dvol, C.c_void_p(out.data_ptr()), _stream_ptr(torch)))
    return out


def integrate(x, domain, axes):
    """Drop-in for jubik.convolve.integrate (convolve.py:179-184): dvol * sum over `axes`.  Set-up-time host code in
    the reference (NumPy; used once to normalise the Chandra PSF, chandra_response.py:123): NumPy arrays are summed
    with NumPy, torch tensors (any device) with torch -- no kernel of ours is involved."""
    axes = tuple(axes)
    dvol = 1.0
    for i in axes:
        dvol *= float(domain.distances[i])
    torch = _torch()
    if isinstance(x, torch.Tensor):
        return dvol * torch.sum(x, dim=axes)
    import numpy as np
    return dvol * np.sum(x, axis=axes)


def build_exposure_function(exposures, exposure_cut=None):
    """response.py:12-45: thresholds `exposures` IN PLACE (like the reference) and returns x -> exposures*x."""
    if exposure_cut is not None:
        if exposure_cut < 0:
            raise ValueError("exposure_cut should be non-negative or None!")
        exposures[exposures < exposure_cut] = 0

    def exposure(x):
        return exposures * x
    return exposure


def build_readout_function(flags, threshold=None, keys=None):
    """response.py:48-112: x (T,E,N,N) -> Vector{key: x[i][~mask[i]]}, mask = thresholded flags == 0."""
    if threshold is not None and threshold < 0:
        raise ValueError("threshold should be positive!")
    if threshold is not None:
        flags[flags < threshold] = 0
    mask = flags == 0
    if keys is None:
        keys = ['masked input']
    elif len(keys) != flags.shape[0]:
        raise ValueError("length of keys should match the number of flag maps.")

    def _apply_readout(x):
        if len(mask.shape) != 4:
            raise ValueError("flags should have shape (n, m, q, l)!")
        if len(x.shape) != 4:
            raise ValueError("input should have shape (n, m, q, l)!")
        return Vector({key: x[i][~mask[i]] for i, key in enumerate(keys)})
    return _apply_readout


def build_psf_operator(kernel, domain, n_patches_per_axis, margin, max_samples=1, cropped=False):
    """psf_op of erosita_response.py:209-212 / chandra_response.py:127-130 with the PSF spectra resident.

    kernel: (T, n*n, E, N, N).  Returns (psf_func, plan); psf_func(x) maps ([S,]E,N,N) -> ([S,]T,E,N,N)."""
    if not isinstance(domain, Domain):
        raise ValueError("domain has to be an instance of jubik.Domain")
    kernel = _as_cuda(kernel)
    plan = ResponsePlan(kernel, domain.shape[-1], n_patches_per_axis, margin, max_samples=max_samples,
                        cropped=cropped, distances=tuple(domain.distances[-2:]))
    return plan.psf, plan


def psf_interpolator(domain, npatch, psf_infos, cropped=False, out=None):
    """Drop-in for jubik's psf_interpolator (instruments/erosita/erosita_psf_utils.py:198-264) on the device: the PSF
    tables of one telescope module (`psf_infos`: one dict per energy with 'psfs', 'rs', 'patch_center_ids',
    'patch_deltas', 'pointing_center' -- eROSITA_PSF._get_psf_infos, erosita_psf.py:211-220) resampled onto the sky
    grid for every patch centre.  Returns a float64 CUDA tensor (npatch**2, E, N, N) in the reference's origin-centred
    layout, or with cropped=True the (npatch**2, E, 2dx, 2dx) centred crop that a ResponsePlan(..., cropped=True) consumes
    -- the full-size array (tens of GB for 7 modules at 1024^2 / 2048^2) is then never materialised.  `out` may be a
    preallocated slice, e.g. kernel[t] of a (T, n*n, E, 2dx, 2dx) tensor."""
    import numpy as np
    from . import _lib
    torch = _torch()
    if not isinstance(domain, Domain):
        raise ValueError("domain has to be an instance of jubik.Domain")
    lib = _lib.load()
    N0, N1 = int(domain.shape[-2]), int(domain.shape[-1])
    for ss in (N0, N1):                                   # erosita_psf_utils.py:209-213
        if ss % npatch != 0 or ss % 2 != 0:
            raise ValueError
    if N0 != N1:
        raise NotImplementedError("psf_interpolator: square skies only")
    E = len(psf_infos)
    tabs = [np.asarray(info['psfs'], dtype=np.float64) for info in psf_infos]
    if len({t.shape for t in tabs}) != 1:
        raise ValueError("psf_interpolator: the tables of all energies must share one shape")
    nrad, px, py = tabs[0].shape
    rs = np.ascontiguousarray([np.asarray(info['rs'], dtype=np.float64) for info in psf_infos])
    cid = np.ascontiguousarray([np.asarray(info['patch_center_ids'], dtype=np.float64) for info in psf_infos])
    if rs.shape != (E, nrad) or cid.shape != (E, nrad, 2):
        raise ValueError
    deltas = {tuple(float(v) for v in info['patch_deltas']) for info in psf_infos}
    points = {tuple(float(v) for v in info['pointing_center']) for info in psf_infos}
    if len(deltas) != 1 or len(points) != 1:
        raise ValueError("psf_interpolator: one module at a time (a single bin size and pointing centre)")
    (dl0, dl1), (p0, p1) = deltas.pop(), points.pop()
    dev_tabs = torch.as_tensor(np.ascontiguousarray(np.stack(tabs))).to("cuda" if out is None else out.device)
    dx = N0 // npatch
    side = 2 * dx if cropped else N0
    shape = (npatch * npatch, E, side, side)
    if out is None:
        out = torch.empty(shape, dtype=torch.float64, device=dev_tabs.device)
    elif tuple(out.shape) != shape or out.dtype != torch.float64 or not out.is_cuda or not out.is_contiguous():
        raise ValueError("out must be a contiguous float64 CUDA tensor of shape %s" % (shape,))
    t = _lib.PsfTables(E, nrad, px, py, dev_tabs.data_ptr(), rs.ctypes.data, cid.ctypes.data, dl0, dl1, p0, p1)
    with torch.cuda.device(out.device):
        _lib.check(lib, lib.jubik_psf_interpolate(C.byref(t), N0, int(npatch), float(domain.distances[-2]),
                                                  float(domain.distances[-1]),
                                                  _lib.KERNEL_CROPPED if cropped else _lib.KERNEL_FULL,
                                                  C.c_void_p(out.data_ptr()), _stream_ptr(torch)))
    return out


def build_mask_adjoint(mask, keys):
    """Transpose of the readout (what `jax.linear_transpose(response_dict['mask'], ...)` yields, data.py:156-160,
    sugar_plot.py:315-323, demos/x-ray-imaging.py:268-269): Vector{key: 1-D array of the unmasked pixels} -> a 1-tuple
    holding the dense (T,E,N,N) array with zeros at the masked pixels -- callers take `[0]` like the reference."""
    def mask_adj(v):
        torch = _torch()
        tree = v.tree if isinstance(v, Vector) else v
        out = torch.zeros(mask.shape, dtype=torch.float64, device=mask.device)
        for i, key in enumerate(keys):
            out[i][~mask[i]] = torch.as_tensor(tree[key]).to(mask.device, torch.float64)
        return (out,)
    return mask_adj


def build_response(kernel, exposures, domain, n_patches_per_axis, margin, pixel_area, exposure_cut=None,
                   effective_area=None, keys=None, max_samples=1, cropped=False, data=None, dynamic_kernel=False,
                   **plan_flags):
    """Response dictionary with the reference's keys ('pix_area', 'psf', 'exposure', 'mask', 'R', and with
    dynamic_kernel=True also 'kernel') plus 'mask_adj' and the fused entries ('R_dense', 'R_dense_T', 'plan').

    R(x) = mask(exposure(psf(x * pixel_area)))  -- chandra_response.py:138-149.
    dynamic_kernel=True is the eROSITA form (erosita_response.py:209-212,349-358): psf(x, k) and R(x, k) take the
    kernel as an argument and the dictionary carries it under 'kernel' (callers branch on that key, data.py:138-142;
    FullModel holds it as a dynamic leaf, erosita_likelihood.py:73-83).  The PSF spectra stay resident: a call with the
    dictionary's own kernel object costs nothing extra, any other tensor is fingerprinted on the device and the
    spectra are rebuilt only if its content differs (jubik_plan_ensure_kernel).
    The exposure cut is applied to both the exposure and the readout mask, as both instruments do
    (SURVEY Appendix A10)."""
    torch = _torch()
    if not isinstance(domain, Domain):
        raise ValueError("domain has to be an instance of jubik.Domain")
    kernel = _as_cuda(kernel)
    expo_raw = _as_cuda(exposures, kernel.device).clone()
    T = kernel.shape[0]
    eff = None if effective_area is None else _as_cuda(effective_area, kernel.device)
    plan = ResponsePlan(kernel, domain.shape[-1], n_patches_per_axis, margin, exposure=expo_raw,
                        exposure_cut=exposure_cut, eff_area=eff, pixel_area=pixel_area, data=data,
                        max_samples=max_samples, cropped=cropped, distances=tuple(domain.distances[-2:]), **plan_flags)
    expo_thr = expo_raw.clone()
    tmp = build_exposure_function(expo_thr, exposure_cut)
    if eff is not None:
        def exposure_func(x):
            return tmp(x) * eff[:, :, None, None]
    else:
        exposure_func = tmp
    key_list = list(keys) if keys is not None else list(range(T))
    flags = expo_raw.clone()
    mask_func = build_readout_function(flags, exposure_cut, key_list)
    mask_adj = build_mask_adjoint(flags == 0, key_list)      # build_readout_function thresholded `flags` in place
    out = {'pix_area': pixel_area, 'exposure': exposure_func, 'mask': mask_func, 'mask_adj': mask_adj,
           'R_dense': plan.response, 'R_dense_T': plan.response_T, 'plan': plan}
    if not dynamic_kernel:
        def response_func(x):
            return mask_func(plan.response(x))
        out.update({'psf': plan.psf, 'R': response_func})
        return out

    def use_kernel(k):
        plan.ensure_kernel(k if getattr(k, "is_cuda", False) else _as_cuda(k, plan.device))

    def psf_func(x, k):
        use_kernel(k)
        return plan.psf(x)

    def response_func(x, k):
        use_kernel(k)
        return mask_func(plan.response(x))

    out.update({'psf': psf_func, 'R': response_func, 'kernel': kernel})
    return out


def create_mock_data(response_dict, sky, seed=0):
    """The hot-path part of jubik.create_mock_data (data.py:138-146): masked_mock_data = R(sky[, kernel]) followed by
    one Poisson draw per key, as int64.  (The reference draws with jax.random; the generator differs, the law does not.)"""
    torch = _torch()
    if 'kernel' in response_dict:
        lam = response_dict['R'](sky, response_dict['kernel'])
    else:
        lam = response_dict['R'](sky)
    out = {}
    for i, (key, v) in enumerate(lam.tree.items()):
        gen = torch.Generator(device=v.device)
        gen.manual_seed(int(seed) + i)
        out[key] = torch.poisson(v.clamp(min=0.0), generator=gen).to(torch.int64)
    return Vector(out)


def calculate_nwr(pos, op, data, response_dict, abs=False, min_counts=None, exposure_mask=True, response=True):
    """jubik.calculate_nwr (diagnostics.py:62-125) on the batched response: noise-weighted residuals
    (R(op(x)) - data) / sqrt(R(op(x))) of every sample x of `pos` (leading axis = samples, the reference's vmap), mapped
    back onto the dense grid with 'mask_adj'.  Returns (res (T,S,E,N,N) like vmap(..., out_axes=1), tot_mask)."""
    torch = _torch()
    adj_mask = response_dict['mask_adj']
    has_k = 'kernel' in response_dict

    def R(x):
        if not response:
            return x
        return response_dict['R'](x, response_dict['kernel']) if has_k else response_dict['R'](x)

    def one(x):
        lam = R(op(x))
        if isinstance(lam, Vector):
            return Vector({k: (lam[k] - data[k]) / torch.sqrt(lam[k]) for k in lam.keys()})
        return (lam - data) / torch.sqrt(lam)

    res = torch.stack([adj_mask(one(x))[0] for x in pos], dim=1)
    if abs:
        res = torch.abs(res)
    min_count_mask = None
    if min_counts is not None:
        tree = data.tree if isinstance(data, Vector) else data
        mi = Vector({k: (torch.as_tensor(v) < min_counts).to(torch.float64) for k, v in tree.items()})
        min_count_mask = adj_mask(mi)[0]
    tot_mask = None
    if exposure_mask:
        ones = torch.ones_like(op(pos[0]))
        exp_mask = response_dict['exposure'](ones) == 0.
        tot_mask = exp_mask if min_count_mask is None else torch.logical_or(min_count_mask.bool(), exp_mask)
    elif min_count_mask is not None:
        tot_mask = min_count_mask
    if tot_mask is not None:
        tot_mask = torch.stack([tot_mask] * len(pos), dim=1)
    return res, tot_mask


class FullModel:
    """erosita_likelihood.py:73-83: the instrument with its kernel held as a (dynamic) leaf and an operator prepended:
    FullModel(kern, instrument, pre_ops)(x) = instrument(x=pre_ops(x), k=kern)."""

    def __init__(self, kern, instrument, pre_ops):
        self.kern, self.instrument, self.pre_ops = kern, instrument, pre_ops

    def __call__(self, x):
        return self.instrument(x=self.pre_ops(x), k=self.kern)


class PoissonLikelihood:
    """jft.Poissonian(masked_data).amend(R) on the fused CUDA path (erosita_likelihood.py:85-89,
    chandra_likelihood.py:37).

    `data` is either dense (T,E,N,N) counts or a Vector/dict of the reference's compacted 1-D arrays per key
    (response.py:110), which is scattered back onto the dense grid once (transpose of the mask).  With the eROSITA
    form of the response dictionary ('kernel' present) every call takes an optional kernel `k` -- the dynamic leaf of
    FullModel; the PSF spectra are rebuilt only when its content changed."""

    def __init__(self, response_dict, data, keys=None):
        torch = _torch()
        self.plan = response_dict['plan']
        self.kern = response_dict.get('kernel')
        if isinstance(data, (Vector, dict)):
            tree = data.tree if isinstance(data, Vector) else data
            keys = list(tree.keys()) if keys is None else keys
            data = response_dict['mask_adj'](Vector({k: tree[k] for k in keys}))[0] if 'mask_adj' in response_dict \
                else self._scatter(response_dict, tree, keys)
        self.plan.set_data(torch.as_tensor(data))

    def _scatter(self, response_dict, tree, keys):
        torch = _torch()
        p = self.plan
        probe = torch.ones((p.T, p.E, p.N, p.N), dtype=torch.float64, device=p.device)
        mask = response_dict['exposure'](probe) != 0
        dense = torch.zeros((p.T, p.E, p.N, p.N), dtype=torch.float64, device=p.device)
        for i, k in enumerate(keys):
            dense[i][mask[i]] = torch.as_tensor(tree[k]).to(p.device, torch.float64)
        return dense

    def _use(self, k):
        if k is not None:
            self.plan.ensure_kernel(k if getattr(k, "is_cuda", False) else _as_cuda(k, self.plan.device))

    def energy(self, x, k=None):
        self._use(k)
        return self.plan.energy(x)

    def value_and_grad(self, x, k=None):
        self._use(k)
        return self.plan.energy_and_grad(x)

    def metric(self, x, v, k=None):
        self._use(k)
        return self.plan.metric(x, v)

    def left_sqrt_metric(self, x, t, k=None):
        """R^T (t / sqrt(R x)) for a dense data-space tangent t ([S,] T, E, N, N); masked pixels are ignored."""
        self._use(k)
        return self.plan.left_sqrt_metric(x, t)

    def value_and_grad_logsky(self, log_sky, k=None):
        """value_and_grad with respect to log_sky, sky = exp(log_sky) (jubik/sky_models.py:340-341)."""
        self._use(k)
        return self.plan.energy_and_grad_logsky(log_sky)

    __call__ = energy
