"""CPU oracle for the J-UBIK instrument-response hot path (TEST INFRASTRUCTURE ONLY).

This file is a NumPy/scipy float64 restatement of the reference algorithm.  It is the
checker for the CUDA path and the CPU baseline of ``bench.py``; it is never imported by
the product package ``jubik_b200`` (which fails loudly when its CUDA library is absent).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.

Reference lines followed (all under /root/reference):
  jubik/convolve.py:18-25    _bilinear_weights
  jubik/convolve.py:28-63    slice_patches              (pad, "xy" patch order, dynamic_slice clamp)
  jubik/convolve.py:66-145   linpatch_convolve          (window, kernel prep, circular conv, overlap-add)
  jubik/convolve.py:148-176  convolve                   (c2c fftn x2, ifftn, dvol, .real, broadcasting branch)
  jubik/convolve.py:179-184  integrate
  jubik/response.py:12-45    build_exposure_function    (in-place threshold, broadcast multiply)
  jubik/response.py:48-112   build_readout_function     (in-place threshold, mask, ragged compaction)
  jubik/instruments/erosita/erosita_response.py:209-212,307,327-328,349-350  vmap over modules, R
  jubik/instruments/chandra/chandra_response.py:127-142                     vmap over observations, R
  demos/x-ray-imaging.py:313  Poisson energy  sum(lambda) - d*log(lambda)   (arithmetic lives in NIFTy)
  jubik/instruments/erosita/erosita_psf_utils.py:17-264  psf_interpolator / get_psf (PSF tables -> patch kernels)

Pinning status
  * slice_patches / its transpose, convolve, _bilinear_weights, integrate, exposure and readout
    are pinned by the reference's own tests, ported in tests/test_oracle_reference_props.py,
    and by golden vectors produced by executing the reference's *own source files* under a
    NumPy-backed stand-in for jax (tests/golden/make_golden.py; jax is not installable here).
  * linpatch_convolve as a whole has no test in the reference; it is pinned only through those
    golden vectors (reference source, jax replaced by NumPy) -- no run of real jax exists.
  * jft.Poissonian is third-party NIFTy code that is absent from /root/reference
    (pyproject.toml:14 pins only ``nifty >= 9.0.0``): PARITY UNPINNED at that boundary; the
    published formula  H = sum(lambda) - d . log(lambda), metric 1/lambda  is restated.
"""
from __future__ import annotations

import os
from functools import reduce
from typing import NamedTuple

import numpy as np

try:  # scipy's pocketfft can use all host threads; same algorithm family as XLA-CPU's FFT
    import scipy.fft as _fft
    _HAVE_SCIPY = True
except Exception:  # pragma: no cover
    import numpy.fft as _fft
    _HAVE_SCIPY = False

_WORKERS = 1


def set_workers(n):
    """Number of host threads the FFTs may use (bench cpu_baseline uses os.cpu_count())."""
    global _WORKERS
    _WORKERS = max(1, int(n))


def _fftn(a, axes):
    if _HAVE_SCIPY:
        return _fft.fftn(a, axes=axes, workers=_WORKERS)
    return _fft.fftn(a, axes=axes)


def _ifftn(a, axes):
    if _HAVE_SCIPY:
        return _fft.ifftn(a, axes=axes, workers=_WORKERS)
    return _fft.ifftn(a, axes=axes)


class Domain(NamedTuple):
    """jubik/data.py:19-29 -- only shape[-2:], distances[-2:] are consumed."""
    shape: tuple
    distances: tuple


class Vector:
    """Minimal stand-in for nifty.re.Vector (a pytree wrapper around a dict)."""

    def __init__(self, tree):
        self.tree = tree

    def __getitem__(self, k):
        return self.tree[k]


# ----------------------------------------------------------------------------- convolve.py
def bilinear_weights(shape):
    """convolve.py:18-25. Uses shape[0] for both axes; odd size raises."""
    half = shape[0] / 2
    if half != int(half):
        raise ValueError("this should happen")
    ramp = np.linspace(0, 1, int(half), dtype="float64")
    b = np.concatenate([ramp, ramp[::-1]])
    return np.outer(b, b)


def patch_offsets(extent, n, dr):
    """Start indices (in the dx//2-padded frame) of the n windows along one axis.

    convolve.py:43-57.  jax.lax.dynamic_slice clamps start indices so the slice stays inside
    the operand; that only bites when dx is odd and the axis is exactly divisible.
    Returns (dx, window_size, offsets).
    """
    dx = int((extent - 2 * dr) / n)
    size = 2 * dx + 2 * dr
    frame = extent + 2 * (dx // 2)
    offs = np.minimum(np.arange(n) * dx, frame - size)
    offs = np.maximum(offs, 0)
    return dx, size, offs


def slice_patches(x, shape, n_patches_per_axis, additional_margin):
    """convolve.py:28-63.  (..., H, W) -> (n*n, ..., 2dx+2dr, 2dy+2dr).

    Patch k = i*n + j sits at row offset j*dx and column offset i*dy (meshgrid "xy" + flatten).
    """
    x = np.asarray(x)
    n = n_patches_per_axis
    dr = additional_margin
    dx, sx, offx = patch_offsets(shape[-2], n, dr)
    dy, sy, offy = patch_offsets(shape[-1], n, dr)
    pad = [(0, 0)] * (x.ndim - 2) + [(dx // 2, dx // 2), (dy // 2, dy // 2)]
    px = np.pad(x, pad, mode="constant", constant_values=0)
    out = np.empty((n * n,) + x.shape[:-2] + (sx, sy), dtype=px.dtype)
    for i in range(n):          # slow index of the flattened "xy" meshgrid
        for j in range(n):      # fast index
            out[i * n + j] = px[..., offx[j]:offx[j] + sx, offy[i]:offy[i] + sy]
    return out


def slice_patches_T(patches, shape, n_patches_per_axis, additional_margin, lead_shape=()):
    """Exact transpose of slice_patches (what jax.linear_transpose yields, convolve.py:141-143):
    overlap-add the windows into the padded frame, then crop the dx//2 pad."""
    patches = np.asarray(patches)
    n = n_patches_per_axis
    dr = additional_margin
    dx, sx, offx = patch_offsets(shape[-2], n, dr)
    dy, sy, offy = patch_offsets(shape[-1], n, dr)
    lead = tuple(patches.shape[1:-2])
    H, W = shape[-2], shape[-1]
    frame = np.zeros(lead + (H + 2 * (dx // 2), W + 2 * (dy // 2)), dtype=np.float64)
    for i in range(n):
        for j in range(n):
            frame[..., offx[j]:offx[j] + sx, offy[i]:offy[i] + sy] += patches[i * n + j]
    return frame[..., dx // 2:dx // 2 + H, dy // 2:dy // 2 + W]


def convolve(x, y, domain, axes):
    """convolve.py:148-176.  dvol * Re ifftn(fftn(x) * fftn(y)); broadcasting branch prints."""
    dlist = [domain.distances[i] for i in axes]
    dvol = float(reduce(lambda a, b: a * b, dlist))
    x = np.asarray(x)
    y = np.asarray(y)
    hx = _fftn(x, axes)
    hy = _fftn(y, axes)
    if y.ndim > x.ndim:
        print("kernel_shape:", x.shape)
        print("signal_shape:", y.shape)
        print("Dimension Inconsistency. Broadcasting PSFs")
        prod = hx[..., np.newaxis, :, :] * hy
    else:
        prod = hx * hy
    res = _ifftn(prod, axes)
    res = dvol * res
    return res.real


def integrate(x, domain, axes):
    """convolve.py:179-184."""
    s = np.sum(x, axis=tuple(axes))
    dlist = [domain.distances[i] for i in axes]
    dvol = float(reduce(lambda a, b: a * b, dlist))
    return dvol * s


def prepare_kernel(kernel, shape, distances, n, margin):
    """convolve.py:104-131: fftshift -> centre crop to (2dx, 2dy) -> zero-pad margin ->
    ifftshift -> divide by (sum * dvol).  kernel: (n*n, ..., H, W) origin-centred."""
    kernel = np.asarray(kernel, dtype=np.float64)
    dx = int(shape[-2] / n)
    dy = int(shape[-1] / n)
    kcx = (shape[-2] - 2 * dx) // 2
    kcy = (shape[-1] - 2 * dy) // 2
    rolled = np.fft.fftshift(kernel, axes=(-2, -1))
    cut = rolled[..., kcx:-kcx, kcy:-kcy]          # kc == 0 -> empty, as in the reference
    pad = [(0, 0)] * (cut.ndim - 2) + [(margin, margin)] * 2
    pk = np.pad(cut, pad, mode="constant", constant_values=0)
    back = np.fft.ifftshift(pk, axes=(-2, -1))
    summed = back.sum((-2, -1))
    dvol = distances[-2] * distances[-1]
    norm = (summed * np.array(dvol))[..., np.newaxis, np.newaxis]
    return back * norm ** -1


def crop_kernel(kernel, shape, n):
    """The centred (2dx, 2dy) crop that convolve.py:107-111 extracts (CROPPED layout helper)."""
    dx = int(shape[-2] / n)
    dy = int(shape[-1] / n)
    kcx = (shape[-2] - 2 * dx) // 2
    kcy = (shape[-1] - 2 * dy) // 2
    rolled = np.fft.fftshift(np.asarray(kernel), axes=(-2, -1))
    return rolled[..., kcx:-kcx, kcy:-kcy]


def prepare_cropped_kernel(cut, distances, margin):
    """Steps of convolve.py:114-131 applied to an already centred-cropped kernel."""
    cut = np.asarray(cut, dtype=np.float64)
    pad = [(0, 0)] * (cut.ndim - 2) + [(margin, margin)] * 2
    pk = np.pad(cut, pad, mode="constant", constant_values=0)
    back = np.fft.ifftshift(pk, axes=(-2, -1))
    summed = back.sum((-2, -1))
    dvol = distances[-2] * distances[-1]
    norm = (summed * np.array(dvol))[..., np.newaxis, np.newaxis]
    return back * norm ** -1


def linpatch_convolve(x, domain, kernel, n_patches_per_axis, margin, cropped=False):
    """convolve.py:66-145.  x: (..., H, W); kernel: (n*n, ..., H, W) -> (..., H, W).

    cropped=True feeds (n*n, ..., 2dx, 2dy) centred kernels (what :110-111 would extract)."""
    if not isinstance(domain, Domain):
        raise ValueError("domain has to be an instance of jubik.Domain")
    n = n_patches_per_axis
    shape = domain.shape
    spatial = [shape[-2], shape[-1]]
    slices = slice_patches(x, shape, n, 0)
    weights = bilinear_weights((slices.shape[-2], slices.shape[-1]))
    weighted = weights * slices
    pad = [(0, 0)] * (weighted.ndim - 2) + [(margin, margin)] * 2
    padded = np.pad(weighted, pad, mode="constant", constant_values=0)
    if cropped:
        normed = prepare_cropped_kernel(kernel, domain.distances, margin)
    else:
        normed = prepare_kernel(kernel, shape, domain.distances, n, margin)
    ndom = Domain((1, *shape), (None, *domain.distances))
    convolved = convolve(normed, padded, ndom, axes=(-2, -1))
    padded_shape = list(shape[:-2]) + [s + 2 * margin for s in spatial]
    padded_res = slice_patches_T(convolved, padded_shape, n, margin)
    return padded_res[..., margin:-margin, margin:-margin]


def linpatch_convolve_T(g, domain, kernel, n_patches_per_axis, margin, cropped=False):
    """Exact adjoint of x -> linpatch_convolve(x, ...) (what XLA's transpose computes):
    pad m -> slice_patches(margin m) -> dvol*Re ifft(conj(K^) fft(.)) -> crop m -> *weights
    -> overlap-add (transpose of slice_patches(margin 0))."""
    n = n_patches_per_axis
    shape = domain.shape
    g = np.asarray(g, dtype=np.float64)
    pad = [(0, 0)] * (g.ndim - 2) + [(margin, margin)] * 2
    gp = np.pad(g, pad, mode="constant", constant_values=0)
    padded_shape = list(shape[:-2]) + [shape[-2] + 2 * margin, shape[-1] + 2 * margin]
    patches = slice_patches(gp, padded_shape, n, margin)
    if cropped:
        normed = prepare_cropped_kernel(kernel, domain.distances, margin)
    else:
        normed = prepare_kernel(kernel, shape, domain.distances, n, margin)
    dvol = float(domain.distances[-2] * domain.distances[-1])
    hk = _fftn(normed, (-2, -1))
    hp = _fftn(patches, (-2, -1))
    corr = dvol * _ifftn(np.conj(hk) * hp, (-2, -1)).real
    inner = corr[..., margin:-margin, margin:-margin]
    weights = bilinear_weights((inner.shape[-2], inner.shape[-1]))
    return slice_patches_T(weights * inner, shape, n, 0)


# ----------------------------------------------------------------------------- response.py
def build_exposure_function(exposures, exposure_cut=None):
    """response.py:12-45 (mutates `exposures` in place like the reference)."""
    if exposure_cut is not None:
        if exposure_cut < 0:
            raise ValueError("exposure_cut should be non-negative or None!")
        exposures[exposures < exposure_cut] = 0

    def exposure(x):
        return exposures * x
    return exposure


def build_readout_function(flags, threshold=None, keys=None):
    """response.py:48-112 (mutates `flags` in place like the reference)."""
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
        if len(np.shape(x)) != 4:
            raise ValueError("input should have shape (n, m, q, l)!")
        return Vector({key: np.asarray(x)[i][~mask[i]] for i, key in enumerate(keys)})
    return _apply_readout


# ------------------------------------------------- instrument response + Poisson likelihood
class Response:
    """R(x) = mask(exposure(psf(x * pixel_area)))  for x:(E,N,N) -> dense (T,E,N,N) and the
    ragged readout (erosita_response.py:349-350, chandra_response.py:138-142).

    kernel: (T, n*n, E, N, N)  [or (T, n*n, E, 2dx, 2dy) with cropped=True]
    exposure: (T,E,N,N) raw; exp_cut thresholds it in place on a private copy
    eff_area: (T,E) or None (erosita_response.py:327-328)."""

    def __init__(self, domain, kernel, n, margin, exposure, exp_cut, pixel_area,
                 eff_area=None, keys=None, cropped=False):
        self.domain = domain
        self.kernel = np.asarray(kernel, dtype=np.float64)
        self.n = n
        self.margin = margin
        self.cropped = cropped
        self.pixel_area = float(pixel_area)
        expo = np.array(exposure, dtype=np.float64, copy=True)
        self._exposure_fn = build_exposure_function(expo, exp_cut)
        flags = np.array(exposure, dtype=np.float64, copy=True)
        self._readout = build_readout_function(flags, exp_cut, keys)
        self.exposure = expo                       # thresholded
        self.mask = expo == 0                      # True = not read out
        self.eff_area = None if eff_area is None else np.asarray(eff_area, dtype=np.float64)
        self.T = self.kernel.shape[0]

    def psf(self, x):
        """vmap(linpatch_convolve, in_axes=(None,None,0,None,None)) -> (T,E,N,N)."""
        return np.stack([linpatch_convolve(x, self.domain, self.kernel[t], self.n, self.margin,
                                           cropped=self.cropped) for t in range(self.T)])

    def psf_T(self, g):
        """(T,E,N,N) -> (E,N,N): sum over modules of the per-module adjoints."""
        return sum(linpatch_convolve_T(g[t], self.domain, self.kernel[t], self.n, self.margin,
                                       cropped=self.cropped) for t in range(self.T))

    def exposure_apply(self, y):
        out = self._exposure_fn(y)
        if self.eff_area is not None:
            out = out * self.eff_area[:, :, np.newaxis, np.newaxis]
        return out

    def dense(self, x):
        """Exposed signal on the dense (T,E,N,N) grid; masked entries set to 0."""
        lam = self.exposure_apply(self.psf(np.asarray(x) * self.pixel_area))
        return np.where(self.mask, 0.0, lam)

    def dense_T(self, g):
        """Adjoint of `dense`."""
        g = np.where(self.mask, 0.0, np.asarray(g, dtype=np.float64)) * self.exposure
        if self.eff_area is not None:
            g = g * self.eff_area[:, :, np.newaxis, np.newaxis]
        return self.psf_T(g) * self.pixel_area

    def __call__(self, x):
        lam = self.exposure_apply(self.psf(np.asarray(x) * self.pixel_area))
        return self._readout(lam)


def poisson_energy(lam_vec, data_vec):
    """jft.Poissonian(d).energy: sum(lambda) - vdot(d, log(lambda)) over every leaf."""
    h = 0.0
    for k in lam_vec.tree:
        lam = lam_vec.tree[k]
        d = data_vec.tree[k]
        h += np.sum(lam) - np.vdot(d, np.log(lam))
    return float(h)


def energy_and_grad(resp: Response, data_dense, x):
    """Fused value + VJP:  H(x) = sum_unmasked lambda - d log lambda,  lambda = R x."""
    x = np.asarray(x, dtype=np.float64)
    lam = resp.exposure_apply(resp.psf(x * resp.pixel_area))
    keep = ~resp.mask
    d = np.asarray(data_dense, dtype=np.float64)
    h = float(np.sum(lam[keep]) - np.vdot(d[keep], np.log(lam[keep])))
    with np.errstate(divide="ignore", invalid="ignore"):
        cot = np.where(keep, 1.0 - d / lam, 0.0)
    grad = resp.dense_T(cot)
    return h, grad


def metric_apply(resp: Response, x, v):
    """Poisson Fisher metric pulled back through R:  R^T diag(1/lambda) R v."""
    lam = resp.dense(x)
    rv = resp.dense(v)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(resp.mask, 0.0, rv / lam)
    return resp.dense_T(w)


def left_sqrt_metric(resp: Response, x, t):
    """Poissonian left_sqrt_metric t / sqrt(lambda) (NIFTy) pulled back through R."""
    lam = resp.dense(x)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(resp.mask, 0.0, np.asarray(t, dtype=np.float64) / np.sqrt(lam))
    return resp.dense_T(w)


def energy_and_grad_logsky(resp: Response, data_dense, log_sky):
    """value_and_grad with the sky-model tail sky = exp(log_sky) (jubik/sky_models.py:340-341)."""
    x = np.exp(np.asarray(log_sky, dtype=np.float64))
    h, g = energy_and_grad(resp, data_dense, x)
    return h, g * x


# ------------------------------------------------- eROSITA PSF interpolation onto the patch grid
def psf_interpolation_weights(rs, r):
    """erosita_psf_utils.py:41-72 get_interpolation_weights: linear weights of the (ascending) radii `rs` at `r`, constant
    extrapolation outside.  Restated with its overlap: at r == rs[-1] both the "back" branch and the last interval fire
    (the weight of the last table becomes 2), exactly as the reference's sum of cond() terms does."""
    rs = np.asarray(rs, dtype=np.float64)
    res = np.zeros(rs.shape)
    if r <= rs[0]:
        res[0] += 1.0
    if r >= rs[-1]:
        res[-1] += 1.0
    dr = rs[1:] - rs[:-1]
    for i in range(rs.size - 1):
        if rs[i] < r and rs[i + 1] >= r:
            wgt = (r - rs[i]) / dr[i]
            res[i] += 1.0 - wgt
            res[i + 1] += wgt
    return res


def psf_patch_centers(shape, distances, npatch):
    """erosita_psf_utils.py:213-246: (c_ra, c_dec) of the n*n patches; patch k = a*n + b has ra = cen1[n-1-a], dec = cen0[b]
    (meshgrid 'xy' + the flip of the ra list)."""
    cen = [np.array([(i * (ss // npatch) + (ss // npatch) / 2) * dd for i in range(npatch)]) for ss, dd in zip(shape, distances)]
    c0, c1 = np.meshgrid(cen[0], cen[1], indexing="xy")
    return np.flip(c1.flatten()), c0.flatten()


def psf_interpolator(domain, npatch, psf_infos):
    """erosita_psf_utils.py:198-264 psf_interpolator with :94-171 get_psf: for every patch centre the PSF tables of the
    two bracketing off-axis radii are rotated to the patch's position angle and bilinearly resampled
    (map_coordinates order=1, mode='nearest') onto the sky grid; result (n*n, E, N0, N1), origin-centred (rolled by N/2)."""
    from scipy import ndimage
    shp = (domain.shape[-2], domain.shape[-1])
    dist = (domain.distances[-2], domain.distances[-1])
    for ss in shp:
        if ss % npatch != 0 or ss % 2 != 0:
            raise ValueError
    c_p = np.meshgrid(*[(np.arange(ss) - ss / 2) * dd for ss, dd in zip(shp, dist)], indexing="ij")
    d_dec, d_ra = c_p[0], -1.0 * c_p[1]
    c_ra, c_dec = psf_patch_centers(shp, dist, npatch)
    out = []
    for info in psf_infos:
        rs = np.array(info["rs"], dtype=float)
        sort = np.argsort(rs)
        rs = rs[sort]
        psfs = np.asarray(info["psfs"], dtype=float)[sort]
        cids = np.asarray(info["patch_center_ids"], dtype=float)[sort]
        if psfs.shape[0] != rs.size or cids.shape[0] != rs.size:
            raise ValueError
        deltas = np.array(list(info["patch_deltas"]), dtype=float)
        pointing = np.array(list(info["pointing_center"]), dtype=float)
        patches = []
        for ra, dec in zip(c_ra, c_dec):
            cc = np.array([ra, dec]) - pointing
            r = np.sqrt(cc[0] ** 2 + cc[1] ** 2)
            phi = np.angle(cc[0] + 1.j * cc[1]) - np.pi / 2.
            wg = psf_interpolation_weights(rs, r)
            # rotate the requested slice to the patch's position angle (to_r_phi, phi -= phi_patch, to_ra_dec)
            ri = np.sqrt(d_ra ** 2 + d_dec ** 2)
            ph = np.angle(d_ra + 1.j * d_dec) - np.pi / 2. - phi
            xr = ri * np.cos(ph + np.pi / 2.)
            yr = ri * np.sin(ph + np.pi / 2.)
            res = np.zeros(shp)
            for ids, pp, ww in zip(cids, psfs, wg):
                q0 = (xr + ids[0] * deltas[0]) / deltas[0]
                q1 = (yr + ids[1] * deltas[1]) / deltas[1]
                res += ww * ndimage.map_coordinates(pp, np.stack([q0, q1]), order=1, mode="nearest")
            patches.append(np.roll(np.roll(res, -shp[0] // 2, axis=0), -shp[1] // 2, axis=1))
        out.append(np.array(patches))
    return np.moveaxis(np.stack(out), 0, -3)


def default_workers():
    return os.cpu_count() or 1
