"""Generate golden vectors by executing the REFERENCE'S OWN SOURCE FILES.

jax / jaxlib / nifty are not installable in this container (no network, no wheels), so the
reference cannot be imported as-is.  Its hot-path files, however, only use a handful of jax
entry points.  This script loads /root/reference/jubik/convolve.py and response.py verbatim
(nothing is copied into this repo) under a small NumPy-backed stand-in for those entry points:

    jax.numpy            -> numpy                    (pad, fft.fftn/ifftn/fftshift/ifftshift, ...)
    jax.lax.dynamic_slice-> NumPy slicing with jax's start-index clamping
    jax.vmap             -> Python loop + stack over axis 0
    jax.linear_transpose -> explicit matrix transpose: the linear function is applied to every
                            basis vector (small sizes only), so no hand-derived adjoint is used
    nifty.re.Vector      -> thin dict wrapper with .tree

Outputs tests/golden/*.npz (inputs + reference outputs).  Run in THIS container only:
    python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read only the committed .npz files.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/jubik"
OUT = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------ jax stand-in
def _dynamic_slice(operand, start_indices, slice_sizes):
    idx = []
    for s, sz, dim in zip(start_indices, slice_sizes, operand.shape):
        s = int(s)
        s = max(0, min(s, dim - sz))          # jax clamps so the slice stays in bounds
        idx.append(slice(s, s + sz))
    return operand[tuple(idx)]


def _vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        nmap = None
        for a, ax in zip(args, axes):
            if ax is not None:
                nmap = np.asarray(a).shape[ax]
        outs = []
        for b in range(nmap):
            call = [a if ax is None else np.take(np.asarray(a), b, axis=ax)
                    for a, ax in zip(args, axes)]
            outs.append(np.asarray(fun(*call)))
        return np.stack(outs, axis=out_axes)
    return mapped


_MATRIX_CACHE = {}


def _closure_key(fun, pshape):
    cells = ()
    if getattr(fun, "__closure__", None):
        cells = tuple(repr(c.cell_contents) for c in fun.__closure__
                      if isinstance(c.cell_contents, (int, float, list, tuple)))
    return (fun.__code__, cells, tuple(pshape))


def _linear_transpose(fun, primal, cache=True):
    """Transpose of a linear map by building its (sparse) matrix column by column: the
    function is applied to every basis vector, so no hand-derived adjoint is involved.
    The matrix is memoised per (code object, closure constants, primal shape)."""
    import scipy.sparse as sp
    pshape = np.shape(primal)
    psize = int(np.prod(pshape))
    key = _closure_key(fun, pshape) if cache else None
    mat = _MATRIX_CACHE.get(key) if cache else None
    if mat is None and psize > 20000:
        mat = _gather_matrix(fun, pshape)       # large primal: exact matrix of a pure gather from ONE probe call
        if mat is not None and cache:
            _MATRIX_CACHE[key] = mat
    if mat is None:
        cols = []
        e = np.zeros(psize)
        for k in range(psize):
            e[k] = 1.0
            cols.append(sp.csr_matrix(np.asarray(fun(e.reshape(pshape))).reshape(1, -1)))
            e[k] = 0.0
        mat = sp.vstack(cols).tocsr()          # row k = image of basis vector k  (= M^T)
        if cache:
            _MATRIX_CACHE[key] = mat

    def transposed(ct):
        ct = np.asarray(ct, dtype=np.float64).reshape(-1)
        return ((mat @ ct).reshape(pshape),)
    return transposed


def _gather_matrix(fun, pshape):
    """Matrix of a linear map that is a pure GATHER (every output element is one input element or a padded zero),
    e.g. the reference's slice_patches: apply the function once to the array of (flat index + 1); the output then
    names the source of every output element.  Verified on a random vector before use (exact equality), so nothing
    about the function is assumed: if it is not a gather the caller falls back to the basis-vector construction."""
    import scipy.sparse as sp
    psize = int(np.prod(pshape))
    probe = np.arange(1, psize + 1, dtype=np.float64).reshape(pshape)
    z = np.asarray(fun(probe)).reshape(-1)
    zi = np.rint(z).astype(np.int64)
    if not np.array_equal(zi.astype(np.float64), z) or zi.min() < 0 or zi.max() > psize:
        return None
    r = np.random.default_rng(123).normal(size=pshape)
    want = np.asarray(fun(r)).reshape(-1)
    got = np.where(zi > 0, r.reshape(-1)[np.maximum(zi - 1, 0)], 0.0)
    if not np.array_equal(want, got):
        return None
    rows = np.nonzero(zi > 0)[0]
    m = sp.csr_matrix((np.ones(rows.size), (rows, zi[rows] - 1)), shape=(z.size, psize))   # M: primal -> output
    return m.T.tocsr()                                                                     # row k = image of basis k


def _install_stand_ins():
    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if not name.startswith("__"):
            setattr(jnp, name, getattr(np, name))
    jnp.ndarray = np.ndarray
    lax = types.ModuleType("jax.lax")
    lax.dynamic_slice = _dynamic_slice
    jax.numpy = jnp
    jax.lax = lax
    jax.vmap = _vmap
    jax.linear_transpose = _linear_transpose
    sys.modules["jax"] = jax
    sys.modules["jax.numpy"] = jnp
    sys.modules["jax.lax"] = lax

    nifty = types.ModuleType("nifty")
    re_ = types.ModuleType("nifty.re")

    class Vector:
        def __init__(self, tree):
            self.tree = tree
    re_.Vector = Vector
    nifty.re = re_
    sys.modules["nifty"] = nifty
    sys.modules["nifty.re"] = re_

    # package shell so that `from .data import Domain` resolves without importing jubik/__init__
    pkg = types.ModuleType("jubik")
    pkg.__path__ = [REF]
    sys.modules["jubik"] = pkg
    data = types.ModuleType("jubik.data")
    from typing import NamedTuple

    class Domain(NamedTuple):      # jubik/data.py:19-29 is a bare NamedTuple(shape, distances)
        shape: tuple
        distances: tuple
    data.Domain = Domain
    sys.modules["jubik.data"] = data
    return Domain


def _load(name):
    spec = importlib.util.spec_from_file_location(f"jubik.{name}", os.path.join(REF, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[f"jubik.{name}"] = mod
    spec.loader.exec_module(mod)
    return mod


def gaussian_kernels(rng, npatch2, E, N, full_support=True, hs=None):
    """Origin-centred (wrapped) un-normalised rotated Gaussians, (npatch2, E, N, N)."""
    idx = np.fft.fftfreq(N, 1.0 / N)
    yy, xx = np.meshgrid(idx, idx, indexing="ij")
    out = np.empty((npatch2, E, N, N))
    for k in range(npatch2):
        for e in range(E):
            s1, s2 = rng.uniform(0.8, 2.5, size=2)
            th = rng.uniform(0, np.pi)
            amp = rng.uniform(0.5, 2.0)
            u = np.cos(th) * yy + np.sin(th) * xx
            v = -np.sin(th) * yy + np.cos(th) * xx
            g = amp * np.exp(-0.5 * (u / s1) ** 2 - 0.5 * (v / s2) ** 2)
            if not full_support:
                g[(np.abs(yy) > hs) | (np.abs(xx) > hs)] = 0.0
            out[k, e] = g
    return out


def main():
    Domain = _install_stand_ins()
    conv = _load("convolve")
    resp = _load("response")

    cases = {
        # name: (E, N, n, m, full_support)
        "lp_N32_n4_m4_full": (2, 32, 4, 4, True),       # wrap-around of period P is exercised
        "lp_N32_n4_m6_compact": (1, 32, 4, 6, False),   # support <= m: equals linear convolution
        "lp_N30_n4_m3_trunc": (1, 30, 4, 3, True),      # N % n != 0 -> int() truncation (A2)
        "lp_N24_n3_m2_full": (1, 24, 3, 2, True),       # smallest legal n (A4)
    }
    for name, (E, N, n, m, full) in cases.items():
        rng = np.random.default_rng(sum(map(ord, name)))
        x = 5e-4 * np.exp(0.5 * rng.normal(size=(E, N, N)))
        kern = gaussian_kernels(rng, n * n, E, N, full_support=full, hs=m)
        dom = Domain((E, N, N), (1.0, 4.0, 4.0))
        y = np.asarray(conv.linpatch_convolve(x, dom, kern, n, m))
        # transpose of the whole chain, again by explicit matrix (reference code only)
        g = rng.normal(size=(E, N, N))
        yT = _linear_transpose(lambda a: conv.linpatch_convolve(a, dom, kern, n, m), x,
                               cache=False)(g)[0]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x, kernel=kern, n=n, m=m,
                            distances=np.array(dom.distances), y=y, g=g, yT=yT)
        print(name, y.shape, float(np.abs(y).max()))

    # ---- the PRODUCTION pipelines pinned to the reference source directly: full-size geometries whose P is
    # a compiled three-stage length, and small geometries that need the period-P alias fold.  Kernels come from
    # tests/synth.py seeds (regenerated by the tests), so only x / y (/ g, yT) are stored.
    sys.path.insert(0, os.path.abspath(os.path.join(OUT, "..", "..")))
    from tests import synth
    big = {
        # name: (N, n, m, sig)                      P    what it pins
        "lp_N256_n4_m16_wide": (256, 4, 16, (3.0, 12.0)),   # 160  BASELINE configs[0] (C1), PSF tails wrap around the period
        "lp_N512_n4_m32_wide": (512, 4, 32, (4.0, 20.0)),   # 320  Chandra demo: sdim 512, npatch 4, margfrac 0.062 (chandra_demo.yaml:13,44-45)
        "lp_N512_n8_m32_wide": (512, 8, 32, (4.0, 20.0)),   # 192  eROSITA demo: sdim 512, npatch 8 (eROSITA_demo.yaml:35-36,48)
    }
    for name, (N, n, m, sig) in big.items():
        kern = synth.psf_full(1, n, 1, N, seed=1, sig=sig)[0]
        x = synth.sky(1, 1, N, seed=0)[0]
        dom = Domain((1, N, N), synth.DIST)
        y = np.asarray(conv.linpatch_convolve(x, dom, kern, n, m))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), N=N, n=n, m=m, sig=np.array(sig), y=y,
                            distances=np.array(dom.distances))
        print(name, y.shape, float(np.abs(y).max()))
    fold = {
        # name: (E, N, n, m, sig): P is NOT a compiled length and the PSF half-support (dx) exceeds the margin
        "lp_N62_n4_m9_fold": (1, 62, 4, 9, (2.0, 8.0)),     # P = 48 -> L = 96, ext = 6
        "lp_N100_n5_m7_fold": (2, 100, 5, 7, (2.0, 10.0)),  # P = 54 -> L = 96, ext = 13
    }
    for name, (E, N, n, m, sig) in fold.items():
        rng = np.random.default_rng(sum(map(ord, name)))
        kern = synth.psf_full(1, n, E, N, seed=1, sig=sig)[0]
        x = 5e-4 * np.exp(0.5 * rng.normal(size=(E, N, N)))
        dom = Domain((E, N, N), synth.DIST)
        y = np.asarray(conv.linpatch_convolve(x, dom, kern, n, m))
        g = rng.normal(size=(E, N, N))
        yT = _linear_transpose(lambda a: conv.linpatch_convolve(a, dom, kern, n, m), x, cache=False)(g)[0]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x, n=n, m=m, sig=np.array(sig), y=y, g=g, yT=yT,
                            distances=np.array(dom.distances))
        print(name, y.shape, float(np.abs(y).max()))

    # slice_patches and its jax.linear_transpose, the fixture of test/test_slice_patches.py
    rng = np.random.default_rng(7)
    shape = (36, 36)
    a = np.arange(36 * 36, dtype=float).reshape(shape)
    sp = np.asarray(conv.slice_patches(a, shape, 16, 2))
    b = rng.random(sp.shape)
    spT = _linear_transpose(lambda f: conv.slice_patches(f, shape, 16, 2), a, cache=False)(b)[0]
    # odd dx with exact divisibility: dynamic_slice clamping path
    a2 = rng.normal(size=(15, 15))
    sp2 = np.asarray(conv.slice_patches(a2, (15, 15), 5, 0))
    np.savez_compressed(os.path.join(OUT, "slice_patches.npz"), a=a, sp=sp, b=b, spT=spT,
                        a2=a2, sp2=sp2)

    # convolve(): plain and broadcasting branch
    rng = np.random.default_rng(11)
    sig = rng.normal(size=(3, 12, 10))
    ker = rng.normal(size=(3, 12, 10))
    dom = Domain((3, 12, 10), (1.0, 0.5, 0.25))
    c1 = np.asarray(conv.convolve(ker, sig, dom, axes=(-2, -1)))
    ker2 = rng.normal(size=(12, 10))
    c2 = np.asarray(conv.convolve(ker2, sig, dom, axes=(-2, -1)))
    w = conv._bilinear_weights((8, 8))
    integ = conv.integrate(sig, dom, axes=(-2, -1))
    np.savez_compressed(os.path.join(OUT, "convolve.npz"), sig=sig, ker=ker, c1=c1, ker2=ker2,
                        c2=c2, w8=w, integ=integ, distances=np.array(dom.distances))

    # exposure + readout
    rng = np.random.default_rng(0)
    expo = rng.uniform(0.0, 3e3, size=(2, 3, 24, 24))
    x4 = rng.normal(size=(2, 3, 24, 24))
    ex = resp.build_exposure_function(expo.copy(), 500.0)(x4)
    ro = resp.build_readout_function(expo.copy(), 500.0, ("tm_1", "tm_2"))(x4)
    np.savez_compressed(os.path.join(OUT, "response.npz"), expo=expo, x=x4, exposed=ex,
                        ro_tm_1=ro.tree["tm_1"], ro_tm_2=ro.tree["tm_2"])
    print("done")


if __name__ == "__main__":
    main()
