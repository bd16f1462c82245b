"""Seeded synthetic inputs of SURVEY.md section 8(d) (NumPy default_rng / PCG64, float64).

Used by the tests (CPU and GPU) and by bench.py; sizes are parameters so the same generator serves the
small parity cases and the BASELINE.json configurations.
"""
import numpy as np

DIST = (1.0, 4.0, 4.0)          # domain distances -> dvol = 16
PIXEL_AREA = 16.0


def sky(S, E, N, seed=0):
    rng = np.random.default_rng(seed)
    return 5e-4 * np.exp(0.5 * rng.standard_normal((S, E, N, N)))


def psf_cropped(T, n, E, dx, half_support=None, seed=1, sig=(1.5, 4.5)):
    """Centred, un-normalised rotated anisotropic Gaussians in CROPPED layout (T, n*n, E, 2dx, 2dx);
    zero beyond `half_support` (compact) if given."""
    rng = np.random.default_rng(seed)
    w = 2 * dx
    d = np.arange(w) - dx
    yy, xx = np.meshgrid(d, d, indexing="ij")
    out = np.empty((T, n * n, E, w, w))
    par = rng.uniform(size=(T, n * n, E, 4))
    for t in range(T):
        for k in range(n * n):
            for e in range(E):
                s1 = sig[0] + (sig[1] - sig[0]) * par[t, k, e, 0]
                s2 = sig[0] + (sig[1] - sig[0]) * par[t, k, e, 1]
                th = np.pi * par[t, k, e, 2]
                amp = 0.5 + 1.5 * par[t, k, e, 3]
                u = np.cos(th) * yy + np.sin(th) * xx
                v = -np.sin(th) * yy + np.cos(th) * xx
                g = amp * np.exp(-0.5 * (u / s1) ** 2 - 0.5 * (v / s2) ** 2)
                if half_support is not None:
                    g[(np.abs(yy) > half_support) | (np.abs(xx) > half_support)] = 0.0
                out[t, k, e] = g
    return out


def psf_full(T, n, E, N, seed=1, sig=(1.5, 4.5)):
    """Origin-centred (wrapped) full-support kernels in the reference layout (T, n*n, E, N, N)."""
    rng = np.random.default_rng(seed)
    idx = np.fft.fftfreq(N, 1.0 / N)
    yy, xx = np.meshgrid(idx, idx, indexing="ij")
    out = np.empty((T, n * n, E, N, N))
    par = rng.uniform(size=(T, n * n, E, 4))
    for t in range(T):
        for k in range(n * n):
            for e in range(E):
                s1 = sig[0] + (sig[1] - sig[0]) * par[t, k, e, 0]
                s2 = sig[0] + (sig[1] - sig[0]) * par[t, k, e, 1]
                th = np.pi * par[t, k, e, 2]
                amp = 0.5 + 1.5 * par[t, k, e, 3]
                u = np.cos(th) * yy + np.sin(th) * xx
                v = -np.sin(th) * yy + np.cos(th) * xx
                out[t, k, e] = amp * np.exp(-0.5 * (u / s1) ** 2 - 0.5 * (v / s2) ** 2)
    return out


def exposure(T, E, N, seed=2):
    return np.random.default_rng(seed).uniform(0.0, 3e3, size=(T, E, N, N))


def eff_area(T, E, seed=6):
    return np.random.default_rng(seed).uniform(0.5, 1.5, size=(T, E))


def counts_from(lam_dense, seed=3):
    """Poisson draw of the (dense, masked = 0) expected counts, int64."""
    rng = np.random.default_rng(seed)
    return rng.poisson(np.maximum(lam_dense, 0.0)).astype(np.int64)


def direction(shape, seed=4):
    return np.random.default_rng(seed).standard_normal(shape)


def psf_infos(E, nrad, px, seed=5, pointing=(100.0, 150.0)):
    """Synthetic stand-in for eROSITA_PSF._get_psf_infos (erosita_psf.py:211-220), one dict per energy: PSF tables
    (nrad, px, px) = rotated Gaussians widening with the off-axis radius, radii given UNSORTED (the reference sorts them),
    centre indices jittered around the middle, 4-unit bins."""
    rng = np.random.default_rng(seed)
    infos = []
    yy, xx = np.meshgrid(np.arange(px), np.arange(px), indexing="ij")
    for e in range(E):
        rs = np.array([150.0, 0.0, 260.0, 60.0, 330.0][:nrad]) * (1.0 + 0.05 * e)
        cid = np.stack([np.array([px // 2 + rng.integers(-1, 2), px // 2 + rng.integers(-1, 2)])
                        for _ in range(nrad)]).astype(float)
        tabs = []
        for q in range(nrad):
            s1, s2 = 1.2 + rs[q] / 120.0 + 0.2 * e, 1.0 + rs[q] / 200.0
            th = rng.uniform(0, np.pi)
            u = np.cos(th) * (yy - cid[q, 0]) + np.sin(th) * (xx - cid[q, 1])
            v = -np.sin(th) * (yy - cid[q, 0]) + np.cos(th) * (xx - cid[q, 1])
            tabs.append(np.exp(-0.5 * (u / s1) ** 2 - 0.5 * (v / s2) ** 2) * rng.uniform(0.5, 2.0))
        infos.append({"psfs": np.array(tabs), "rs": rs, "patch_center_ids": cid, "patch_deltas": (4.0, 4.0),
                      "pointing_center": pointing})
    return infos
