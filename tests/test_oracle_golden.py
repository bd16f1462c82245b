"""Pin the NumPy oracle: (1) against golden vectors produced by executing the reference's OWN source files
(tests/golden/make_golden.py), (2) against the properties the reference's own tests assert
(test/test_slice_patches.py, test/test_convolve.py, test/test_build_exposure_function.py,
test/test_build_exposure_readout_function.py), (3) adjointness and finite differences of the fused path."""
import os

import numpy as np
import pytest

from oracle import jubik_oracle as orc
from tests import synth

GOLDEN = ["lp_N32_n4_m4_full", "lp_N32_n4_m6_compact", "lp_N30_n4_m3_trunc", "lp_N24_n3_m2_full"]


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max())


@pytest.mark.parametrize("name", GOLDEN)
def test_linpatch_matches_reference_source(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    dom = orc.Domain(tuple(z["x"].shape), tuple(z["distances"]))
    n, m = int(z["n"]), int(z["m"])
    assert rel(orc.linpatch_convolve(z["x"], dom, z["kernel"], n, m), z["y"]) < 1e-13
    assert rel(orc.linpatch_convolve_T(z["g"], dom, z["kernel"], n, m), z["yT"]) < 1e-13
    # CROPPED layout helper is the same operator
    cut = orc.crop_kernel(z["kernel"], dom.shape, n)
    assert rel(orc.linpatch_convolve(z["x"], dom, cut, n, m, cropped=True), z["y"]) < 1e-13


WIDE = ["lp_N256_n4_m16_wide", "lp_N512_n4_m32_wide", "lp_N512_n8_m32_wide"]
FOLD = ["lp_N62_n4_m9_fold", "lp_N100_n5_m7_fold"]


@pytest.mark.parametrize("name", WIDE)
def test_linpatch_production_geometries_match_reference_source(golden_dir, name):
    """C1 and the reference's two shipped demo geometries (512^2, 4x4 / 8x8, margfrac 0.062) with PSFs whose tails
    wrap around the period: forward output of the reference's own convolve.py; kernels regenerated from the seed."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    N, n, m = int(z["N"]), int(z["n"]), int(z["m"])
    kern = synth.psf_full(1, n, 1, N, seed=1, sig=tuple(z["sig"]))[0]
    x = synth.sky(1, 1, N, seed=0)[0]
    dom = orc.Domain((1, N, N), tuple(z["distances"]))
    assert rel(orc.linpatch_convolve(x, dom, kern, n, m), z["y"]) < 1e-13
    cut = orc.crop_kernel(kern, dom.shape, n)
    assert rel(orc.linpatch_convolve(x, dom, cut, n, m, cropped=True), z["y"]) < 1e-13


@pytest.mark.parametrize("name", FOLD)
def test_linpatch_wide_psf_on_non_smooth_period_matches_reference_source(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    x, n, m = z["x"], int(z["n"]), int(z["m"])
    E, N = x.shape[0], x.shape[-1]
    kern = synth.psf_full(1, n, E, N, seed=1, sig=tuple(z["sig"]))[0]
    dom = orc.Domain(tuple(x.shape), tuple(z["distances"]))
    assert rel(orc.linpatch_convolve(x, dom, kern, n, m), z["y"]) < 1e-13
    assert rel(orc.linpatch_convolve_T(z["g"], dom, kern, n, m), z["yT"]) < 1e-13


def test_slice_patches_matches_reference_source(golden_dir):
    z = np.load(os.path.join(golden_dir, "slice_patches.npz"))
    sp = orc.slice_patches(z["a"], (36, 36), 16, 2)
    assert sp.shape == z["sp"].shape and np.array_equal(sp, z["sp"])
    assert np.allclose(orc.slice_patches_T(z["b"], (36, 36), 16, 2), z["spT"], rtol=0, atol=1e-13)
    assert np.array_equal(orc.slice_patches(z["a2"], (15, 15), 5, 0), z["sp2"])  # dynamic_slice clamping


def test_convolve_weights_integrate_match_reference_source(golden_dir):
    z = np.load(os.path.join(golden_dir, "convolve.npz"))
    dom = orc.Domain((3, 12, 10), tuple(z["distances"]))
    assert rel(orc.convolve(z["ker"], z["sig"], dom, (-2, -1)), z["c1"]) < 1e-13
    assert rel(orc.convolve(z["ker2"], z["sig"], dom, (-2, -1)), z["c2"]) < 1e-13
    assert np.array_equal(orc.bilinear_weights((8, 8)), z["w8"])
    assert np.allclose(orc.integrate(z["sig"], dom, (-2, -1)), z["integ"], rtol=1e-14)


def test_exposure_readout_match_reference_source(golden_dir):
    z = np.load(os.path.join(golden_dir, "response.npz"))
    assert np.array_equal(orc.build_exposure_function(z["expo"].copy(), 500.0)(z["x"]), z["exposed"])
    ro = orc.build_readout_function(z["expo"].copy(), 500.0, ("tm_1", "tm_2"))(z["x"])
    assert np.array_equal(ro.tree["tm_1"], z["ro_tm_1"]) and np.array_equal(ro.tree["tm_2"], z["ro_tm_2"])


# ---- properties asserted by the reference's own tests ------------------------------------------
def test_slice_patches_shape_linearity_adjointness():
    # test/test_slice_patches.py:10-19,39-66
    shape, n, ov = (36, 36), 16, 2
    rng = np.random.default_rng(42)
    a, b = rng.normal(size=shape), rng.normal(size=shape)
    fa = orc.slice_patches(a, shape, n, ov)
    dx = int((36 - 2 * ov) / n)
    assert fa.shape == (n * n, 2 * dx + 2 * ov, 2 * dx + 2 * ov)
    assert np.allclose(orc.slice_patches(2 * a + 3 * b, shape, n, ov), 2 * fa + 3 * orc.slice_patches(b, shape, n, ov))
    c = rng.normal(size=fa.shape)
    assert np.isclose(np.vdot(c, fa), np.vdot(orc.slice_patches_T(c, shape, n, ov), a), rtol=1e-12)


def test_bilinear_weights_properties_and_odd_raises():
    # test/test_convolve.py:8-20
    w = orc.bilinear_weights((8, 8))
    assert w.shape == (8, 8) and (w >= 0).all() and np.allclose(w, w.T)
    with pytest.raises(ValueError):
        orc.bilinear_weights((7, 7))


def test_integrate_known_answer_and_delta_identity(capsys):
    # test/test_convolve.py:23-55
    dom = orc.Domain((3, 4), (0.5, 0.25))
    assert np.isclose(orc.integrate(np.ones((3, 4)), dom, (0, 1)), 3 * 4 * 0.5 * 0.25)
    rng = np.random.default_rng(0)
    sig = rng.normal(size=(8, 8))
    delta = np.zeros((8, 8))
    delta[0, 0] = 1.0
    dom = orc.Domain((8, 8), (1.0, 1.0))
    assert np.allclose(orc.convolve(delta, sig, dom, (0, 1)), sig, rtol=1e-10, atol=1e-10)
    out = orc.convolve(delta, rng.normal(size=(2, 8, 8)), orc.Domain((2, 8, 8), (1.0, 1.0, 1.0)), (-2, -1))
    assert out.shape == (2, 8, 8) and "Broadcasting" in capsys.readouterr().out


def test_exposure_and_readout_reference_behaviour():
    # test/test_build_exposure_function.py:29-52, test/test_build_exposure_readout_function.py:48-108
    rng = np.random.default_rng(1)
    expo = rng.uniform(0, 3e3, size=(2, 2, 6, 6))
    x = rng.normal(size=expo.shape)
    cut = expo.copy()
    cut[cut < 500] = 0
    assert np.array_equal(orc.build_exposure_function(expo.copy(), 500)(x), cut * x)
    assert np.array_equal(orc.build_exposure_function(expo.copy(), None)(x), expo * x)
    with pytest.raises(ValueError):
        orc.build_exposure_function(expo.copy(), -1)
    ro = orc.build_readout_function(expo.copy(), 500, ["a", "b"])(x)
    for i, k in enumerate(["a", "b"]):
        assert np.array_equal(ro.tree[k], x[i][cut[i] != 0])
    with pytest.raises(ValueError):
        orc.build_readout_function(expo.copy(), 500, ["a"])
    with pytest.raises(ValueError):
        orc.build_readout_function(expo.copy(), -2, ["a", "b"])
    with pytest.raises(ValueError):
        orc.build_readout_function(expo.copy(), 500, ["a", "b"])(x[0])
    assert list(orc.build_readout_function(expo[:1].copy(), None)(x[:1]).tree) == ["masked input"]
    assert orc.build_readout_function(expo[:1].copy(), None)(x[:1]).tree["masked input"].size == x[:1].size
    with pytest.raises(ValueError):
        orc.linpatch_convolve(x[0], ((2, 6, 6), (1, 1, 1)), x, 3, 1)


# ---- oracle self-consistency on the synthetic inputs --------------------------------------------
def _small_response():
    N, n, m, T, E = 48, 4, 4, 2, 2
    dom = orc.Domain((E, N, N), synth.DIST)
    kern = synth.psf_full(T, n, E, N, sig=(0.8, 2.0))
    R = orc.Response(dom, kern, n, m, synth.exposure(T, E, N), 500.0, synth.PIXEL_AREA, eff_area=synth.eff_area(T, E),
                     keys=("tm1", "tm2"))
    x = synth.sky(1, E, N)[0]
    d = synth.counts_from(R.dense(synth.sky(1, E, N, seed=5)[0]))
    return R, x, d


def test_response_adjointness_and_fd_gradient():
    R, x, d = _small_response()
    g = synth.direction((R.T,) + x.shape)
    assert np.isclose(np.vdot(g, R.dense(x)), np.vdot(R.dense_T(g), x), rtol=1e-12)
    assert np.isclose(np.vdot(g, R.psf(x)), np.vdot(R.psf_T(g), x), rtol=1e-12)
    e0, grad = orc.energy_and_grad(R, d, x)
    v = synth.direction(x.shape, seed=9) * x
    eps = 1e-4
    fd = (orc.energy_and_grad(R, d, x + eps * v)[0] - orc.energy_and_grad(R, d, x - eps * v)[0]) / (2 * eps)
    assert np.isclose(fd, np.vdot(grad, v), rtol=1e-6)
    lam = R(x)
    dv = orc.Vector({k: d[i][~R.mask[i]] for i, k in enumerate(lam.tree)})
    assert np.isclose(orc.poisson_energy(lam, dv), e0, rtol=1e-13)
    mv = orc.metric_apply(R, x, v)
    assert np.vdot(v, mv) > 0  # Fisher metric is positive


def test_partition_of_unity_interior():
    # with a delta PSF the response is x * coverage, coverage == 1 in the interior
    N, n, m = 64, 4, 4
    kern = np.zeros((n * n, 1, N, N))
    kern[..., 0, 0] = 2.5
    x = synth.sky(1, 1, N)[0]
    y = orc.linpatch_convolve(x, orc.Domain((1, N, N), synth.DIST), kern, n, m)
    dx = N // n
    inner = slice(dx // 2 + 1, N - dx // 2 - 1)
    assert np.abs(y[:, inner, inner] / x[:, inner, inner] - 1).max() < 1e-13


@pytest.mark.parametrize("name", ["psf_interp_N48_n4", "psf_interp_N40_n5"])
def test_oracle_psf_interpolator_matches_reference_source(golden_dir, name):
    """erosita_psf_utils.py:198-264 psf_interpolator executed from the reference's own source (make_golden_psf.py) on
    synthetic PSF tables: the oracle's restatement reproduces it bit for bit (unsorted radii, extrapolation at both ends
    of the radius range, rotation to the patch's position angle, 'nearest' edge handling)."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    E, N, n = int(z["E"]), int(z["N"]), int(z["n"])
    infos = synth.psf_infos(E, int(z["nrad"]), int(z["px"]), seed=int(z["seed"]))
    dist = float(z["dist"])
    k = orc.psf_interpolator(orc.Domain((E, N, N), (1.0, dist, dist)), n, infos)
    assert k.shape == z["kernel"].shape and np.array_equal(k, z["kernel"])
    with pytest.raises(ValueError):
        orc.psf_interpolator(orc.Domain((E, N + 1, N + 1), (1.0, dist, dist)), n, infos)
    w = orc.psf_interpolation_weights(np.array([0.0, 1.0, 3.0]), 2.0)
    assert np.allclose(w, [0.0, 0.5, 0.5]) and np.allclose(orc.psf_interpolation_weights(np.array([0.0, 1.0]), -1.0), [1, 0])
