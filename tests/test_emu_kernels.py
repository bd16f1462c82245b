"""CPU check of the CUDA kernels' logic: the SAME sources (jubik_b200/csrc) compiled with -DJB_EMU for host
threads (tests/emu/build_emu.py), driven through the same C ABI, against the golden vectors and the oracle.
This is test infrastructure for a box without a GPU; the product never loads the emulation library.
The GPU parity tests proper are tests/test_gpu_parity.py (-m gpu)."""
import os

import numpy as np
import pytest

from oracle import jubik_oracle as orc
from tests import synth
from tests.emu.emu_plan import EmuPlan

TOL = 1e-10
GOLDEN = ["lp_N32_n4_m4_full", "lp_N32_n4_m6_compact", "lp_N30_n4_m3_trunc", "lp_N24_n3_m2_full"]


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max())


@pytest.mark.parametrize("name", GOLDEN)
def test_emulated_kernels_match_reference_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    n, m = int(z["n"]), int(z["m"])
    p = EmuPlan(z["kernel"][None], z["x"].shape[-1], n, m)
    assert p.info["exact_period"] == 1
    assert rel(p.psf(z["x"])[0], z["y"]) < TOL
    assert rel(p.psf_T(z["g"][None]), z["yT"]) < TOL


@pytest.mark.parametrize("name", ["lp_N62_n4_m9_fold", "lp_N100_n5_m7_fold"])
def test_emulated_alias_fold_matches_reference_golden(golden_dir, name):
    """Period-P alias fold (P not a compiled length, PSF wider than the margin) against the reference source's output."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    x, n, m = z["x"], int(z["n"]), int(z["m"])
    E, N = x.shape[0], x.shape[-1]
    kern = synth.psf_full(1, n, E, N, seed=1, sig=tuple(z["sig"]))
    p = EmuPlan(kern, N, n, m)
    assert p.info["alias_fold"] == 1 and p.info["L"] == 96 and p.info["alias_ext"] == N // n - m
    assert rel(p.psf(x)[0], z["y"]) < TOL
    assert rel(p.psf_T(z["g"][None]), z["yT"]) < TOL


@pytest.mark.parametrize("case", [
    (64, 4, 6, 2, 2, 2, False, True),     # P=44 -> padded L=96 (two-pass rows, three-stage TMA-pipelined columns in
                                          # their emulated form), compact PSF, S/T/E batching
    (62, 4, 9, 1, 1, 1, True, False),     # N % n != 0, P = 48 is not compiled and the PSF is wider than the margin (half-support
                                          # 15 > 9): period-P alias fold on L = 160 >= w + 2*support = 60
    (64, 4, 3, 2, 1, 2, True, True),      # P = 38, full support 16 > 3: wide alias bands (ext = 13), T/S batching
])
def test_emulated_two_pass_fft_against_oracle(case):
    N, n, m, T, E, S, full, eff = case
    dx = N // n
    dom = orc.Domain((E, N, N), synth.DIST)
    if full:
        kern, cropped = synth.psf_full(T, n, E, N), False
    else:
        kern, cropped = synth.psf_cropped(T, n, E, dx, half_support=m), True
    expo = synth.exposure(T, E, N)
    ea = synth.eff_area(T, E) if eff else None
    x = synth.sky(S, E, N)
    R = orc.Response(dom, kern, n, m, expo, 500.0, synth.PIXEL_AREA, eff_area=ea, cropped=cropped)
    d = synth.counts_from(R.dense(synth.sky(1, E, N, seed=5)[0]))
    p = EmuPlan(kern, N, n, m, exposure=expo, exposure_cut=500.0, eff_area=ea, pixel_area=synth.PIXEL_AREA,
                data=d, max_samples=S, cropped=cropped, fft_length=160 if full else 96)   # (emulated threads are slow: the
    assert p.info["L"] == (160 if full else 96) and p.info["exact_period"] == 0            #  batched case takes the shorter length)
    assert p.info["alias_fold"] == (1 if full else 0)
    if full:
        assert p.info["alias_ext"] == dx - m and p.info["fast_path"] == 0
    g = synth.direction((S, T, E, N, N))
    assert rel(p.psf_T(g), np.stack([R.psf_T(g[s]) for s in range(S)])) < TOL
    assert rel(p.psf(x), np.stack([R.psf(x[s]) for s in range(S)])) < TOL
    assert rel(p.response_T(g), np.stack([R.dense_T(g[s]) for s in range(S)])) < TOL
    e, gr = p.energy_and_grad(x)
    eo, go = zip(*[orc.energy_and_grad(R, d, x[s]) for s in range(S)])
    assert np.abs((e - np.array(eo)) / np.array(eo)).max() < TOL
    assert rel(gr, np.stack(go)) < TOL
    v = synth.direction((S, E, N, N), seed=9)
    mo = np.stack([orc.metric_apply(R, x[s], v[s]) for s in range(S)])
    assert rel(p.metric(x, v), mo) < TOL
    # cached linearisation point: set_point once, then any number of applies
    p.set_point(x)
    assert rel(p.metric_at_point(v), mo) < TOL
    v2 = synth.direction((S, E, N, N), seed=10)
    assert rel(p.metric_at_point(v2), np.stack([orc.metric_apply(R, x[s], v2[s]) for s in range(S)])) < TOL
    assert rel(p.left_sqrt_metric(x, g), np.stack([orc.left_sqrt_metric(R, x[s], g[s]) for s in range(S)])) < TOL
    # sky-model tail fused into the row kernels (exp on load, * exp(log_sky) on store)
    ls = np.log(x)
    el, gl = p.energy_and_grad_logsky(ls)
    elo, glo = zip(*[orc.energy_and_grad_logsky(R, d, ls[s]) for s in range(S)])
    assert np.abs((el - np.array(elo)) / np.array(elo)).max() < TOL
    assert rel(gl, np.stack(glo)) < TOL
    eh, gh = p.energy_and_grad_host(x)
    assert np.array_equal(eh, e) and np.array_equal(gh, gr)
    # the sky model's padded log-sky: sky = exp(log_sky[:E, :N, :N]) read in place, gradient zero outside the crop
    Sx, Ex, Nx = x.shape[0], x.shape[1], x.shape[-1]
    lsp = np.random.default_rng(9).normal(size=(Sx, Ex + 1, Nx + 6, Nx + 6))
    lsp[:, :Ex, :Nx, :Nx] = np.log(x)
    ep_, gp_ = p.energy_and_grad_logsky(lsp)
    assert np.array_equal(ep_, el)
    assert np.array_equal(gp_[:, :Ex, :Nx, :Nx], gl)
    gp_[:, :Ex, :Nx, :Nx] = 0.0
    assert not gp_.any()


def test_emulated_sharded_entry_point_single_rank():
    """jubik_poisson_energy_and_grad_sharded without a communicator: layout of the flat all-reduce buffer,
    the sample fold kernel (fixed summation order) and the module-sharded in-place layout."""
    from jubik_b200 import _lib
    N, n, m, T, E, S = 64, 4, 6, 2, 2, 3
    dx = N // n
    kern = synth.psf_cropped(T, n, E, dx, half_support=m)
    expo = synth.exposure(T, E, N)
    x = synth.sky(S, E, N)
    R = orc.Response(orc.Domain((E, N, N), synth.DIST), kern, n, m, expo, 500.0, synth.PIXEL_AREA, cropped=True)
    d = synth.counts_from(R.dense(synth.sky(1, E, N, seed=5)[0]))
    p = EmuPlan(kern, N, n, m, exposure=expo, exposure_cut=500.0, pixel_area=synth.PIXEL_AREA, data=d,
                max_samples=S, cropped=True, fft_length=96)
    eo, go = zip(*[orc.energy_and_grad(R, d, x[s]) for s in range(S)])
    eo, go = np.array(eo), np.stack(go)
    npx = E * N * N
    for want_local in (True, False):
        flat, gl = p.energy_and_grad_sharded(x, _lib.SHARD_SAMPLES, want_local)
        assert flat.shape == (npx + S,)
        assert np.abs((flat[npx:] - eo) / eo).max() < TOL
        assert rel(flat[:npx].reshape(E, N, N), go.sum(0)) < TOL
        if want_local:
            assert rel(gl, go) < TOL
    flat, _ = p.energy_and_grad_sharded(x[:1], _lib.SHARD_SAMPLES)       # one local sample: gradient written in place
    assert rel(flat[:npx].reshape(E, N, N), go[0]) < TOL and abs(flat[npx] - eo[0]) / abs(eo[0]) < TOL
    flat, _ = p.energy_and_grad_sharded(x, _lib.SHARD_MODULES)
    assert flat.shape == (S * npx + S,)
    assert rel(flat[:S * npx].reshape(S, E, N, N), go) < TOL
    assert np.abs((flat[S * npx:] - eo) / eo).max() < TOL
    # more samples than the workspace holds: chunks of max_samples folded into flat one after the other
    p2 = EmuPlan(kern, N, n, m, exposure=expo, exposure_cut=500.0, pixel_area=synth.PIXEL_AREA, data=d,
                 max_samples=2, cropped=True, fft_length=96)
    for want_local in (True, False):
        flat, gl = p2.energy_and_grad_sharded(x, _lib.SHARD_SAMPLES, want_local)
        assert np.abs((flat[npx:] - eo) / eo).max() < TOL
        assert rel(flat[:npx].reshape(E, N, N), go.sum(0)) < TOL
        if want_local:
            assert rel(gl, go) < TOL
    flat, _ = p2.energy_and_grad_sharded(x, _lib.SHARD_MODULES)
    assert rel(flat[:S * npx].reshape(S, E, N, N), go) < TOL and np.abs((flat[S * npx:] - eo) / eo).max() < TOL


def test_emulated_standalone_fft_convolve_and_kernel_fingerprint():
    """jubik_convolve on the FFT kernels (rectangular, broadcasting branch) and jubik_plan_ensure_kernel (fingerprint hit
    / miss) through the emulation build -- index logic only, no GPU."""
    import ctypes as C
    from jubik_b200 import _lib
    from tests.emu.emu_plan import emu_lib, _p
    lib = emu_lib()
    rng = np.random.default_rng(3)
    for lead, x_lead, n0, n1 in ((2, 2, 96, 160), (3, 1, 160, 96), (1, 1, 24, 28)):
        x, y = rng.normal(size=(x_lead, n0, n1)), rng.normal(size=(lead, n0, n1))
        out = np.empty_like(y)
        _lib.check(lib, lib.jubik_convolve(_p(x), _p(y), lead, x_lead, n0, n1, 0.25, _p(out), None))
        ref = 0.25 * np.fft.ifft2(np.fft.fft2(x) * np.fft.fft2(y)).real
        assert rel(out, ref) < 1e-12
    N, n, m = 64, 4, 6
    k1 = synth.psf_cropped(1, n, 1, N // n, half_support=m)
    k2 = synth.psf_cropped(1, n, 1, N // n, half_support=m, seed=8)
    p = EmuPlan(k1, N, n, m, cropped=True, fft_length=160)
    flag = C.c_int(-1)
    k1c = k1.copy()                     # equal content at another address (kept alive across the call)
    p.raw.call("jubik_plan_ensure_kernel", _p(k1c), _lib.KERNEL_CROPPED, C.byref(flag), None)
    assert flag.value == 0
    p.raw.call("jubik_plan_ensure_kernel", _p(k2), _lib.KERNEL_CROPPED, C.byref(flag), None)
    assert flag.value == 1
    x = synth.sky(1, 1, N)
    dom = orc.Domain((1, N, N), synth.DIST)
    assert rel(p.psf(x)[0, 0], orc.linpatch_convolve(x[0], dom, k2[0], n, m, cropped=True)) < TOL


@pytest.mark.parametrize("name", ["psf_interp_N48_n4", "psf_interp_N40_n5"])
def test_emulated_psf_interpolator_matches_reference_golden(golden_dir, name):
    """jubik_psf_interpolate (k_psf_interp + its host-side weight / rotation set-up) through the emulation build against
    the reference source's output, in both layouts."""
    import ctypes as C
    from jubik_b200 import _lib
    from tests.emu.emu_plan import emu_lib, _p
    lib = emu_lib()
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    E, N, n, nrad, px, dist = int(z["E"]), int(z["N"]), int(z["n"]), int(z["nrad"]), int(z["px"]), float(z["dist"])
    infos = synth.psf_infos(E, nrad, px, seed=int(z["seed"]))
    tabs = np.ascontiguousarray(np.stack([i["psfs"] for i in infos]))
    rs = np.ascontiguousarray(np.stack([i["rs"] for i in infos]))
    cid = np.ascontiguousarray(np.stack([i["patch_center_ids"] for i in infos]))
    t = _lib.PsfTables(E, nrad, px, px, tabs.ctypes.data, rs.ctypes.data, cid.ctypes.data, 4.0, 4.0, *infos[0]["pointing_center"])
    full = np.empty((n * n, E, N, N))
    _lib.check(lib, lib.jubik_psf_interpolate(C.byref(t), N, n, dist, dist, _lib.KERNEL_FULL, _p(full), None))
    assert rel(full, z["kernel"]) < 1e-12
    dx = N // n
    crop = np.empty((n * n, E, 2 * dx, 2 * dx))
    _lib.check(lib, lib.jubik_psf_interpolate(C.byref(t), N, n, dist, dist, _lib.KERNEL_CROPPED, _p(crop), None))
    assert rel(crop, orc.crop_kernel(z["kernel"], (N, N), n)) < 1e-12
    assert lib.jubik_psf_interpolate(C.byref(t), N + 1, n, dist, dist, _lib.KERNEL_FULL, _p(full), None) == 1
