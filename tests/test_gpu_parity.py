"""GPU parity tests: the CUDA path (through the C ABI, via jubik_b200.ResponsePlan) against
(1) the golden vectors generated from the reference's own sources (tests/golden/make_golden.py) and
(2) the NumPy oracle on the same seeded inputs.  Tolerance: 1e-10 relative (BASELINE.json north_star),
measured as max|a-b| / max|b| per array and relative error on the scalar energy.
"""
import os

import numpy as np
import pytest

from tests import synth
from oracle import jubik_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def dev(torch, a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


GOLDEN = ["lp_N32_n4_m4_full", "lp_N32_n4_m6_compact", "lp_N30_n4_m3_trunc", "lp_N24_n3_m2_full"]


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_linpatch_forward_and_transpose(torch_cuda, golden_dir, name):
    import jubik_b200 as jb
    torch = torch_cuda
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    x, kern, n, m = z["x"], z["kernel"], int(z["n"]), int(z["m"])
    N = x.shape[-1]
    plan = jb.ResponsePlan(dev(torch, kern[None]), N, n, m)
    y = plan.psf(dev(torch, x))[0].cpu().numpy()
    assert rel(y, z["y"]) < TOL
    yT = plan.psf_T(dev(torch, z["g"][None])).cpu().numpy()
    assert rel(yT, z["yT"]) < TOL
    # the reference-named entry point
    dom = jb.Domain(tuple(x.shape), tuple(z["distances"]))
    y2 = jb.linpatch_convolve(dev(torch, x), dom, dev(torch, kern), n, m).cpu().numpy()
    assert rel(y2, z["y"]) < TOL
    plan.close()


WIDE = ["lp_N256_n4_m16_wide", "lp_N512_n4_m32_wide", "lp_N512_n8_m32_wide"]


@pytest.mark.parametrize("name", WIDE)
def test_golden_production_pipeline_vs_reference_source(torch_cuda, golden_dir, name):
    """The PRODUCTION kernels (TMA column pipeline, warp-private c2r rows, prefetching r2c rows; P = 160 / 320 / 192 are
    three-stage lengths) against forward outputs of the reference's own convolve.py -- BASELINE configs[0] and the two
    shipped demo geometries (512^2, npatch 4 / 8, margfrac 0.062: chandra_demo.yaml:13,44-45, eROSITA_demo.yaml:35-36,48)
    with PSFs whose tails wrap around the period (full-support kernels in the reference layout)."""
    import jubik_b200 as jb
    torch = torch_cuda
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    N, n, m = int(z["N"]), int(z["n"]), int(z["m"])
    kern = synth.psf_full(1, n, 1, N, seed=1, sig=tuple(z["sig"]))
    x = synth.sky(1, 1, N, seed=0)[0]
    plan = jb.ResponsePlan(dev(torch, kern), N, n, m)
    assert plan.info["exact_period"] == 1 and plan.info["fast_path"] == 1 and plan.info["psf_half_support"] > m
    assert rel(plan.psf(dev(torch, x))[0].cpu().numpy(), z["y"]) < TOL
    # the transpose through the same kernels is the exact adjoint of that forward
    g = dev(torch, synth.direction((1, 1, 1, N, N)))
    xd = dev(torch, x[None])
    lhs = torch.vdot(g.flatten(), plan.psf(xd).flatten())
    rhs = torch.vdot(plan.psf_T(g).flatten(), xd.flatten())
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-12
    plan.close()


@pytest.mark.parametrize("name", ["lp_N62_n4_m9_fold", "lp_N100_n5_m7_fold"])
def test_golden_alias_fold_vs_reference_source(torch_cuda, golden_dir, name):
    """P is not a compiled length and the PSF is wider than the margin: linear convolution on L >= w + 2*support folded
    modulo P, forward and transpose against the reference source's outputs."""
    import jubik_b200 as jb
    torch = torch_cuda
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    x, n, m = z["x"], int(z["n"]), int(z["m"])
    E, N = x.shape[0], x.shape[-1]
    kern = synth.psf_full(1, n, E, N, seed=1, sig=tuple(z["sig"]))
    plan = jb.ResponsePlan(dev(torch, kern), N, n, m)
    assert plan.info["alias_fold"] == 1 and plan.info["exact_period"] == 0 and plan.info["alias_ext"] == N // n - m
    assert rel(plan.psf(dev(torch, x))[0].cpu().numpy(), z["y"]) < TOL
    assert rel(plan.psf_T(dev(torch, z["g"][None])).cpu().numpy(), z["yT"]) < TOL
    plan.close()


# (N, n, m, T, E, S, full_support, eff_area)  -> FFT length exercised
CASES = [
    (256, 4, 16, 1, 1, 1, True, False),    # C1: P=160 exact period, full-support PSF wraps around
    (240, 4, 10, 2, 2, 2, False, True),    # P=140 -> L=160 padded transform, compact PSF, S/T/E batching
    (250, 4, 18, 1, 2, 1, True, False),    # N % n != 0 truncation, P=160
    (408, 4, 64, 1, 1, 1, False, False),   # C3 patch geometry: P=332 -> L=336
    (512, 4, 64, 1, 2, 1, True, False),    # C2 patch geometry: P=384 exact, full support
    (768, 6, 64, 2, 1, 1, False, True),    # P=384, more patches, two modules
    (1024, 4, 64, 1, 1, 1, False, False),  # C4 patch geometry: P=640
    (816, 4, 128, 1, 1, 1, False, False),  # C5 patch geometry: P=664 -> L=672
    (512, 8, 32, 2, 2, 2, True, True),     # eROSITA demo geometry: P=192 exact, full support, S/T/E batching
    (512, 4, 32, 1, 3, 1, True, False),    # Chandra demo geometry: P=320 exact, full support, 3 bins
    (408, 4, 64, 1, 2, 1, True, False),    # C3 patch geometry with a FULL-support PSF: P=332, alias fold on L=448
    (1024, 8, 126, 1, 1, 1, False, False), # P=508 -> L=512 padded; (2048/16 patch geometry, odd demo margin)
    (260, 4, 15, 1, 2, 1, True, False),    # ODD margin, P=160 exact: TMA boxes land on odd tile rows (r2c memcheck finding)
    (512, 4, 31, 2, 1, 2, False, True),    # odd margin, P=318 -> L=320 padded (ceil(0.062*2048) = 127 is odd too)
]


def build_case(N, n, m, T, E, S, full, eff):
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
    return dom, kern, cropped, expo, ea, x, R, d


@pytest.mark.parametrize("case", CASES, ids=lambda c: "N%d_n%d_m%d_T%d_E%d_S%d_%s" % (c[:6] + ("full" if c[6] else "compact",)))
def test_oracle_parity_all_entry_points(torch_cuda, case):
    import jubik_b200 as jb
    torch = torch_cuda
    N, n, m, T, E, S, full, eff = case
    dom, kern, cropped, expo, ea, x, R, d = build_case(*case)
    plan = jb.ResponsePlan(dev(torch, kern), N, n, m, exposure=dev(torch, expo), exposure_cut=500.0,
                           eff_area=None if ea is None else dev(torch, ea), pixel_area=synth.PIXEL_AREA,
                           data=dev(torch, d), max_samples=S, cropped=cropped, host_call=True)
    if full and plan.info["P"] != plan.info["L"]:
        assert plan.info["alias_fold"] == 1 and plan.info["L"] >= plan.info["w"] + 2 * plan.info["psf_half_support"]
    xd = dev(torch, x)
    g = synth.direction((S, T, E, N, N))
    gd = dev(torch, g)
    assert rel(plan.psf(xd).cpu().numpy(), np.stack([R.psf(x[s]) for s in range(S)])) < TOL
    assert rel(plan.psf_T(gd).cpu().numpy(), np.stack([R.psf_T(g[s]) for s in range(S)])) < TOL
    assert rel(plan.response(xd).cpu().numpy(), np.stack([R.dense(x[s]) for s in range(S)])) < TOL
    assert rel(plan.response_T(gd).cpu().numpy(), np.stack([R.dense_T(g[s]) for s in range(S)])) < TOL
    e, gr = plan.energy_and_grad(xd)
    eo, go = zip(*[orc.energy_and_grad(R, d, x[s]) for s in range(S)])
    assert np.abs((e.cpu().numpy() - np.array(eo)) / np.array(eo)).max() < TOL
    assert rel(gr.cpu().numpy(), np.stack(go)) < TOL
    v = synth.direction((S, E, N, N), seed=9)
    mo = np.stack([orc.metric_apply(R, x[s], v[s]) for s in range(S)])
    assert rel(plan.metric(xd, dev(torch, v)).cpu().numpy(), mo) < TOL
    # left-sqrt-metric and the log-sky tail
    lo = np.stack([orc.left_sqrt_metric(R, x[s], g[s]) for s in range(S)])
    assert rel(plan.left_sqrt_metric(xd, gd).cpu().numpy(), lo) < TOL
    # cached linearisation point (what NIFTy's CG does: one position, many applies)
    plan.set_point(xd)
    v2 = synth.direction((S, E, N, N), seed=10)
    assert rel(plan.metric_at_point(dev(torch, v)).cpu().numpy(), mo) < TOL
    assert rel(plan.metric_at_point(dev(torch, v2)).cpu().numpy(),
               np.stack([orc.metric_apply(R, x[s], v2[s]) for s in range(S)])) < TOL
    assert rel(plan.left_sqrt_metric_at_point(gd).cpu().numpy(), lo) < TOL
    ls = np.log(x)
    # the sky model's padded array (sky_models.py:340: exp(log_diffuse(x)[:edim, :sdim, :sdim])), cropped in place
    lsp = np.random.default_rng(9).normal(size=(S, E + 1, N + 8, N + 8))
    lsp[:, :E, :N, :N] = ls
    elp, glp = plan.energy_and_grad_logsky(dev(torch, lsp))
    el, gl = plan.energy_and_grad_logsky(dev(torch, ls))
    assert torch.equal(elp, el) and torch.equal(glp[:, :E, :N, :N], gl)
    glp[:, :E, :N, :N] = 0.0
    assert not bool(glp.any())
    elo, glo = zip(*[orc.energy_and_grad_logsky(R, d, ls[s]) for s in range(S)])
    assert np.abs((el.cpu().numpy() - np.array(elo)) / np.array(elo)).max() < TOL
    assert rel(gl.cpu().numpy(), np.stack(glo)) < TOL
    # host-buffer end-to-end entry point gives the same numbers
    xh = torch.as_tensor(x).pin_memory()
    eh = torch.empty((S,), dtype=torch.float64).pin_memory()
    gh = torch.empty_like(xh).pin_memory()
    plan.energy_and_grad_host(xh, eh, gh)
    assert torch.equal(eh, e.cpu()) and torch.equal(gh, gr.cpu())
    plan.close()


def test_readout_and_likelihood_wrappers_match_reference_layout(torch_cuda):
    """response_dict['R'] returns the reference's ragged {key: 1-D} tree (response.py:110) and the
    likelihood accepts the compacted data tree (data.pkl layout)."""
    import jubik_b200 as jb
    torch = torch_cuda
    N, n, m, T, E = 240, 4, 10, 2, 2
    dom_o, kern, cropped, expo, ea, x, R, d = build_case(N, n, m, T, E, 1, False, True)
    dom = jb.Domain(dom_o.shape, dom_o.distances)
    rd = jb.build_response(dev(torch, kern), dev(torch, expo), dom, n, m, synth.PIXEL_AREA, exposure_cut=500.0,
                           effective_area=dev(torch, ea), keys=["tm1", "tm2"], cropped=True)
    out = rd["R"](dev(torch, x[0]))
    ref = R(x[0])
    for k_ref, k in zip(ref.tree, ["tm1", "tm2"]):
        a, b = out[k].cpu().numpy(), ref.tree[k_ref]
        assert a.shape == b.shape and rel(a, b) < TOL
    # compacted data tree -> dense scatter -> same energy as the dense oracle
    mask = R.mask
    tree = {k: torch.as_tensor(d[i][~mask[i]]) for i, k in enumerate(["tm1", "tm2"])}
    lh = jb.PoissonLikelihood(rd, tree)
    e, gr = lh.value_and_grad(dev(torch, x[0]))
    eo, go = orc.energy_and_grad(R, d, x[0])
    assert abs(float(e) - eo) / abs(eo) < TOL and rel(gr.cpu().numpy(), go) < TOL


def test_delta_kernel_is_window_weighted_identity(torch_cuda):
    """Analogue of test/test_convolve.py:32-41 (delta kernel => identity) for linpatch_convolve: a delta PSF
    gives x times the window coverage, which is exactly 1 away from the dx/2 border (Appendix A3)."""
    import jubik_b200 as jb
    torch = torch_cuda
    N, n, m = 256, 4, 16
    dx = N // n
    kern = np.zeros((1, n * n, 1, N, N))
    kern[..., 0, 0] = 3.0  # un-normalised delta
    x = synth.sky(1, 1, N)[0]
    y = jb.ResponsePlan(dev(torch, kern), N, n, m).psf(dev(torch, x))[0].cpu().numpy()
    inner = slice(dx // 2 + 1, N - dx // 2 - 1)
    assert np.abs(y[:, inner, inner] / x[:, inner, inner] - 1).max() < 1e-12


def test_full_size_chandra_adjointness_linearity_and_fd(torch_cuda):
    """BASELINE.json configs[1] at full size (1024^2, 3 bins, 8x8 patches, margin 64): size-independent
    properties -- dot-product adjointness (test_slice_patches.py:55-66 style), linearity (:48-53), and a
    finite-difference check of the fused energy gradient."""
    import jubik_b200 as jb
    torch = torch_cuda
    N, n, m, T, E, S = 1024, 8, 64, 1, 3, 1
    dx = N // n
    kern = dev(torch, synth.psf_cropped(T, n, E, dx, half_support=m))
    expo = dev(torch, synth.exposure(T, E, N))
    plan = jb.ResponsePlan(kern, N, n, m, exposure=expo, exposure_cut=500.0, pixel_area=synth.PIXEL_AREA,
                           cropped=True, max_samples=S)
    x = dev(torch, synth.sky(S, E, N))
    g = dev(torch, synth.direction((S, T, E, N, N)))
    lhs = torch.vdot(g.flatten(), plan.response(x).flatten())
    rhs = torch.vdot(plan.response_T(g).flatten(), x.flatten())
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-12
    lhs = torch.vdot(g.flatten(), plan.psf(x).flatten())
    rhs = torch.vdot(plan.psf_T(g).flatten(), x.flatten())
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-12
    x2 = dev(torch, synth.sky(S, E, N, seed=11))
    lin = plan.psf(2.0 * x + 3.0 * x2) - (2.0 * plan.psf(x) + 3.0 * plan.psf(x2))
    assert float(lin.abs().max()) / float(plan.psf(x).abs().max()) < 1e-12
    # Poisson data from the model itself, then FD of the energy along a random direction
    lam = plan.response(dev(torch, synth.sky(1, E, N, seed=5)))
    torch.manual_seed(1234)
    d = torch.poisson(lam.clamp(min=0.0))[0]
    plan.set_data(d)
    e0, gr = plan.energy_and_grad(x)
    v = dev(torch, synth.direction((S, E, N, N), seed=9)) * x  # relative perturbation keeps x > 0
    eps = 1e-4
    ep = plan.energy(x + eps * v)
    em = plan.energy(x - eps * v)
    fd = float((ep - em).sum()) / (2 * eps)
    an = float(torch.vdot(gr.flatten(), v.flatten()))
    # the directional derivative along a random direction can be small by chance (E[lambda - d] = 0 near the
    # truth), so the tolerance is relative to |g||v| as well as to |g.v|
    scale = max(abs(an), 1e-4 * float(gr.norm() * v.norm()))
    assert abs(fd - an) / scale < 1e-6
    # determinism: bitwise identical on repetition
    e1, gr1 = plan.energy_and_grad(x)
    assert torch.equal(e0, e1) and torch.equal(gr, gr1)
    plan.close()


def test_full_size_erosita_like_c3_oracle_parity_and_adjointness(torch_cuda):
    """BASELINE.json configs[2] at full size (7 modules x 3 bins, 1024^2, 10x10 patches, margin 64: N % n != 0,
    P = 332 -> L = 336 padded transform, per-module effective areas): fused energy/gradient against the oracle
    for every module and bin, plus dot-product adjointness of the full response."""
    import jubik_b200 as jb
    torch = torch_cuda
    N, n, m, T, E, S = 1024, 10, 64, 7, 3, 1
    dom, kern, cropped, expo, ea, x, R, d = build_case(N, n, m, T, E, S, False, True)
    plan = jb.ResponsePlan(dev(torch, kern), N, n, m, exposure=dev(torch, expo), exposure_cut=500.0,
                           eff_area=dev(torch, ea), pixel_area=synth.PIXEL_AREA, data=dev(torch, d),
                           max_samples=S, cropped=True)
    assert plan.info["L"] == 336 and plan.info["P"] == 332
    xd = dev(torch, x)
    e, gr = plan.energy_and_grad(xd)
    eo, go = orc.energy_and_grad(R, d, x[0])
    assert abs(float(e[0]) - eo) / abs(eo) < TOL
    assert rel(gr[0].cpu().numpy(), go) < TOL
    g = dev(torch, synth.direction((S, T, E, N, N)))
    lhs = torch.vdot(g.flatten(), plan.response(xd).flatten())
    rhs = torch.vdot(plan.response_T(g).flatten(), xd.flatten())
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-12
    plan.close()


def test_full_size_jwst_like_c4_oracle_parity_and_adjointness(torch_cuda):
    """BASELINE.json configs[3] at full size (4096^2 sky, 16x16 patches, 128-px kernels = margin 64, P = 640):
    forward PSF against the oracle, adjointness and the fused energy gradient's directional derivative."""
    import jubik_b200 as jb
    torch = torch_cuda
    N, n, m, T, E, S = 4096, 16, 64, 1, 1, 1
    dx = N // n
    kern = synth.psf_cropped(T, n, E, dx, half_support=m)
    expo = synth.exposure(T, E, N)
    x = synth.sky(S, E, N)
    plan = jb.ResponsePlan(dev(torch, kern), N, n, m, exposure=dev(torch, expo), exposure_cut=500.0,
                           pixel_area=synth.PIXEL_AREA, max_samples=S, cropped=True)
    assert plan.info["L"] == 640 and plan.info["exact_period"] == 1
    xd = dev(torch, x)
    y = plan.psf(xd)
    yo = orc.linpatch_convolve(x[0], orc.Domain((E, N, N), synth.DIST), kern[0], n, m, cropped=True)
    assert rel(y[0, 0].cpu().numpy(), yo) < TOL
    g = dev(torch, synth.direction((S, T, E, N, N)))
    lhs = torch.vdot(g.flatten(), plan.response(xd).flatten())
    rhs = torch.vdot(plan.response_T(g).flatten(), xd.flatten())
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-12
    lam = plan.response(dev(torch, synth.sky(1, E, N, seed=5)))
    torch.manual_seed(7)
    plan.set_data(torch.poisson(lam.clamp(min=0.0))[0])
    e0, gr = plan.energy_and_grad(xd)
    v = dev(torch, synth.direction((S, E, N, N), seed=9)) * xd
    eps = 1e-4
    fd = float((plan.energy(xd + eps * v) - plan.energy(xd - eps * v)).sum()) / (2 * eps)
    an = float(torch.vdot(gr.flatten(), v.flatten()))
    assert abs(fd - an) / max(abs(an), 1e-4 * float(gr.norm() * v.norm())) < 1e-6
    plan.close()


def test_full_size_geovi_batch_c5_per_gpu_share(torch_cuda):
    """BASELINE.json configs[4], the share of ONE GPU (2 of the 16 posterior samples x 7 modules x 3 bins at 2048^2,
    10x10 patches, margin 128: P = 664 -> L = 672): sample batching is bitwise the same as one sample at a time,
    adjointness of the batched response, and a finite-difference check of the fused gradient."""
    import jubik_b200 as jb
    torch = torch_cuda
    N, n, m, T, E, S = 2048, 10, 128, 7, 3, 2
    dx = N // n
    kern = synth.psf_cropped(T, n, E, dx, half_support=m)
    expo = synth.exposure(T, E, N)
    ea = synth.eff_area(T, E)
    plan = jb.ResponsePlan(dev(torch, kern), N, n, m, exposure=dev(torch, expo), exposure_cut=500.0,
                           eff_area=dev(torch, ea), pixel_area=synth.PIXEL_AREA, max_samples=S, cropped=True)
    del kern
    assert plan.info["L"] == 672 and plan.info["P"] == 664
    x = dev(torch, synth.sky(S, E, N))
    lam = plan.response(dev(torch, synth.sky(1, E, N, seed=5)))
    torch.manual_seed(11)
    plan.set_data(torch.poisson(lam.clamp(min=0.0))[0])
    del lam
    e, gr = plan.energy_and_grad(x)
    for s in range(S):
        es, gs = plan.energy_and_grad(x[s:s + 1])
        assert torch.equal(es[0], e[s]) and torch.equal(gs[0], gr[s])
    g = dev(torch, synth.direction((S, T, E, N, N)))
    lhs = torch.vdot(g.flatten(), plan.response(x).flatten())
    rhs = torch.vdot(plan.response_T(g).flatten(), x.flatten())
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-12
    del g
    v = dev(torch, synth.direction((S, E, N, N), seed=9)) * x
    eps = 1e-4
    fd = float((plan.energy(x + eps * v) - plan.energy(x - eps * v)).sum()) / (2 * eps)
    an = float(torch.vdot(gr.flatten(), v.flatten()))
    assert abs(fd - an) / max(abs(an), 1e-4 * float(gr.norm() * v.norm())) < 1e-6
    plan.close()


def test_host_buffer_entry_point_is_bitwise_equal_to_device_call(torch_cuda):
    """jubik_poisson_energy_and_grad_host: (sample, bin) planes alternating between two streams with two workspaces give
    exactly the batched device call's numbers (same kernels, same summation order), call after call; and the entry
    point refuses to run on a plan that was created without its staging (no allocation after plan creation)."""
    import jubik_b200 as jb
    torch = torch_cuda
    case = (512, 4, 64, 2, 3, 2, False, True)
    N, n, m, T, E, S, full, eff = case
    dom, kern, cropped, expo, ea, x, R, d = build_case(*case)

    def mk(**kw):
        return jb.ResponsePlan(dev(torch, kern), N, n, m, exposure=dev(torch, expo), exposure_cut=500.0,
                               eff_area=dev(torch, ea), pixel_area=synth.PIXEL_AREA, data=dev(torch, d), max_samples=S,
                               cropped=cropped, **kw)
    xd = dev(torch, x)
    p0 = mk(host_call=True)
    e0, g0 = p0.energy_and_grad(xd)
    xh = torch.as_tensor(x).pin_memory()
    eh = torch.empty((S,), dtype=torch.float64).pin_memory()
    gh = torch.empty_like(xh).pin_memory()
    for _ in range(3):
        gh.zero_()
        p0.energy_and_grad_host(xh, eh, gh)
        assert torch.equal(eh, e0.cpu()) and torch.equal(gh, g0.cpu())
    with pytest.raises((TypeError, ValueError)):            # undersized / wrong-device buffers are rejected in Python
        p0.energy_and_grad_host(xh, eh[:1], gh)
    with pytest.raises((TypeError, ValueError)):
        p0.energy_and_grad_host(xd, eh, gh)
    bytes_before = p0.info["bytes_total"]
    p1 = mk()
    assert p1.info["bytes_total"] < bytes_before
    with pytest.raises(jb.JubikError, match="JUBIK_PLAN_HOST_CALL"):
        p1.energy_and_grad_host(xh, eh, gh)
    p0.close()
    p1.close()


def test_set_kernel_and_linpatch_cache_never_confuse_two_psfs(torch_cuda):
    """The eROSITA call form has the kernel as a DYNAMIC argument (erosita_response.py:349-350): jubik_plan_set_kernel
    swaps the PSF of a live plan, and linpatch_convolve called with two different NumPy kernels of identical shape
    (fresh device temporaries that may reuse one address -- ADVICE r1) returns two different, correct results."""
    import jubik_b200 as jb
    torch = torch_cuda
    N, n, m, E = 256, 4, 16, 1
    dom_o = orc.Domain((E, N, N), synth.DIST)
    dom = jb.Domain(dom_o.shape, dom_o.distances)
    k1 = synth.psf_full(1, n, E, N, seed=1)[0]
    k2 = synth.psf_full(1, n, E, N, seed=77, sig=(2.0, 9.0))[0]
    x = synth.sky(1, E, N)[0]
    y1 = orc.linpatch_convolve(x, dom_o, k1, n, m)
    y2 = orc.linpatch_convolve(x, dom_o, k2, n, m)
    assert rel(y1, y2) > 1e-3
    for _ in range(2):
        assert rel(jb.linpatch_convolve(x, dom, k1, n, m).cpu().numpy(), y1) < TOL
        assert rel(jb.linpatch_convolve(x, dom, k2, n, m).cpu().numpy(), y2) < TOL
    kd = dev(torch, k1)
    assert rel(jb.linpatch_convolve(dev(torch, x), dom, kd, n, m).cpu().numpy(), y1) < TOL
    kd.copy_(dev(torch, k2))          # same tensor object, new content
    assert rel(jb.linpatch_convolve(dev(torch, x), dom, kd, n, m).cpu().numpy(), y2) < TOL
    plan = jb.ResponsePlan(dev(torch, k1[None]), N, n, m)
    assert rel(plan.psf(dev(torch, x))[0].cpu().numpy(), y1) < TOL
    plan.set_kernel(dev(torch, k2[None]))
    assert rel(plan.psf(dev(torch, x))[0].cpu().numpy(), y2) < TOL
    # padded length: a new PSF that no longer fits the plan's transform length is refused, not approximated
    N2, n2, m2 = 240, 4, 10       # P = 140 -> L = 160 with a compact PSF
    kc = synth.psf_cropped(1, n2, 1, N2 // n2, half_support=m2)
    p2 = jb.ResponsePlan(dev(torch, kc), N2, n2, m2, cropped=True)
    assert p2.info["L"] == 160 and p2.info["alias_fold"] == 0
    with pytest.raises(jb.JubikError, match="create a new plan"):
        p2.set_kernel(dev(torch, synth.psf_cropped(1, n2, 1, N2 // n2, half_support=None, sig=(4.0, 20.0))))
    plan.close()
    p2.close()


def test_exact_benchmark_config_c2_two_samples_oracle_parity(torch_cuda):
    """The configuration bench.py --workload C2 times (1024^2, 8x8 patches, margin 64, E=3, TWO samples -- the n <= 8 form
    of the c2r kernels) against the ORACLE: energy, gradient, psf and response_T (this exact config had
    only self-consistency checks)."""
    import jubik_b200 as jb
    torch = torch_cuda
    case = (1024, 8, 64, 1, 3, 2, False, False)
    N, n, m, T, E, S, full, eff = case
    dom, kern, cropped, expo, ea, x, R, d = build_case(*case)
    plan = jb.ResponsePlan(dev(torch, kern), N, n, m, exposure=dev(torch, expo), exposure_cut=500.0,
                           pixel_area=synth.PIXEL_AREA, data=dev(torch, d), max_samples=S, cropped=True)
    assert plan.info["L"] == 384 and plan.info["fast_path"] == 1
    xd = dev(torch, x)
    e, gr = plan.energy_and_grad(xd)
    eo, go = zip(*[orc.energy_and_grad(R, d, x[s]) for s in range(S)])
    assert np.abs((e.cpu().numpy() - np.array(eo)) / np.array(eo)).max() < TOL
    assert rel(gr.cpu().numpy(), np.stack(go)) < TOL
    assert rel(plan.psf(xd).cpu().numpy(), np.stack([R.psf(x[s]) for s in range(S)])) < TOL
    g = synth.direction((S, T, E, N, N))
    assert rel(plan.response_T(dev(torch, g)).cpu().numpy(), np.stack([R.dense_T(g[s]) for s in range(S)])) < TOL
    plan.close()


def test_c5_plane_and_c4_energy_gradient_oracle_parity(torch_cuda):
    """One (module, bin) plane of the C5 share (2048^2, 10x10 patches, margin 128: P = 664 -> L = 672, the WIDE c2r form)
    and C4 (4096^2, 16x16, P = 640): fused energy + gradient against the oracle at full image size."""
    import jubik_b200 as jb
    torch = torch_cuda
    for case, L in (((2048, 10, 128, 1, 1, 1, False, True), 672), ((4096, 16, 64, 1, 1, 1, False, False), 640)):
        N, n, m, T, E, S, full, eff = case
        dom, kern, cropped, expo, ea, x, R, d = build_case(*case)
        plan = jb.ResponsePlan(dev(torch, kern), N, n, m, exposure=dev(torch, expo), exposure_cut=500.0,
                               eff_area=None if ea is None else dev(torch, ea), pixel_area=synth.PIXEL_AREA,
                               data=dev(torch, d), max_samples=S, cropped=True)
        assert plan.info["L"] == L and plan.info["fast_path"] == 1
        e, gr = plan.energy_and_grad(dev(torch, x))
        eo, go = orc.energy_and_grad(R, d, x[0])
        assert abs(float(e[0]) - eo) / abs(eo) < TOL
        assert rel(gr[0].cpu().numpy(), go) < TOL
        plan.close()
        del kern, expo, x, R, d


def test_full_size_c3_with_full_support_psf_alias_fold(torch_cuda):
    """BASELINE configs[2] geometry (1024^2, 10x10 patches, margin 64: P = 332 = 4*83 is not a smooth length) with a PSF
    that fills the whole 2dx crop (half-support 102 > 64): the reference's period-332 wrap-around is reproduced by the
    alias fold on L = 448.  Two modules, one bin (the oracle's cost is per plane)."""
    import jubik_b200 as jb
    torch = torch_cuda
    N, n, m, T, E, S = 1024, 10, 64, 2, 1, 1
    dx = N // n
    kern = synth.psf_cropped(T, n, E, dx, half_support=None, sig=(4.0, 25.0))
    expo = synth.exposure(T, E, N)
    ea = synth.eff_area(T, E)
    x = synth.sky(S, E, N)
    R = orc.Response(orc.Domain((E, N, N), synth.DIST), kern, n, m, expo, 500.0, synth.PIXEL_AREA, eff_area=ea, cropped=True)
    d = synth.counts_from(R.dense(synth.sky(1, E, N, seed=5)[0]))
    plan = jb.ResponsePlan(dev(torch, kern), N, n, m, exposure=dev(torch, expo), exposure_cut=500.0, eff_area=dev(torch, ea),
                           pixel_area=synth.PIXEL_AREA, data=dev(torch, d), max_samples=S, cropped=True)
    assert plan.info["P"] == 332 and plan.info["alias_fold"] == 1 and plan.info["L"] == 448
    xd = dev(torch, x)
    assert rel(plan.psf(xd)[0].cpu().numpy(), R.psf(x[0])) < TOL
    e, gr = plan.energy_and_grad(xd)
    eo, go = orc.energy_and_grad(R, d, x[0])
    assert abs(float(e[0]) - eo) / abs(eo) < TOL and rel(gr[0].cpu().numpy(), go) < TOL
    # the padded transform WITHOUT the fold is what the reference does not compute: it must differ visibly
    compact = synth.psf_cropped(T, n, E, dx, half_support=m, sig=(4.0, 25.0))
    Rc = orc.Response(orc.Domain((E, N, N), synth.DIST), compact, n, m, expo, 500.0, synth.PIXEL_AREA, eff_area=ea, cropped=True)
    assert rel(Rc.psf(x[0]), R.psf(x[0])) > 1e-3
    plan.close()


def test_edge_cases_of_survey_appendix_a(torch_cuda):
    """A5: an all-zero kernel makes the reference divide by zero (NaN output); the CUDA path mirrors that (NaN over at
    least the same footprint, no crash).
    A12: lambda <= 0 on an unmasked pixel gives a NaN / inf energy in the reference (log of a non-positive number);
    mirrored: a non-positive sky yields a non-finite energy instead of a silently wrong number."""
    import jubik_b200 as jb
    torch = torch_cuda
    N, n, m = 256, 4, 16
    kern = synth.psf_full(1, n, 1, N)
    kern[0, 5] = 0.0                                     # one patch with an all-zero PSF
    y = jb.ResponsePlan(dev(torch, kern), N, n, m).psf(dev(torch, synth.sky(1, 1, N)[0]))
    assert bool(torch.isnan(y).any())
    yo = orc.linpatch_convolve(synth.sky(1, 1, N)[0], orc.Domain((1, N, N), synth.DIST), kern[0], n, m)
    # the reference's NaN footprint (the P x P window of that patch) is NaN here too; it may extend over the neighbouring
    # column block (two blocks share one packed c2r transform) -- everywhere else the numbers agree
    yg = y[0].cpu().numpy()
    assert np.all(np.isnan(yg)[np.isnan(yo)])
    both = ~np.isnan(yg)
    assert both.sum() > 0.4 * both.size and np.abs(yg[both] - yo[both]).max() / np.abs(yo[both]).max() < TOL
    case = (256, 4, 16, 1, 1, 1, True, False)
    dom, kern, cropped, expo, ea, x, R, d = build_case(*case)
    plan = jb.ResponsePlan(dev(torch, kern), N, n, m, exposure=dev(torch, expo), exposure_cut=500.0,
                           pixel_area=synth.PIXEL_AREA, data=dev(torch, d))
    e = plan.energy(dev(torch, -x))
    assert not bool(torch.isfinite(e).all())


def test_standalone_slice_patches_and_convolve_match_reference_golden(torch_cuda, golden_dir, capsys):
    """The stand-alone public helpers against outputs of the reference's own source (tests/golden/make_golden.py):
    slice_patches (test/test_slice_patches.py fixture, incl. dynamic_slice clamping), its linear transpose, and
    convolve with its broadcasting branch (test/test_convolve.py)."""
    import jubik_b200 as jb
    torch = torch_cuda
    z = np.load(os.path.join(golden_dir, "slice_patches.npz"))
    sp = jb.slice_patches(dev(torch, z["a"]), (36, 36), 16, 2).cpu().numpy()
    assert sp.shape == z["sp"].shape and np.array_equal(sp, z["sp"])
    spT = jb.slice_patches_T(dev(torch, z["b"]), (36, 36), 16, 2).cpu().numpy()
    assert np.allclose(spT, z["spT"], rtol=0, atol=1e-13)
    assert np.array_equal(jb.slice_patches(dev(torch, z["a2"]), (15, 15), 5, 0).cpu().numpy(), z["sp2"])
    # adjointness with a leading axis (test_slice_patches.py:55-66)
    a = synth.direction((3, 36, 36), seed=21)
    b = synth.direction((256, 3, 8, 8), seed=22)
    lhs = np.vdot(b, jb.slice_patches(dev(torch, a), (36, 36), 16, 2).cpu().numpy())
    rhs = np.vdot(jb.slice_patches_T(dev(torch, b), (36, 36), 16, 2).cpu().numpy(), a)
    assert abs(lhs - rhs) / abs(lhs) < 1e-13
    c = np.load(os.path.join(golden_dir, "convolve.npz"))
    dom = jb.Domain(tuple(c["sig"].shape), tuple(c["distances"]))
    c1 = jb.convolve(dev(torch, c["ker"]), dev(torch, c["sig"]), dom, axes=(-2, -1)).cpu().numpy()
    assert rel(c1, c["c1"]) < 1e-12
    capsys.readouterr()
    c2 = jb.convolve(dev(torch, c["ker2"]), dev(torch, c["sig"]), dom, axes=(-2, -1)).cpu().numpy()
    assert "Dimension Inconsistency. Broadcasting PSFs" in capsys.readouterr().out
    assert c2.shape == c["sig"].shape and rel(c2, c["c2"]) < 1e-12
    # delta kernel => identity (test_convolve.py:32-41)
    k = np.zeros((4, 4)); k[0, 0] = 1.0
    sig = np.random.default_rng(0).normal(size=(4, 4))
    res = jb.convolve(dev(torch, k), dev(torch, sig), jb.Domain((4, 4), (1.0, 1.0)), axes=(-2, -1)).cpu().numpy()
    np.testing.assert_allclose(res, sig, rtol=1e-10, atol=1e-10)


def test_device_psf_interpolator_feeds_the_plan_without_the_full_size_array(torch_cuda, golden_dir):
    """psf_interpolator (erosita_psf_utils.py:198-264) on the device against the reference source's output
    (both layouts), and end to end: the cropped kernels it writes for two modules go straight into a ResponsePlan whose
    PSF output equals the oracle's linpatch_convolve with the reference's full-size interpolated kernels."""
    import jubik_b200 as jb
    torch = torch_cuda
    for name in ("psf_interp_N48_n4", "psf_interp_N40_n5"):
        z = np.load(os.path.join(golden_dir, name + ".npz"))
        E, N, n, dist = int(z["E"]), int(z["N"]), int(z["n"]), float(z["dist"])
        infos = synth.psf_infos(E, int(z["nrad"]), int(z["px"]), seed=int(z["seed"]))
        dom = jb.Domain((E, N, N), (1.0, dist, dist))
        full = jb.psf_interpolator(dom, n, infos)
        assert tuple(full.shape) == z["kernel"].shape and rel(full.cpu().numpy(), z["kernel"]) < 1e-12
        crop = jb.psf_interpolator(dom, n, infos, cropped=True)
        assert rel(crop.cpu().numpy(), orc.crop_kernel(z["kernel"], (N, N), n)) < 1e-12
        with pytest.raises(ValueError):
            jb.psf_interpolator(jb.Domain((E, N + 2, N + 2), (1.0, dist, dist)), n, infos)
    # two modules with different pointings -> (T, n*n, E, 2dx, 2dx) written module by module -> plan
    E, N, n, m, dist = 2, 240, 4, 10, 8.0
    dx = N // n
    dom = jb.Domain((E, N, N), (1.0, dist, dist))
    dom_o = orc.Domain(dom.shape, dom.distances)
    mods = [synth.psf_infos(E, 4, 33, seed=31, pointing=(500.0, 900.0)), synth.psf_infos(E, 4, 33, seed=32, pointing=(1200.0, 300.0))]
    kern = torch.empty((2, n * n, E, 2 * dx, 2 * dx), dtype=torch.float64, device="cuda")
    for t, infos in enumerate(mods):
        jb.psf_interpolator(dom, n, infos, cropped=True, out=kern[t])
    plan = jb.ResponsePlan(kern, N, n, m, cropped=True, distances=dom.distances[-2:])
    x = synth.sky(1, E, N)[0]
    y = plan.psf(dev(torch, x)).cpu().numpy()
    for t, infos in enumerate(mods):
        kfull = orc.psf_interpolator(dom_o, n, infos)
        assert rel(y[t], orc.linpatch_convolve(x, dom_o, kfull, n, m)) < TOL
    plan.close()


def test_general_convolve_runs_on_the_fft_kernels(torch_cuda, capsys):
    """jubik.convolve (convolve.py:148-176) for arrays whose sides are compiled transform lengths: square and
    rectangular, any leading shape, NumPy-style broadcasting of the leading axes and the N-D form of the reference's
    broadcasting branch (hx[..., newaxis, :, :] * hy) -- against the oracle's c2c fftn form."""
    import jubik_b200 as jb
    torch = torch_cuda
    rng = np.random.default_rng(12)
    for shape_x, shape_y in (((256, 256), (256, 256)),            # square, no leading axes
                             ((2, 3, 96, 160), (2, 3, 96, 160)),  # rectangular, leading shape
                             ((1, 3, 128, 128), (2, 1, 128, 128)),  # leading axes broadcast against each other
                             ((512, 384), (2, 512, 384)),         # broadcasting branch, 2-D kernel
                             ((3, 160, 96), (3, 4, 160, 96)),     # broadcasting branch, N-D kernel: (3,1,..) * (3,4,..)
                             ((1024, 1024), (1024, 1024)),        # the largest compiled length
                             ((40, 28), (2, 40, 28))):            # not compiled: direct sum
        xk, ys = rng.normal(size=shape_x), rng.normal(size=shape_y)
        dom_o = orc.Domain(tuple(shape_y), (1.0,) * (len(shape_y) - 2) + (0.5, 2.0))
        capsys.readouterr()
        ref = orc.convolve(xk, ys, dom_o, axes=(-2, -1))
        ref_out = capsys.readouterr().out
        got = jb.convolve(dev(torch, xk), dev(torch, ys), jb.Domain(dom_o.shape, dom_o.distances), axes=(-2, -1))
        assert capsys.readouterr().out == ref_out
        assert tuple(got.shape) == ref.shape and rel(got.cpu().numpy(), ref) < 1e-12, (shape_x, shape_y)
    with pytest.raises(jb.JubikError, match="compiled transform length"):
        jb.convolve(dev(torch, rng.normal(size=(300, 300))), dev(torch, rng.normal(size=(300, 300))),
                    jb.Domain((300, 300), (1.0, 1.0)), axes=(-2, -1))


def test_error_behaviour_matches_reference(torch_cuda):
    import jubik_b200 as jb
    torch = torch_cuda
    x = torch.zeros((1, 32, 32), dtype=torch.float64, device="cuda")
    k = torch.ones((16, 1, 32, 32), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):  # convolve.py:82-83
        jb.linpatch_convolve(x, ((1, 32, 32), (1, 1, 1)), k, 4, 4)
    with pytest.raises(jb.JubikError):  # n <= 2 breaks the reference's kernel crop (Appendix A4)
        jb.ResponsePlan(torch.ones((1, 4, 1, 32, 32), dtype=torch.float64, device="cuda"), 32, 2, 4)
    with pytest.raises(jb.JubikError):  # negative cut, response.py:37-39
        jb.ResponsePlan(k[None], 32, 4, 4, exposure=torch.ones((1, 1, 32, 32), dtype=torch.float64, device="cuda"),
                        exposure_cut=-1.0)
    # non-smooth P (42) with a PSF wider than the margin: served by the alias fold since r2 (was JUBIK_ERR_UNSUPPORTED)
    p = jb.ResponsePlan(torch.ones((1, 16, 1, 64, 64), dtype=torch.float64, device="cuda"), 64, 4, 5)
    assert p.info["alias_fold"] == 1
    with pytest.raises(jb.JubikError, match="fft_length"):   # a forced length that is too short for the fold
        jb.ResponsePlan(torch.ones((1, 16, 1, 64, 64), dtype=torch.float64, device="cuda"), 64, 4, 5, fft_length=96 // 2 + 0)
    with pytest.raises((TypeError, ValueError)):     # an input on the wrong device / a CPU tensor never reaches the C ABI
        p.psf(torch.zeros((1, 64, 64), dtype=torch.float64))


def test_erosita_call_form_dynamic_kernel_mask_adj_mock_data_and_nwr(torch_cuda):
    """The eROSITA form of the response dictionary (erosita_response.py:349-358): 'kernel' in the dict, R(x, k) and
    psf(x, k) with the kernel as an argument (re-planned only when its content changes), FullModel with the kernel as
    a dynamic leaf (erosita_likelihood.py:73-83), 'mask_adj' (data.py:156-160), the mock-data draw (data.py:138-146)
    and the noise-weighted residuals over a batch of samples (diagnostics.py:96-105) -- all against the oracle."""
    import jubik_b200 as jb
    torch = torch_cuda
    N, n, m, T, E, S = 256, 4, 16, 2, 2, 3
    dx = N // n
    dom_o = orc.Domain((E, N, N), synth.DIST)
    dom = jb.Domain(dom_o.shape, dom_o.distances)
    k1 = synth.psf_cropped(T, n, E, dx, half_support=None)
    k2 = synth.psf_cropped(T, n, E, dx, half_support=None, seed=9, sig=(2.0, 7.0))
    expo, ea, x = synth.exposure(T, E, N), synth.eff_area(T, E), synth.sky(S, E, N)
    keys = ["tm%d" % (t + 1) for t in range(T)]
    R1 = orc.Response(dom_o, k1, n, m, expo, 500.0, synth.PIXEL_AREA, eff_area=ea, keys=keys, cropped=True)
    R2 = orc.Response(dom_o, k2, n, m, expo, 500.0, synth.PIXEL_AREA, eff_area=ea, keys=keys, cropped=True)
    rd = jb.build_response(dev(torch, k1), dev(torch, expo), dom, n, m, synth.PIXEL_AREA, exposure_cut=500.0,
                           effective_area=dev(torch, ea), keys=keys, max_samples=S, cropped=True, dynamic_kernel=True)
    assert 'kernel' in rd and 'mask_adj' in rd
    plan = rd['plan']
    xd = dev(torch, x)

    def same(vec, ref):
        assert list(vec.keys()) == keys
        for k in keys:
            assert vec[k].dim() == 1 and rel(vec[k].cpu().numpy(), ref.tree[k]) < TOL

    l0 = plan.launch_count
    same(rd['R'](xd[0], rd['kernel']), R1(x[0]))
    first = plan.launch_count - l0
    same(rd['R'](xd[0], rd['kernel']), R1(x[0]))
    assert plan.launch_count - l0 == 2 * first              # the dictionary's own kernel: no fingerprint, no re-plan
    k2d = dev(torch, k2)
    same(rd['R'](xd[0], k2d), R2(x[0]))                     # another kernel: spectra rebuilt
    assert rel(rd['psf'](xd[0], k2d).cpu().numpy(), R2.psf(x[0] * 1.0)) < TOL
    assert plan.ensure_kernel(k2d.clone()) is False         # equal content in a new tensor: fingerprint hit
    assert plan.ensure_kernel(dev(torch, k1)) is True
    model = jb.FullModel(kern=rd['kernel'], instrument=rd['R'], pre_ops=lambda v: 2.0 * v)
    same(model(xd[1]), R1(2.0 * x[1]))
    # mask_adj: transpose of the readout
    lam = R1(x[0])
    dense = rd['mask_adj'](jb.Vector({k: dev(torch, lam.tree[k]) for k in keys}))[0]
    assert rel(dense.cpu().numpy(), R1.dense(x[0])) < TOL
    # mock data: Poisson draw of R(sky, kernel) per key
    mock = jb.create_mock_data(rd, xd[0], seed=11)
    for k in keys:
        assert mock[k].dtype == torch.int64 and mock[k].shape == (lam.tree[k].size,)
        assert abs(float(mock[k].double().mean()) / lam.tree[k].mean() - 1.0) < 0.02
    # likelihood on the compacted data, kernel as a per-call argument
    d_dense = synth.counts_from(R1.dense(synth.sky(1, E, N, seed=5)[0]))
    d_vec = jb.Vector({k: dev(torch, d_dense[t][~R1.mask[t]].astype(np.float64)) for t, k in enumerate(keys)})
    lh = jb.PoissonLikelihood(rd, d_vec)
    e, g = lh.value_and_grad(xd[:1], k=rd['kernel'])
    eo, go = orc.energy_and_grad(R1, d_dense, x[0])
    assert abs(float(e[0]) - eo) / abs(eo) < TOL and rel(g[0].cpu().numpy(), go) < TOL
    e2, _ = lh.value_and_grad(xd[:1], k=k2d)
    assert abs(float(e2[0]) - orc.energy_and_grad(R2, d_dense, x[0])[0]) / abs(eo) < TOL
    # noise-weighted residuals of S samples (the reference's vmap over samples = the batched response)
    res, tot = jb.calculate_nwr([xd[s] for s in range(S)], lambda v: v, d_vec, rd, abs=True, min_counts=1)
    assert tuple(res.shape) == (T, S, E, N, N) and tuple(tot.shape) == (T, S, E, N, N)
    for s in range(S):
        lam_s = R1.dense(x[s])
        with np.errstate(divide="ignore", invalid="ignore"):
            ref = np.where(R1.mask, 0.0, np.abs(lam_s - d_dense) / np.sqrt(lam_s))
        assert rel(res[:, s].cpu().numpy(), ref) < 1e-9
    assert np.array_equal(tot[:, 0].cpu().numpy(), R1.mask | (d_dense < 1))
    plan.close()


def test_sharded_call_in_chunks_and_its_host_buffer_form(torch_cuda):
    """More samples than the workspace holds (bench.py's C5 on one GPU: 16 samples in chunks of 2): the chunked sharded
    call returns every sample's energy bit for bit and the sample-summed gradient of a plan that holds all samples at
    once; the host-buffer form (x ring on a copy stream, one bin at a time) is bitwise equal to the device form."""
    import jubik_b200 as jb
    torch = torch_cuda
    case = (512, 4, 64, 2, 3, 5, False, True)
    N, n, m, T, E, S, full, eff = case
    dom, kern, cropped, expo, ea, x, R, d = build_case(*case)

    def mk(smax, **kw):
        return jb.ResponsePlan(dev(torch, kern), N, n, m, exposure=dev(torch, expo), exposure_cut=500.0,
                               eff_area=dev(torch, ea), pixel_area=synth.PIXEL_AREA, data=dev(torch, d), max_samples=smax,
                               cropped=cropped, **kw)
    xd = dev(torch, x)
    big = mk(S)
    e0, g0 = big.energy_and_grad(xd)
    eo, go = zip(*[orc.energy_and_grad(R, d, x[s]) for s in range(2)])
    assert rel(g0[:2].cpu().numpy(), np.stack(go)) < TOL
    npx = E * N * N
    p = mk(2, host_call=True, sharded=True)
    for want_local in (False, True):
        gl = torch.empty_like(xd) if want_local else None
        en, gs = p.energy_and_grad_sharded(xd, None, jb.SHARD_SAMPLES, grad_local=gl)
        assert torch.equal(en.flatten(), e0)
        assert rel(gs.cpu().numpy(), g0.sum(0).cpu().numpy()) < 1e-13
        if want_local:
            assert torch.equal(gl, g0)
    flat = torch.empty(npx + S, dtype=torch.float64, device="cuda")
    p.energy_and_grad_sharded(xd, None, jb.SHARD_SAMPLES, flat=flat)
    xh = torch.as_tensor(x).pin_memory()
    fh = torch.empty(npx + S, dtype=torch.float64).pin_memory()
    for _ in range(3):
        fh.zero_()
        p.energy_and_grad_sharded_host(xh, None, fh)
        assert torch.equal(fh, flat.cpu())
    em, gm = p.energy_and_grad_sharded(xd, None, jb.SHARD_MODULES)
    assert torch.equal(em, e0) and torch.equal(gm, g0)
    with pytest.raises(jb.JubikError, match="JUBIK_PLAN_HOST_CALL"):
        big.energy_and_grad_sharded_host(xh, None, fh)
    p.close()
    big.close()


def test_sharded_entry_point_with_one_rank_nccl_communicator(torch_cuda):
    """jubik_poisson_energy_and_grad_sharded through a real (one-rank) NCCL communicator created by the C library:
    dlopen of libnccl, ncclCommInitRank, the fold kernel and an in-place ncclAllReduce on the plan's stream."""
    import jubik_b200 as jb
    torch = torch_cuda
    case = (240, 4, 10, 2, 2, 3, False, True)
    N, n, m, T, E, S, full, eff = case
    dom, kern, cropped, expo, ea, x, R, d = build_case(*case)
    plan = jb.ResponsePlan(dev(torch, kern), N, n, m, exposure=dev(torch, expo), exposure_cut=500.0,
                           eff_area=dev(torch, ea), pixel_area=synth.PIXEL_AREA, data=dev(torch, d), max_samples=S,
                           cropped=cropped, sharded=True)
    comm = jb.Communicator(jb.Communicator.unique_id(), 0, 1)
    assert comm.size == 1 and comm.rank == 0
    xd = dev(torch, x)
    e0, g0 = plan.energy_and_grad(xd)
    gl = torch.empty_like(xd)
    e, g = plan.energy_and_grad_sharded(xd, comm, jb.SHARD_SAMPLES, grad_local=gl)
    assert e.shape == (1, S) and torch.equal(e[0], e0) and torch.equal(gl, g0)
    eo, go = zip(*[orc.energy_and_grad(R, d, x[s]) for s in range(S)])
    assert rel(g.cpu().numpy(), np.stack(go).sum(0)) < TOL
    assert torch.equal(g, (g0[0] + g0[1]) + g0[2])        # fixed summation order
    e1, g1 = plan.energy_and_grad_sharded(xd[:1], None, jb.SHARD_SAMPLES)       # no communicator, gradient in place
    assert torch.equal(e1[0], e0[:1]) and torch.equal(g1, g0[0])
    em, gm = plan.energy_and_grad_sharded(xd, comm, jb.SHARD_MODULES)
    assert torch.equal(em, e0) and torch.equal(gm, g0)
    comm.close()
    plan.close()


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import jubik_b200 as jb
    from jubik_b200.sharding import NcclShardedLikelihood, owned
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)       # plumbing for the ncclUniqueId only
    case = (240, 4, 10, 2, 2, 4, False, True)
    N, n, m, T, E, S, full, eff = case
    dom, kern, cropped, expo, ea, x, R, d = build_case(*case)

    def mk(tsel, smax):
        return jb.ResponsePlan(torch.as_tensor(kern[tsel]).cuda(), N, n, m, exposure=torch.as_tensor(expo[tsel]).cuda(),
                               exposure_cut=500.0, eff_area=torch.as_tensor(ea[tsel]).cuda(),
                               pixel_area=synth.PIXEL_AREA, data=torch.as_tensor(d[tsel]).cuda(), max_samples=smax,
                               cropped=cropped, sharded=True)
    comm = jb.Communicator.from_torch_distributed()
    sl = S // world
    lh = NcclShardedLikelihood(mk(list(range(T)), sl), "samples", comm=comm)
    e_s, g_s = lh.energy_and_grad(torch.as_tensor(x[rank * sl:(rank + 1) * sl]).cuda())
    lm = NcclShardedLikelihood(mk(owned(T, rank, world), S), "modules", comm=comm)
    e_m, g_m = lm.energy_and_grad(torch.as_tensor(x).cuda())
    torch.cuda.synchronize()
    if rank == 0:
        q.put((e_s.cpu().numpy(), g_s.cpu().numpy(), e_m.cpu().numpy(), g_m.cpu().numpy()))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_two_gpu_nccl_sharding_matches_oracle(torch_cuda):
    """Two ranks, two GPUs: sample sharding and module sharding with the ncclAllReduce issued by the C library."""
    import socket
    torch = torch_cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    e_s, g_s, e_m, g_m = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    case = (240, 4, 10, 2, 2, 4, False, True)
    dom, kern, cropped, expo, ea, x, R, d = build_case(*case)
    eo, go = zip(*[orc.energy_and_grad(R, d, x[s]) for s in range(4)])
    eo, go = np.array(eo), np.stack(go)
    assert np.abs((e_s - eo) / eo).max() < TOL and rel(g_s, go.sum(0)) < TOL
    assert np.abs((e_m - eo) / eo).max() < TOL and rel(g_m, go) < TOL
