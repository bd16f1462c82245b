"""The jax.ffi boundary (BASELINE north_star; VERDICT r1 row b').

CPU part (always runs): csrc/jax_ffi.cc is well-formed C++ against the shape of jaxlib's xla/ffi/api/ffi.h (a mock of the
subset it uses, tests/mock_xla -- jax is not installable in the build image) and every FFI target the Python half
registers has its handler symbol in the C++ half.
GPU + jax part (skipped without jax): jit / grad / vmap / linear_transpose of the bound callables against the oracle.
"""
import ast
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
CC = os.path.join(ROOT, "jubik_b200", "csrc", "jax_ffi.cc")
PY = os.path.join(ROOT, "jubik_b200", "jax_binding.py")


def _targets():
    tree = ast.parse(open(PY).read())
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "TARGETS":
            return ast.literal_eval(node.value)
    raise AssertionError("TARGETS not found")


def test_ffi_handlers_compile_against_the_ffi_api_shape():
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "mock_xla"),
                        "-I", "/usr/local/cuda/include", CC], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]


def test_every_registered_target_has_a_handler_and_every_c_abi_call_exists():
    src = open(CC).read()
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", src))
    targets = _targets()
    assert set(targets.values()) == defined
    header = open(os.path.join(ROOT, "include", "jubik_b200.h")).read()
    for fn in set(re.findall(r"\b(jubik_[a-z_A-Z0-9]+)\(", src)):
        assert re.search(r"\b%s\(" % fn, header), fn + " is not declared in include/jubik_b200.h"


# ------------------------------------------------------------------------------------------------ needs jax + a GPU
@pytest.fixture(scope="module")
def jbx():
    jax = pytest.importorskip("jax")
    try:
        if not any(d.platform == "gpu" for d in jax.devices()):
            pytest.skip("jax has no CUDA device")
    except RuntimeError:
        pytest.skip("jax has no CUDA device")
    jax.config.update("jax_enable_x64", True)
    import jubik_b200.jax_binding as m
    return m


@pytest.mark.gpu
def test_jax_response_likelihood_jit_grad_vmap_transpose(jbx):
    import jax
    import jax.numpy as jnp
    from oracle import jubik_oracle as orc
    from tests import synth
    import jubik_b200 as jb
    N, n, m, T, E, S = 256, 4, 16, 2, 2, 3
    dx = N // n
    dom_o = orc.Domain((E, N, N), synth.DIST)
    kern = synth.psf_cropped(T, n, E, dx, half_support=None)
    k2 = synth.psf_cropped(T, n, E, dx, half_support=None, seed=9, sig=(2.0, 7.0))
    expo, ea, x = synth.exposure(T, E, N), synth.eff_area(T, E), synth.sky(S, E, N)
    keys = ["tm1", "tm2"]
    R = orc.Response(dom_o, kern, n, m, expo, 500.0, synth.PIXEL_AREA, eff_area=ea, keys=keys, cropped=True)
    R2 = orc.Response(dom_o, k2, n, m, expo, 500.0, synth.PIXEL_AREA, eff_area=ea, keys=keys, cropped=True)
    d = synth.counts_from(R.dense(synth.sky(1, E, N, seed=5)[0]))
    rd = jbx.build_response(kern, expo, jb.Domain(dom_o.shape, dom_o.distances), n, m, synth.PIXEL_AREA, exposure_cut=500.0,
                            effective_area=ea, keys=keys, max_samples=2, cropped=True, dynamic_kernel=True)
    tol = 1e-10

    def rel(a, b):
        return np.abs(np.asarray(a) - b).max() / np.abs(b).max()
    lam = jax.jit(rd['R'])(x[0], rd['kernel'])
    for k in keys:
        assert rel(lam[k], R(x[0]).tree[k]) < tol
    lam2 = rd['R'](x[0], jnp.asarray(k2))
    for k in keys:
        assert rel(lam2[k], R2(x[0]).tree[k]) < tol
    # vmap over samples (3 samples, plan holds 2: chunked inside the handler) and the transpose
    dense = jax.vmap(rd['R_dense'])(jnp.asarray(x))
    assert rel(dense, np.stack([R2.dense(x[s]) for s in range(S)])) < tol
    rd['R'](x[0], rd['kernel'])
    g = synth.direction((T, E, N, N))
    (gt,) = jax.linear_transpose(rd['R_dense'], jnp.asarray(x[0]))(jnp.asarray(g))
    assert rel(gt, R.dense_T(g)) < tol
    # fused likelihood: value_and_grad is one custom call
    data_vec = {k: d[t][~R.mask[t]] for t, k in enumerate(keys)}
    H = jbx.poisson_energy(rd, data_vec)
    e, gr = jax.jit(jax.value_and_grad(H))(jnp.asarray(x[0]))
    eo, go = orc.energy_and_grad(R, d, x[0])
    assert abs(float(e) - eo) / abs(eo) < tol and rel(gr, go) < tol
    es = jax.vmap(H)(jnp.asarray(x))
    assert rel(es, np.array([orc.energy_and_grad(R, d, x[s])[0] for s in range(S)])) < tol
    plan = rd['plan']
    v = synth.direction((E, N, N), seed=10)
    assert rel(plan.metric(x[0], v), orc.metric_apply(R, x[0], v)) < tol
    tok = plan.set_point(x[0])
    assert rel(plan.metric_at_point(v, tok), orc.metric_apply(R, x[0], v)) < tol
    assert rel(plan.left_sqrt_metric(x[0], g), orc.left_sqrt_metric(R, x[0], g)) < tol
    Hl = jbx.poisson_energy(rd, data_vec, logsky=True)
    el, gl = jax.value_and_grad(Hl)(jnp.log(jnp.asarray(x[0])))
    elo, glo = orc.energy_and_grad_logsky(R, d, np.log(x[0]))
    assert abs(float(el) - elo) / abs(elo) < tol and rel(gl, glo) < tol
