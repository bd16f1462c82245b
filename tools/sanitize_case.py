"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once.

    compute-sanitizer --tool racecheck python tools/sanitize_case.py
Cases: C1 (256^2, 4x4, P = L = 160: TMA column pipeline + warp-private c2r rows + prefetching r2c rows), a padded length
(P = 140 -> L = 160), the WIDE c2r form (10x10 patches), the module-sharing column kernels with their k1-group barriers (two modules, L = 96), the alias fold on the generic kernels, and the stand-alone helpers.
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import jubik_b200 as jb  # noqa: E402
from tests import synth  # noqa: E402


def run(N, n, m, T, E, S, full, tag):
    dx = N // n
    kern = synth.psf_full(T, n, E, N) if full else synth.psf_cropped(T, n, E, dx, half_support=m)
    expo = synth.exposure(T, E, N)
    x = torch.as_tensor(synth.sky(S, E, N)).cuda()
    plan = jb.ResponsePlan(torch.as_tensor(kern).cuda(), N, n, m, exposure=torch.as_tensor(expo).cuda(), exposure_cut=500.0,
                           pixel_area=16.0, max_samples=S, cropped=not full, host_call=True, sharded=True)
    lam = plan.response(x)[0]
    torch.manual_seed(3)
    plan.set_data(torch.poisson(lam.clamp(min=0.0)))
    e, g = plan.energy_and_grad(x)
    mv = plan.metric(x, x)
    el, gl = plan.energy_and_grad_logsky(torch.log(x))
    flat = torch.empty(E * N * N + S, dtype=torch.float64, device="cuda")
    plan.energy_and_grad_sharded(x, None, jb.SHARD_SAMPLES, flat=flat)
    xh = x.cpu().pin_memory()
    fh = torch.empty(E * N * N + S, dtype=torch.float64).pin_memory()
    plan.energy_and_grad_sharded_host(xh, None, fh)
    torch.cuda.synchronize()
    print("%s: L=%d fast=%d fold=%d energy %.6e |g| %.3e |Mv| %.3e host==dev %s" % (
        tag, plan.info["L"], plan.info["fast_path"], plan.info["alias_fold"], float(e[0]), float(g.abs().max()),
        float(mv.abs().max()), bool(torch.equal(fh, flat.cpu()))))
    plan.close()


run(256, 4, 16, 1, 2, 2, True, "C1 exact period")
run(240, 4, 10, 2, 1, 1, False, "padded length")
run(330, 10, 17, 1, 1, 1, False, "wide c2r form")
run(256, 8, 16, 2, 1, 1, False, "module-sharing column kernels (L = 96 = 4*4*6: register spectrum / accumulator, k1-group barriers)")
run(100, 5, 7, 1, 2, 1, True, "alias fold")
rng = np.random.default_rng(0)
a = torch.as_tensor(rng.normal(size=(2, 96, 160))).cuda()
c = jb.convolve(a, a, jb.Domain((2, 96, 160), (1.0, 1.0, 1.0)), axes=(-2, -1))
sp = jb.slice_patches(a[:, :96, :96].contiguous(), (96, 96), 4, 2)
spT = jb.slice_patches_T(sp, (96, 96), 4, 2)
torch.cuda.synchronize()
print("helpers ok", float(c.abs().max()), tuple(sp.shape), tuple(spT.shape))
