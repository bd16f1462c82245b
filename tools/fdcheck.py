import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import jubik_b200 as jb
from tests import synth
N, n, m, T, E, S = 1024, 8, 64, 1, 3, 1
dx = N // n
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
kern = dev(synth.psf_cropped(T, n, E, dx, half_support=m)); expo = dev(synth.exposure(T, E, N))
plan = jb.ResponsePlan(kern, N, n, m, exposure=expo, exposure_cut=500.0, pixel_area=synth.PIXEL_AREA, cropped=True, max_samples=S)
x = dev(synth.sky(S, E, N))
lam = plan.response(dev(synth.sky(1, E, N, seed=5)))
torch.manual_seed(0)
d = torch.poisson(lam.clamp(min=0.0))[0]
plan.set_data(d)
v = dev(synth.direction((S, E, N, N), seed=9)) * x
for rep in range(3):
    e0, gr = plan.energy_and_grad(x)
    for eps in (1e-3, 1e-4, 1e-5):
        ep = plan.energy(x + eps * v); em = plan.energy(x - eps * v)
        fd = float((ep - em).sum()) / (2 * eps); an = float(torch.vdot(gr.flatten(), v.flatten()))
        print(rep, eps, "e0 %.10e fd %.8e an %.8e rel %.2e" % (float(e0[0]), fd, an, abs(fd-an)/abs(an)))
    e1, gr1 = plan.energy_and_grad(x)
    print("deterministic:", torch.equal(e0, e1), torch.equal(gr, gr1), float((gr-gr1).abs().max()))
