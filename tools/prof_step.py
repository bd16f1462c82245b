"""Tiny driver for ncu: build a plan of one bench workload (optionally with fewer modules, so that a `--set full`
capture with its ~40 replays per launch stays short) and run a few fused fwd+VJP steps.

    python tools/prof_step.py [C5|C2|...] [steps] [samples] [modules]
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402
import jubik_b200 as jb  # noqa: E402
from bench import WORKLOADS, sample_skies  # noqa: E402
from tests import synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C5"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
N, n, m, T, E, S_job, chunk, _ = WORKLOADS[wl]
S = int(sys.argv[3]) if len(sys.argv) > 3 else chunk
T = int(sys.argv[4]) if len(sys.argv) > 4 else T
kern = synth.psf_cropped(T, n, E, N // n, half_support=m)
expo = synth.exposure(T, E, N)
plan = jb.ResponsePlan(torch.as_tensor(kern).cuda(), N, n, m, exposure=torch.as_tensor(expo).cuda(),
                       exposure_cut=500.0, pixel_area=16.0, cropped=True, max_samples=S, metric=False)
xd = torch.as_tensor(sample_skies(wl, 0, S)).cuda()
lam = plan.response(torch.as_tensor(synth.sky(1, E, N, seed=5)).cuda())[0]
plan.set_data(torch.poisson(lam.clamp(min=0.0)))
for _ in range(steps):
    e, g = plan.energy_and_grad(xd)
torch.cuda.synchronize()
print(float(e[0]))
