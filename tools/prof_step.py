"""Tiny driver for ncu: build the C2 plan (BASELINE.json configs[1]) and run a few fused fwd+VJP steps."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402
import jubik_b200 as jb  # noqa: E402
from bench import WORKLOADS, make_inputs  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1
N, n, m, T, E, _ = WORKLOADS[wl]
kern, expo, x = make_inputs(wl, S)
plan = jb.ResponsePlan(torch.as_tensor(kern).cuda(), N, n, m, exposure=torch.as_tensor(expo).cuda(),
                       exposure_cut=500.0, pixel_area=16.0, cropped=True, max_samples=S)
xd = torch.as_tensor(x).cuda()
lam = plan.response(torch.as_tensor(make_inputs(wl, 1, seed_shift=5)[2]).cuda())[0]
plan.set_data(torch.poisson(lam.clamp(min=0.0)))
for _ in range(steps):
    e, g = plan.energy_and_grad(xd)
torch.cuda.synchronize()
print(float(e[0]))
