"""Quick per-kernel timing of the fused fwd+VJP step (C2 by default) for kernel tuning: prints ms/step and the
per-kernel CUDA-event times.  Not the bench contract (see bench.py)."""
import os
import statistics
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402
import jubik_b200 as jb  # noqa: E402
from bench import WORKLOADS, make_inputs  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
N, n, m, T, E, _ = WORKLOADS[wl]
kern, expo, x = make_inputs(wl, 1)
plan = jb.ResponsePlan(torch.as_tensor(kern).cuda(), N, n, m, exposure=torch.as_tensor(expo).cuda(),
                       exposure_cut=500.0, pixel_area=16.0, cropped=True)
xd = torch.as_tensor(x).cuda()
lam = plan.response(torch.as_tensor(make_inputs(wl, 1, seed_shift=5)[2]).cuda())[0]
torch.manual_seed(3)
plan.set_data(torch.poisson(lam.clamp(min=0.0)))
flush = torch.empty(192 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
for _ in range(3):
    plan.energy_and_grad(xd)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, b in ev:
    flush.zero_()
    a.record()
    e, g = plan.energy_and_grad(xd)
    b.record()
torch.cuda.synchronize()
ms = statistics.mean(a.elapsed_time(b) for a, b in ev)
acc = {}
for _ in range(10):
    flush.zero_()
    _, _, k = plan.profile(xd)
    for kk, v in k.items():
        acc.setdefault(kk, []).append(v)
print("%s %s: %.4f ms/step  energy %.10e | " % (os.environ.get("JUBIK_B200_LIB", "default").split("/")[-1], wl, ms, float(e[0])) +
      " ".join("%s %.1f" % (k, 1e3 * statistics.mean(v)) for k, v in acc.items()))
