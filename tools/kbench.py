"""Quick per-kernel timing of the fused fwd+VJP step for kernel tuning: prints ms/step, the per-kernel CUDA-event
times and a checksum of the results (variants must reproduce it).  Not the bench contract (see bench.py).

    [JUBIK_B200_LIB=jubik_b200/_variants/lib_x.so] python tools/kbench.py [C2|C5|...] [samples]
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402
import jubik_b200 as jb  # noqa: E402
from bench import WORKLOADS, make_inputs, sample_skies  # noqa: E402
from tests import synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
N, n, m, T, E, S_job, chunk, _ = WORKLOADS[wl]
S = int(sys.argv[2]) if len(sys.argv) > 2 else chunk
kern, expo = make_inputs(wl)
plan = jb.ResponsePlan(torch.as_tensor(kern).cuda(), N, n, m, exposure=torch.as_tensor(expo).cuda(),
                       exposure_cut=500.0, pixel_area=16.0, cropped=True, max_samples=S, metric=False)
del kern, expo
xd = torch.as_tensor(sample_skies(wl, 0, S)).cuda()
lam = plan.response(torch.as_tensor(synth.sky(1, E, N, seed=5)).cuda())[0]
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
import hashlib  # noqa: E402
h = hashlib.sha1(g.cpu().numpy().tobytes() + e.cpu().numpy().tobytes()).hexdigest()[:12]
px = S * T * E * N * N
print("%s %s S=%d: %.4f ms/step %.2f Gpx/s sha %s | " % (os.environ.get("JUBIK_B200_LIB", "default").split("/")[-1], wl, S, ms,
                                                       px / ms / 1e6, h) +
      " ".join("%s %.1f" % (k, 1e3 * statistics.mean(v)) for k, v in acc.items()))
