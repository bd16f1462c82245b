"""Timing of the Fisher-metric apply (what NIFTy's CG issues 60-70x per sample, eROSITA_demo.yaml:161): the stand-alone
call `metric(x, v)` (two forward passes + one adjoint) against the cached linearisation point
(`set_point(x)` once, then `metric_at_point(v)` = one forward + one adjoint).  CUDA events, L2 flushed between calls.

    python tools/metric_bench.py [C2|C3|C5] [samples]
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
                       exposure_cut=500.0, pixel_area=16.0, cropped=True, max_samples=S, metric=True)
del kern, expo
x = torch.as_tensor(sample_skies(wl, 0, S)).cuda()
v = torch.as_tensor(synth.direction((S, E, N, N), seed=9)).cuda()
flush = torch.empty(192 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in ev)


t_full = timed(lambda: plan.metric(x, v))
t_set = timed(lambda: plan.set_point(x))
t_at = timed(lambda: plan.metric_at_point(v))
ref, got = plan.metric(x, v), plan.metric_at_point(v)
err = float((ref - got).abs().max() / ref.abs().max())
print("%s S=%d: metric(x, v) %.3f ms | set_point %.3f ms, metric_at_point %.3f ms (%.2fx of the stand-alone call) | "
      "max rel. difference %.1e" % (wl, S, t_full, t_set, t_at, t_at / t_full, err))
