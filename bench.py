#!/usr/bin/env python
"""bench.py -- fused forward + VJP of response + Poisson log-likelihood, fp64, in sky-pixels/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C1|C3s|C3|C4|C5g]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU)

Workload at N=1: BASELINE.json configs[1] ("C2", Chandra-like): 1024x1024 sky, 3 energy bins, 8x8
spatially-variant PSF patches, margin 64, exposure + readout mask + Poisson likelihood, S=1, T=1.
One step = one evaluation of (energy[s], d energy[s] / d sky[s]) for the S samples of the rank; sky-pixels =
S*T*E*N^2.
Every rank (also at N=1) evaluates `--samples-per-gpu` posterior samples of that workload per step (default 2, the
per-GPU share of BASELINE.json configs[4]: 16 geoVI samples over 8 GPUs) -- weak scaling, sample sharding;
one step is ONE C-ABI call (jubik_poisson_energy_and_grad_sharded): the seven kernels, the fold of the rank's samples
into a flat buffer [sample-summed gradient | per-sample energies of all ranks] and a SINGLE ncclAllReduce of that buffer,
issued by the library on its own NCCL communicator, all on one stream inside the timed region.

`value`  : device-resident throughput (inputs already in HBM), CUDA events, max over ranks.
`e2e`    : the same metric through ResponsePlan.energy_and_grad_host (C ABI jubik_poisson_energy_and_grad_host)
           with pinned HOST buffers: H2D of x and D2H of energy+gradient inside the timed region.
`roofline`: dominant kernel (column pass k_cols, forward direction) -- algorithmic bytes / live CUDA-event
           duration against MEASURED_PEAKS.json hbm_gbs; whole-path HBM and FP64 fractions alongside.
`cpu_baseline`: the NumPy/scipy oracle (the reference's op sequence; jax is not installable here) timed on
           this box's host cores on a bounded sample.  --impl reference prints that arm alone.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "fwd+VJP response+loglik sky-pixels/s (fp64)"
UNIT = "sky-pixels/s"

WORKLOADS = {
    # name: (N, n, m, T, E, description)
    "C1": (256, 4, 16, 1, 1, "C1: 256x256 sky, 4x4 PSF patches, margin 16, 1 energy bin"),
    "C2": (1024, 8, 64, 1, 3, "C2 Chandra-like: 1024x1024 sky, 3 energy bins, 8x8 PSF patches, margin 64, "
                              "exposure + mask + Poisson fwd+VJP"),
    "C3s": (1020, 10, 64, 2, 3, "C3-like slice: 1020x1020 sky, 2 modules x 3 bins, 10x10 PSF patches, margin 64"),
    # the other single-GPU configurations of BASELINE.json (parity-test cases; benched on request for DESIGN.md section 4)
    "C3": (1024, 10, 64, 7, 3, "C3 eROSITA-like: 7 telescope modules x 3 energy bins, 1024x1024 sky, 10x10 PSF patches, "
                               "margin 64 (P=332 -> L=336), per-module exposure/readout masks"),
    "C5g": (2048, 10, 128, 7, 3, "C5 geoVI batch, one GPU's share (2 of 16 samples with --samples-per-gpu 2): 7 telescope modules "
                                 "x 3 energy bins at 2048x2048, 10x10 PSF patches, margin 128 (P=664 -> L=672)"),
    "C4": (4096, 16, 64, 1, 1, "C4 JWST-like: 4096x4096 sky, 16x16 PSF patches with 128-px kernels (margin 64, P=640), "
                               "single filter"),
}


def make_inputs(wl, S, seed_shift=0):
    from tests import synth
    N, n, m, T, E, _ = WORKLOADS[wl]
    dx = N // n
    kern = synth.psf_cropped(T, n, E, dx, half_support=m)
    expo = synth.exposure(T, E, N)
    x = synth.sky(S, E, N, seed=0 + seed_shift)
    return kern, expo, x


def alg_bytes_flops(N, n, m, T, E, S):
    """compulsory HBM bytes and textbook FFT flops of one fused fwd+VJP evaluation."""
    dx = N // n
    P = 2 * dx + 2 * m
    khat = n * n * P * (P // 2 + 1) * 16
    b = S * E * N * N * 16 + S * T * E * (16 * N * N + 2 * khat)
    f = S * T * E * n * n * 2 * (2 * 2.5 * P * P * np.log2(P * P) + 6 * P * (P // 2 + 1))
    return float(b), float(f)


# ----------------------------------------------------------------------------------- CPU arm
def cpu_arm(wl, steps, warmup, bounded_bins=1):
    """Oracle (reference op sequence: c2c fftn x2 + ifftn, kernel prep every call, gather/scatter) on the host
    cores.  One step = fwd+VJP on `bounded_bins` energy bin(s) of the workload, T=1, S=1."""
    from oracle import jubik_oracle as orc
    from tests import synth
    N, n, m, T, E, desc = WORKLOADS[wl]
    Eb = min(bounded_bins, E)
    cores = os.cpu_count() or 1
    orc.set_workers(cores)
    dx = N // n
    kern = synth.psf_cropped(1, n, Eb, dx, half_support=m)
    expo = synth.exposure(1, Eb, N)
    x = synth.sky(1, Eb, N)[0]
    dom = orc.Domain((Eb, N, N), synth.DIST)
    R = orc.Response(dom, kern, n, m, expo, 500.0, synth.PIXEL_AREA, cropped=True)
    d = synth.counts_from(R.dense(synth.sky(1, Eb, N, seed=5)[0]))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.energy_and_grad(R, d, x)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    px = Eb * N * N
    tot = sum(times)
    return {"value": px * len(times) / tot, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of %d energy bin(s) of %s, T=1, S=1 per step; NumPy/scipy.fft oracle with workers=%d; "
                      "%d steps, %.2f s/step" % (Eb, E, wl, cores, len(times), tot / len(times))}, tot / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N, n, m, T, E, desc = WORKLOADS[args.workload]
    cb, s_per_step = cpu_arm(args.workload, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "N": N, "n_patches": n, "margin": m, "T": T, "E": E, "S": 1},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "app_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import jubik_b200 as jb
    from jubik_b200.plan import measure_fp64_tflops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything a library writes to fd 1 (NCCL prints its version banner there)
    # is sent to stderr instead, and the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        # the one collective is a ~25 MB fp64 all-reduce: on this NVSwitch box NCCL's ring beats its default (NVLS)
        # choice at that size (106 vs 124 us measured at 8 ranks, tools/ar_bench.py)
        os.environ.setdefault("NCCL_ALGO", "Ring")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, n, m, T, E, desc = WORKLOADS[args.workload]
    S = args.samples_per_gpu
    kern, expo, x_np = make_inputs(args.workload, S, seed_shift=100 * rank)
    dev = torch.device("cuda", local)
    kern_d = torch.as_tensor(kern).to(dev)
    expo_d = torch.as_tensor(expo).to(dev)
    plan = jb.ResponsePlan(kern_d, N, n, m, exposure=expo_d, exposure_cut=500.0, pixel_area=16.0,
                           max_samples=S, cropped=True)
    del kern_d
    x = torch.as_tensor(x_np).to(dev)
    truth = torch.as_tensor(make_inputs(args.workload, 1, seed_shift=5)[2]).to(dev)
    lam = plan.response(truth)[0]
    torch.manual_seed(3)
    plan.set_data(torch.poisson(lam.clamp(min=0.0)))
    del lam, truth
    flush = torch.empty(192 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 192 MiB > 126 MB L2

    # one flat buffer = [sample-summed sky gradient (E,N,N) | energies of all world*S samples]: the partial
    # log-likelihoods and the gradient are combined with a SINGLE NCCL all-reduce (north_star)
    npx = E * N * N
    flat = torch.zeros(npx + world * S, dtype=torch.float64, device=dev)
    gbuf = flat[:npx].view(1, E, N, N) if S == 1 else torch.empty((S, E, N, N), dtype=torch.float64, device=dev)
    ebuf = flat[npx + rank * S: npx + (rank + 1) * S]

    # default: the whole step -- kernels, the fold of the rank's samples into `flat` and the ONE ncclAllReduce -- is a
    # single C-ABI call (jubik_poisson_energy_and_grad_sharded) on one stream; the library owns the NCCL communicator
    # (torch.distributed only hands the ncclUniqueId around).  JUBIK_B200_TORCH_ALLREDUCE=1 keeps the collective in torch.
    via_torch = os.environ.get("JUBIK_B200_TORCH_ALLREDUCE") == "1"
    comm = None
    if world > 1 and not via_torch:
        comm = jb.Communicator.from_torch_distributed()
    glocal = gbuf if S > 1 else None     # per-sample gradients of this rank (S == 1: the gradient lands in flat directly)

    def step_device():
        if not via_torch:
            return plan.energy_and_grad_sharded(x, comm, jb.SHARD_SAMPLES, flat=flat, grad_local=glocal)
        if world > 1:
            flat[npx:].zero_()
        e, g = plan.energy_and_grad(x, out_energy=ebuf, out_grad=gbuf)
        if S > 1:
            torch.sum(g, 0, out=flat[:npx].view(E, N, N))
        if world > 1:
            dist.all_reduce(flat)
        return e, g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    launches0 = plan.launch_count
    clocks = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clocks.start()
    barrier()
    for a, b in ev:
        flush.zero_()                   # L2 flush between timed iterations (outside the events)
        a.record()
        step_device()
        b.record()
    barrier()
    clk = clocks.stop()
    launches = plan.launch_count - launches0
    t_dev = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev = float(tt)
    px_step = world * S * T * E * N * N
    value = px_step * args.steps / t_dev

    # ---- end to end through the host-buffer C-ABI call (pinned host memory)
    xh = torch.as_tensor(x_np).pin_memory()
    eh = torch.empty((S,), dtype=torch.float64).pin_memory()
    gh = torch.empty_like(xh).pin_memory()
    for _ in range(3):
        plan.energy_and_grad_host(xh, eh, gh)
    barrier()
    t0 = time.perf_counter()
    e2e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in e2e_ev:
        flush.zero_()
        a.record()
        plan.energy_and_grad_host(xh, eh, gh)
        if world > 1:
            et = eh.to(dev, non_blocking=True)
            dist.all_reduce(et)
        b.record()
    barrier()
    t_e2e = sum(a.elapsed_time(b) for a, b in e2e_ev) * 1e-3
    tt = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_e2e = float(tt)
    e2e_value = px_step * args.steps / t_e2e
    h2d = xh.numel() * 8
    d2h = gh.numel() * 8 + eh.numel() * 8

    line = None
    if rank == 0:
        # ---- per-kernel durations, live (CUDA events between launches on the launching stream)
        acc = {}
        for _ in range(max(3, min(args.steps, 10))):
            flush.zero_()
            _, _, ms = plan.profile(x)
            for k, v in ms.items():
                acc.setdefault(k, []).append(v)
        kms = {k: statistics.mean(v) for k, v in acc.items()}
        info = plan.info
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        fp64 = measure_fp64_tflops()
        L, P, PH, w = info["L"], info["P"], info["PH"], info["w"]
        # column pass, forward: read w rows, read the resident spectrum (L*PH), write P rows, all complex128
        cols_bytes = S * T * E * n * n * (w + P + L) * PH * 16
        cols_s = kms["cols_fwd"] * 1e-3
        # DRAM bytes of the same kernel from the committed `ncu --set full` capture (profiles/*_traffic.json)
        traffic = None
        try:
            import glob
            tf = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))[-1]
            tj = json.load(open(tf))
            if tj.get("workload") == args.workload and tj.get("samples", 1) == S:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        b_alg, f_alg = alg_bytes_flops(N, n, m, T, E, S)
        t_step = t_dev / args.steps
        roof = {"bound": "hbm", "kernel": "k_cols (forward column pass: FFT * K^ * IFFT)",
                "achieved": cols_bytes / cols_s / 1e9, "peak": hbm, "unit": "GB/s",
                "frac": cols_bytes / cols_s / 1e9 / hbm, "peak_source": hbm_src, "traffic": traffic,
                "algorithmic_bytes_per_launch": cols_bytes, "launch_ms": kms["cols_fwd"],
                "kernel_ms": kms,
                "path": {"B_alg_bytes": b_alg, "F_alg_flops": f_alg,
                         "frac_hbm": b_alg / t_step / 1e9 / hbm,
                         "fp64_peak_tflops_measured": fp64,
                         "frac_fp64": f_alg / t_step / 1e12 / fp64 if fp64 > 0 else None,
                         "binding": "fp64" if f_alg / (fp64 * 1e12) > b_alg / (hbm * 1e9) else "hbm"}}
        cb, _ = cpu_arm(args.workload, 3, 1) if not args.no_cpu else ({"value": None, "unit": UNIT, "cores": 0,
                                                                      "kind": "port", "sample": "skipped"}, 0)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "N": N, "n_patches": n, "margin": m, "T": T, "E": E,
                           "samples_per_gpu": S, "fft_length": L, "parallelism": "sample-sharded x%d" % world,
                           "collective": ("torch.distributed all_reduce" if via_torch else
                                          "one ncclAllReduce of [sum-grad | energies] issued inside the C-ABI call")
                           if world > 1 else "none",
                           "l2": "flushed (192 MiB write) between timed iterations; working set %.0f MB > L2"
                                 % (info["bytes_total"] / 1e6)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": t_e2e / args.steps * 1e3},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cb}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        if comm is not None:
            comm.close()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--samples-per-gpu", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
