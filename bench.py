#!/usr/bin/env python
"""bench.py -- fused forward + VJP of response + Poisson log-likelihood, fp64, in sky-pixels/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C5|C1|C2|C3|C4]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU)

Default workload: BASELINE.json configs[4] ("C5", the configuration the metric "... at 1/2/4/8 B200" is quoted on, and
the largest one): 16 posterior samples x 7 telescope modules x 3 energy bins at 2048x2048, 10x10 PSF patches, margin 128.
The job is FIXED (strong scaling): N ranks evaluate 16/N samples each, in chunks of two (the plan's workspace holds two
samples; one GPU loops over 8 chunks).  One step = ONE C-ABI call per rank (jubik_poisson_energy_and_grad_sharded): the
kernels of every chunk, the fold of the rank's samples into a flat buffer [sample-summed gradient | energies of all
samples] and the NCCL all-reduce of that buffer (issued by the library, one energy bin at a time so that it overlaps
the next bin's kernels).  sky-pixels per step = samples * T * E * N^2.  --workload selects the other
BASELINE configurations (C1..C4, their own sample counts).

`value`  : device-resident throughput (inputs already in HBM), CUDA events, max over ranks.
`e2e`    : the same operation through jubik_poisson_energy_and_grad_sharded_host with pinned HOST buffers: H2D of every
           sample and D2H of [gradient | energies] inside the timed region.
`roofline`: SURVEY 8(d) whole-path fraction against the binding roof (FP64 pipe, measured DFMA peak) with the HBM
           fraction beside it; `dominant_kernel` = the slower of the two column passes against its compulsory bytes (the
           resident PSF spectra it must read), `traffic` = its DRAM bytes per launch from the committed ncu capture of
           the same workload (profiles/r2*_traffic.json); per-kernel durations measured live with CUDA events.
`parity` : N = 1: one (module, bin) plane of the workload against the CPU oracle; N > 1: rank 0 re-evaluates all samples
           alone and compares with the all-reduced result.
`cpu_baseline`: the NumPy/scipy oracle (the reference's op sequence; jax is not installable here) timed on this box's
           host cores on one (module, bin) plane of the workload (N = 1 only).  --impl reference prints that arm alone.
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

# torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core
if "--impl" in sys.argv and "reference" in sys.argv:
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402

METRIC = "fwd+VJP response+loglik sky-pixels/s (fp64)"
UNIT = "sky-pixels/s"

WORKLOADS = {
    # name: (N, n, m, T, E, samples of the job, samples per chunk, description)
    "C1": (256, 4, 16, 1, 1, 1, 1, "C1: 256x256 sky, 4x4 PSF patches, margin 16, 1 energy bin"),
    "C2": (1024, 8, 64, 1, 3, 2, 2, "C2 Chandra-like: 1024x1024 sky, 3 energy bins, 8x8 PSF patches, margin 64, "
                                    "exposure + mask + Poisson fwd+VJP, 2 samples"),
    "C3": (1024, 10, 64, 7, 3, 1, 1, "C3 eROSITA-like: 7 telescope modules x 3 energy bins, 1024x1024 sky, 10x10 PSF "
                                     "patches, margin 64 (P=332 -> L=336), per-module exposure/readout masks"),
    "C4": (4096, 16, 64, 1, 1, 1, 1, "C4 JWST-like: 4096x4096 sky, 16x16 PSF patches with 128-px kernels (margin 64, "
                                     "P=640), single filter"),
    "C5": (2048, 10, 128, 7, 3, 16, 2, "C5 geoVI batch: 16 posterior samples x 7 telescope modules x 3 energy bins at "
                                       "2048x2048, 10x10 PSF patches, margin 128 (P=664 -> L=672), sample-sharded"),
}


def config_of(wl):
    """The workload keys both arms print (identical dicts: the driver's same_config check)."""
    N, n, m, T, E, S, chunk, desc = WORKLOADS[wl]
    return {"workload": desc, "name": wl, "N": N, "n_patches": n, "margin": m, "T": T, "E": E, "samples": S}


def alg_bytes_flops(N, n, m, T, E, S):
    """compulsory HBM bytes and textbook FFT flops of one fused fwd+VJP evaluation of S samples."""
    dx = N // n
    P = 2 * dx + 2 * m
    khat = n * n * P * (P // 2 + 1) * 16
    b = S * E * N * N * 16 + S * T * E * (16 * N * N + 2 * khat)
    f = S * T * E * n * n * 2 * (2 * 2.5 * P * P * np.log2(P * P) + 6 * P * (P // 2 + 1))
    return float(b), float(f)


# ----------------------------------------------------------------------------------- CPU arm
class CpuPlane:
    """One (module 0, bin 0) plane of the workload at full image size on the oracle (reference op sequence: c2c fftn
    x2 + ifftn, kernel prep every call, gather/scatter) -- the bounded sample of the CPU arm, and the parity probe."""

    def __init__(self, wl):
        from oracle import jubik_oracle as orc
        from tests import synth
        self.orc = orc
        N, n, m, T, E, S, chunk, desc = WORKLOADS[wl]
        self.wl, self.N, self.n, self.m = wl, N, n, m
        self.cores = os.cpu_count() or 1
        orc.set_workers(self.cores)
        dx = N // n
        self.kern = synth.psf_cropped(1, n, 1, dx, half_support=m)
        self.expo = synth.exposure(1, 1, N)
        self.x = synth.sky(1, 1, N)[0]
        dom = orc.Domain((1, N, N), synth.DIST)
        self.R = orc.Response(dom, self.kern, n, m, self.expo, 500.0, synth.PIXEL_AREA, cropped=True)
        self.d = synth.counts_from(self.R.dense(synth.sky(1, 1, N, seed=5)[0]))
        self.sample = "one (module, energy bin) plane of %s at %dx%d (of T=%d x E=%d x %d samples) per step" % (
            wl, N, N, T, E, S)

    def step(self):
        return self.orc.energy_and_grad(self.R, self.d, self.x)

    def time(self, steps, warmup):
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            self.step()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        tot = sum(times)
        px = self.N * self.N
        cb = {"value": px * len(times) / tot, "unit": UNIT, "cores": self.cores, "kind": "port",
              "sample": self.sample + "; NumPy/scipy.fft oracle with workers=%d; %d steps, %.2f s/step"
                        % (self.cores, len(times), tot / len(times))}
        return cb, tot / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cb, s_per_step = CpuPlane(args.workload).time(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args.workload),
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
def make_inputs(wl):
    from tests import synth
    N, n, m, T, E, S, chunk, _ = WORKLOADS[wl]
    return synth.psf_cropped(T, n, E, N // n, half_support=m), synth.exposure(T, E, N)


def sample_skies(wl, first, count):
    """Skies of the job's samples [first, first+count): one seed per global sample index."""
    from tests import synth
    N, n, m, T, E, S, chunk, _ = WORKLOADS[wl]
    return np.concatenate([synth.sky(1, E, N, seed=1000 + s) for s in range(first, first + count)], axis=0)


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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, n, m, T, E, S_job, chunk, desc = WORKLOADS[args.workload]
    if args.samples:
        S_job = args.samples
    if S_job % world:
        raise SystemExit("the job's %d samples do not divide over %d ranks" % (S_job, world))
    S = S_job // world                   # this rank's samples
    chunk = min(chunk, S)
    dev = torch.device("cuda", local)
    kern, expo = make_inputs(args.workload)
    kern_d = torch.as_tensor(kern).to(dev)
    expo_d = torch.as_tensor(expo).to(dev)
    del kern, expo
    plan = jb.ResponsePlan(kern_d, N, n, m, exposure=expo_d, exposure_cut=500.0, pixel_area=16.0,
                           max_samples=chunk, cropped=True, host_call=True, sharded=True, metric=False)
    del kern_d
    from tests import synth
    truth = torch.as_tensor(synth.sky(1, E, N, seed=5)).to(dev)
    lam = plan.response(truth)[0]
    torch.manual_seed(3)
    plan.set_data(torch.poisson(lam.clamp(min=0.0)))
    del lam, truth
    x_np = sample_skies(args.workload, rank * S, S)
    x = torch.as_tensor(x_np).to(dev)
    flush = torch.empty(192 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 192 MiB > 126 MB L2

    # one flat buffer = [sample-summed sky gradient (E,N,N) | energies of all world*S samples]: the partial
    # log-likelihoods and the gradient are combined by the library's NCCL all-reduce (north_star)
    npx = E * N * N
    flat = torch.zeros(npx + world * S, dtype=torch.float64, device=dev)
    comm = jb.Communicator.from_torch_distributed() if world > 1 else None

    def step_device():
        return plan.energy_and_grad_sharded(x, comm, jb.SHARD_SAMPLES, flat=flat)

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
    px_step = S_job * T * E * N * N
    value = px_step * args.steps / t_dev
    res_dev = flat.clone()

    # ---- end to end through the host-buffer C-ABI call (pinned host memory): the same operation, same reduction
    xh = torch.as_tensor(x_np).pin_memory()
    fh = torch.empty((npx + world * S,), dtype=torch.float64).pin_memory()
    for _ in range(3):
        plan.energy_and_grad_sharded_host(xh, comm, fh)
    barrier()
    e2e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in e2e_ev:
        flush.zero_()
        a.record()
        plan.energy_and_grad_sharded_host(xh, comm, fh)
        b.record()
    barrier()
    t_e2e = sum(a.elapsed_time(b) for a, b in e2e_ev) * 1e-3
    tt = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_e2e = float(tt)
    e2e_value = px_step * args.steps / t_e2e
    h2d = xh.numel() * 8 * world
    d2h = fh.numel() * 8 * world
    e2e_same = bool(torch.equal(fh, res_dev.cpu()))   # host form == device form, bit for bit

    # ---- parity leg
    parity = {"checked": False}
    cpu = None
    if world > 1:
        # rank 0 evaluates ALL samples alone (chunked, no communicator) and compares with the all-reduced result
        if rank == 0:
            xa = torch.as_tensor(sample_skies(args.workload, 0, S_job)).to(dev)
            fl1 = torch.zeros(npx + S_job, dtype=torch.float64, device=dev)
            plan.energy_and_grad_sharded(xa, None, jb.SHARD_SAMPLES, flat=fl1)
            g_err = float((fl1[:npx] - res_dev[:npx]).abs().max() / fl1[:npx].abs().max())
            e_err = float(((fl1[npx:] - res_dev[npx:]) / fl1[npx:]).abs().max())
            parity = {"checked": True, "ok": bool(g_err < 1e-11 and e_err < 1e-12 and e2e_same),
                      "what": "all-reduced [sum-gradient | energies] of %d ranks vs one rank evaluating all %d samples; "
                              "host-buffer form bitwise equal to the device form: %s" % (world, S_job, e2e_same),
                      "max_rel_err_grad": g_err, "max_rel_err_energy": e_err}
            del xa, fl1
    elif not args.no_cpu:
        cpu = CpuPlane(args.workload)
        p1 = jb.ResponsePlan(torch.as_tensor(cpu.kern).to(dev), N, n, m, exposure=torch.as_tensor(cpu.expo).to(dev),
                             exposure_cut=500.0, pixel_area=16.0, data=torch.as_tensor(cpu.d).to(dev), cropped=True,
                             metric=False)
        e1, g1 = p1.energy_and_grad(torch.as_tensor(cpu.x).to(dev))
        eo, go = cpu.step()
        g_err = float(np.abs(g1.cpu().numpy() - go).max() / np.abs(go).max())
        e_err = abs(float(e1) - eo) / abs(eo)
        parity = {"checked": True, "ok": bool(g_err < 1e-10 and e_err < 1e-10 and e2e_same),
                  "what": "energy + gradient of one (module, bin) plane of the workload at full size vs the CPU oracle; "
                          "host-buffer form bitwise equal to the device form: %s" % e2e_same,
                  "max_rel_err_grad": g_err, "max_rel_err_energy": e_err}
        p1.close()

    if rank == 0:
        # ---- per-kernel durations of ONE chunk, live (CUDA events between launches on the launching stream)
        acc = {}
        for _ in range(max(3, min(args.steps, 10))):
            flush.zero_()
            _, _, ms = plan.profile(x[:chunk])
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
        b_alg, f_alg = alg_bytes_flops(N, n, m, T, E, S_job)
        t_step = t_dev / args.steps
        ach_f = f_alg / t_step / 1e12 / world          # per GPU
        ach_b = b_alg / t_step / 1e9 / world
        binding = "fp64" if f_alg / (fp64 * 1e12) > b_alg / (hbm * 1e9) else "hbm"
        # dominant kernel = the slower of the two column passes (same compulsory bytes and textbook flops); compulsory
        # bytes = the resident PSF spectra it has to read, once per sample (no intermediate counted),
        # P*(P/2+1) complex per patch
        khat_b = chunk * T * E * n * n * P * (P // 2 + 1) * 16
        dom = "cols_adj" if kms.get("cols_adj", 0.0) > kms["cols_fwd"] else "cols_fwd"
        cols_s = kms[dom] * 1e-3
        if T > 1 and info.get("fast_path"):
            dom_name = {"cols_fwd": "k_cols3_fwd_mt (forward column pass, forward transform shared by the modules: FFT, "
                                    "per module * K^ and IFFT)",
                        "cols_adj": "k_cols3_adj_mt (adjoint column pass, modules summed in the spectral domain: per module "
                                    "FFT * conj K^, one IFFT)"}[dom]
        else:
            dom_name = {"cols_fwd": "k_cols3_pipe (forward column pass: FFT * K^ * IFFT)",
                        "cols_adj": "k_cols3_pipe (adjoint column pass: FFT * conj K^ * IFFT)"}[dom]
        # textbook flops of that launch: forward + inverse column FFTs over PH columns + the spectrum multiply
        cols_f = chunk * T * E * n * n * PH * (2 * 5 * P * np.log2(P) + 6 * P)
        # DRAM bytes of that kernel per launch from the committed `ncu --set full` capture of the same workload and chunk
        # (profiles/r2*_traffic.json; null when no capture of this workload / kernel exists)
        traffic_ncu, traffic = None, None
        try:
            import glob
            tf = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2*_traffic.json")))[-1]
            tj = json.load(open(tf))
            if tj.get("workload") == args.workload and tj.get("samples") == chunk and dom in tj.get("kernels", {}):
                traffic_ncu = dict(tj["kernels"][dom], file=os.path.relpath(tf, ROOT), source=tj.get("source"))
                traffic = traffic_ncu["dram_bytes_read"] + traffic_ncu["dram_bytes_write"]
        except Exception:
            pass
        roof = {"bound": binding, "what": "whole path, SURVEY 8(d): textbook FFT flops (compulsory bytes) of one step / "
                                          "device time, per GPU",
                "achieved": ach_f if binding == "fp64" else ach_b,
                "peak": fp64 if binding == "fp64" else hbm,
                "unit": "TFLOP/s" if binding == "fp64" else "GB/s",
                "frac": (ach_f / fp64) if binding == "fp64" else (ach_b / hbm),
                "peak_source": "DFMA micro-benchmark run inside this process (jubik_measure_fp64_tflops)"
                               if binding == "fp64" else hbm_src,
                "frac_hbm": ach_b / hbm, "achieved_hbm_gbs": ach_b, "peak_hbm_gbs": hbm, "peak_hbm_source": hbm_src,
                "frac_fp64": ach_f / fp64 if fp64 > 0 else None, "achieved_fp64_tflops": ach_f, "peak_fp64_tflops": fp64,
                "B_alg_bytes": b_alg, "F_alg_flops": f_alg,
                "traffic": traffic, "traffic_ncu_profile": traffic_ncu,
                "dominant_kernel": {"name": "%s, one chunk of %d samples" % (dom_name, chunk),
                                    "launch_ms": kms[dom],
                                    "compulsory_bytes_per_launch": khat_b,
                                    "achieved_gbs": khat_b / cols_s / 1e9, "frac_hbm": khat_b / cols_s / 1e9 / hbm,
                                    "textbook_flops_per_launch": cols_f,
                                    "achieved_tflops": cols_f / cols_s / 1e12,
                                    "frac_fp64": cols_f / cols_s / 1e12 / fp64 if fp64 > 0 else None},
                "kernel_ms_one_chunk": kms}
        if cpu is not None:
            cb, _ = cpu.time(3, 1)
        else:
            cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                  "sample": "not run (N > 1 or --no-cpu): the CPU arm is timed at N = 1 only"}
        cfg = config_of(args.workload)
        cfg["samples"] = S_job
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "detail": {"samples_per_gpu": S, "samples_per_chunk": chunk, "fft_length": L,
                           "parallelism": "sample-sharded x%d" % world,
                           "collective": "ncclAllReduce of [sum-grad | energies] issued inside the C-ABI call, one energy "
                                         "bin at a time on a second stream (overlaps the next bin's kernels)"
                           if world > 1 else "none",
                           "l2": "flushed (192 MiB write) between timed iterations; working set %.0f MB > L2"
                                 % (info["bytes_total"] / 1e6)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": t_e2e / args.steps * 1e3,
                        "call": "jubik_poisson_energy_and_grad_sharded_host (pinned host x -> host [sum-grad | energies])"},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cb, "parity": parity,
                "parity_ok": parity.get("ok")}
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
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="override the job's sample count")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / oracle parity leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
