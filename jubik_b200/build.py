"""Build libjubik_b200.so (sm_100a) in-tree with nvcc.

One object per FFT length listed in csrc/fft_list.h (compiled in parallel) plus csrc/capi.cu.
Usage:  python -m jubik_b200.build [--force] [--jobs N] [--verbose]
The same recipe is what __graft_entry__.build() runs.  No torch extension machinery is involved:
the product is a plain C-ABI shared library (include/jubik_b200.h).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libjubik_b200.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xptxas", "-v",
    "-Xcudafe", "--diag_suppress=177",
]


def fft_list():
    txt = open(os.path.join(CSRC, "fft_list.h")).read()
    return [tuple(int(v) for v in m.groups())
            for m in re.finditer(r"X\((\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\)", txt)]


STAMP = os.path.join(OBJ, "sources.sha256")


def _sources_digest(flags):
    """sha256 over every source/header that goes into the library plus the compiler flags: the rebuild decision
    follows the CONTENT (an mtime comparison breaks on checkouts and on the copy that travels to the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f == "jax_ffi.cc":          # the jax.ffi shim is a separate library (jax_binding.build_ffi_library)
                continue
            h.update(f.encode())
            with open(os.path.join(root, f), "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout[-4000:]))
    return r.stdout


def build(force=False, jobs=None, verbose=False):
    extra = os.environ.get("JUBIK_B200_NVCC_EXTRA", "").split()   # experiment hook, e.g. "-DJB_R5_SW=1"
    if extra:
        NVCC_FLAGS.extend(f for f in extra if f not in NVCC_FLAGS)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: jubik_b200 needs the CUDA toolkit to build")
    digest = _sources_digest(NVCC_FLAGS)
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    # one build at a time: concurrent callers (e.g. several tools/build_variant.py processes) would write the same objects
    import fcntl
    lock = open(os.path.join(OBJ, ".lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB          # another process built it while this one waited
    jobs = jobs or os.cpu_count() or 4
    tasks = []
    for (L, r1, r2, cb, pipe, q1, q2, q3) in fft_list():
        obj = os.path.join(OBJ, "fft_%d.o" % L)
        cmd = [nvcc] + NVCC_FLAGS + ["-DJB_INST_L=%d" % L, "-DJB_INST_R1=%d" % r1, "-DJB_INST_R2=%d" % r2,
                                    "-DJB_INST_CB=%d" % cb, "-DJB_INST_PIPE=%d" % pipe, "-DJB_INST_Q1=%d" % q1, "-DJB_INST_Q2=%d" % q2, "-DJB_INST_Q3=%d" % q3, "-c", os.path.join(CSRC, "fft_inst.cu"), "-o", obj]
        tasks.append((cmd, obj))
    obj = os.path.join(OBJ, "capi.o")
    tasks.append(([nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, "capi.cu"), "-o", obj], obj))
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        futs = {ex.submit(_run, cmd, o + ".log"): o for cmd, o in tasks}
        for f in cf.as_completed(futs):
            out = f.result()
            if verbose:
                print(out)
    objs = [o for _, o in tasks]
    _run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"], os.path.join(OBJ, "link.log"))
    with open(STAMP, "w") as fh:
        fh.write(digest + "\n")
    return LIB


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.jobs, a.verbose))


if __name__ == "__main__":
    sys.exit(main())
