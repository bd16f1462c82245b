"""Host-thread EMULATION build of the CUDA kernels -- test infrastructure only.

Compiles jubik_b200/csrc/*.cu with g++ and -DJB_EMU (threads of a CTA become std::threads, __syncthreads a
std::barrier, device memory plain malloc) into tests/emu/_build/libjubik_emu.so.  It exists so that the index
logic of the kernels can be checked against the oracle on a box without a GPU (`pytest -m "not gpu"`).  The
jubik_b200 package never loads it: the product path is libjubik_b200.so on a CUDA device and nothing else.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from jubik_b200.build import fft_list, CSRC  # noqa: E402

EXTRA = os.environ.get("JUBIK_B200_EMU_EXTRA", "").split()     # kernel-variant switches, e.g. -DJB_COLS_WP=1 (tuning only)
OUT = os.path.join(HERE, "_build" + ("_" + "".join(c for c in "".join(EXTRA) if c.isalnum()) if EXTRA else ""))
LIB = os.path.join(OUT, "libjubik_emu.so")
# -fno-gnu-unique / -Bsymbolic: keep this library's symbols from merging with libjubik_b200.so when a test
# process loads both (they are built from the same sources)
FLAGS = ["-std=c++20", "-O1", "-fPIC", "-DJB_EMU", "-x", "c++", "-pthread", "-w", "-fno-gnu-unique",
         "-fvisibility-inlines-hidden"] + EXTRA


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError(" ".join(cmd) + "\n" + r.stdout[-6000:])


def build(force=False, lengths=None):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "emu_ctx.cpp")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(s) for s in srcs):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    tasks = []
    for (L, r1, r2, cb, pipe, q1, q2, q3) in fft_list():
        o = os.path.join(OUT, "fft_%d.o" % L)
        tasks.append((["g++"] + FLAGS + ["-DJB_INST_L=%d" % L, "-DJB_INST_R1=%d" % r1, "-DJB_INST_R2=%d" % r2,
                                         "-DJB_INST_CB=%d" % cb, "-DJB_INST_PIPE=%d" % pipe, "-DJB_INST_Q1=%d" % q1, "-DJB_INST_Q2=%d" % q2, "-DJB_INST_Q3=%d" % q3, "-c", os.path.join(CSRC, "fft_inst.cu"), "-o", o], o))
    for name in ("capi.cu",):
        o = os.path.join(OUT, name + ".o")
        tasks.append((["g++"] + FLAGS + ["-c", os.path.join(CSRC, name), "-o", o], o))
    o = os.path.join(OUT, "emu_ctx.o")
    tasks.append((["g++", "-std=c++20", "-O1", "-fPIC", "-DJB_EMU", "-I", CSRC, "-c", os.path.join(HERE, "emu_ctx.cpp"), "-o", o], o))
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for f in [ex.submit(_run, c) for c, _ in tasks]:
            f.result()
    _run(["g++", "-shared", "-Wl,-Bsymbolic", "-o", LIB] + [o for _, o in tasks] + ["-pthread"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
