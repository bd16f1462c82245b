"""Build a tuning variant of the CUDA library: recompile ONE FFT length with extra -D flags and link it with capi.o
and null stubs for every other length (a variant serves only geometries of its own length; ~10 MB instead of ~100 MB,
the gpurun snapshot is capped) into jubik_b200/_variants/lib_<tag>.so (loaded via JUBIK_B200_LIB=...).

    python tools/build_variant.py <tag> <L> [--cb=2] [--q=8,12,7] -DJB_R5_B=6 -DJB_R5_CPS=5
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from jubik_b200 import build as b  # noqa: E402

tag, L, extra = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
cb_override = None
q_override = None
for a in list(extra):
    if a.startswith("--cb="):
        cb_override = int(a[5:])
        extra.remove(a)
    if a.startswith("--rq="):         # factor order of the c2r row kernels, e.g. --rq=8,7,12
        extra += ["-DJB_INST_RQ%d=%s" % (i + 1, v) for i, v in enumerate(a[5:].split(","))]
        extra.remove(a)
    if a.startswith("--q="):          # three-stage factor order, e.g. --q=8,12,7
        q_override = tuple(int(v) for v in a[4:].split(","))
        extra.remove(a)
b.build()
out = os.path.join(b.HERE, "_variants")
os.makedirs(out, exist_ok=True)
objs = []
stubs = "struct JbFftOps;\n"
for (l, r1, r2, cb, pipe, q1, q2, q3) in b.fft_list():
    o = None
    if l != L:
        stubs += "const JbFftOps* jb_fft_ops_%d() { return nullptr; }\n" % l
    if l == L:
        if cb_override:
            cb = cb_override
        if q_override:
            q1, q2, q3 = q_override
        o = os.path.join(out, "fft_%d_%s.o" % (l, tag))
        cmd = ["nvcc"] + b.NVCC_FLAGS + extra + ["-DJB_INST_L=%d" % l, "-DJB_INST_R1=%d" % r1, "-DJB_INST_R2=%d" % r2,
                                                "-DJB_INST_CB=%d" % cb, "-DJB_INST_PIPE=%d" % pipe, "-DJB_INST_Q1=%d" % q1,
                                                "-DJB_INST_Q2=%d" % q2, "-DJB_INST_Q3=%d" % q3, "-c",
                                                os.path.join(b.CSRC, "fft_inst.cu"), "-o", o]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            sys.exit(r.stdout[-4000:])
        for line in r.stdout.splitlines():
            if "k_rows_c2r5" in line or "k_cols3_" in line:
                print(line[:160])
            if "spill" in line and "0 bytes spill stores" not in line:
                print("   ", line.strip())
    if o:
        objs.append(o)
stub_c, stub_o = os.path.join(out, "stubs_%d.cpp" % L), os.path.join(out, "stubs_%d.o" % L)
open(stub_c, "w").write(stubs)
subprocess.run(["g++", "-fPIC", "-c", stub_c, "-o", stub_o], check=True)
objs += [stub_o, os.path.join(b.OBJ, "capi.o")]
lib = os.path.join(out, "lib_%s.so" % tag)
subprocess.run(["nvcc", "-shared", "-o", lib] + objs + ["-lcudart"], check=True)
print(lib)
