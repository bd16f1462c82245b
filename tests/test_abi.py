"""The C-ABI shared library: builds for sm_100a without a GPU, exports every symbol include/jubik_b200.h
declares, and the product refuses to run without a device (no CPU fallback, no oracle import)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib_path():
    from jubik_b200 import build
    return build.build(force=False)


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "jubik_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(jubik_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), name
    from jubik_b200 import _lib
    assert sorted(_lib.EXPORTS) == names


def test_library_contains_sm100a_code(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out


def test_supported_lengths_and_versions(lib_path):
    from jubik_b200 import _lib
    lib = _lib.load(require_device=False)
    assert lib.jubik_version() >= 200
    for L in (160, 336, 384, 640, 672):     # P of BASELINE configs C1..C5 (332 -> 336, 664 -> 672)
        assert lib.jubik_fft_length_supported(L) == 1
    for L in (192, 320):                    # the reference's shipped demos: 512^2 / 8x8 and 4x4, margfrac 0.062
        assert lib.jubik_fft_length_supported(L) == 1     # (demos/configs/eROSITA_demo.yaml:35-36,48; chandra_demo.yaml:13,44-45)
    for L in (96, 128, 256, 448, 512, 768, 896, 1024):
        assert lib.jubik_fft_length_supported(L) == 1
    assert lib.jubik_fft_length_supported(332) == 0


def test_kernel_selection_reads_no_environment():
    """The plan is a pure function of its descriptor: no getenv on the product path.  The one
    environment hook left names the NCCL library file to dlopen."""
    csrc = os.path.join(ROOT, "jubik_b200", "csrc")
    hits = []
    for f in os.listdir(csrc):
        for ln, line in enumerate(open(os.path.join(csrc, f)), 1):
            if "getenv" in line:
                hits.append((f, ln, line.strip()))
    assert all("JUBIK_B200_NCCL_LIB" in h[2] for h in hits), hits


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_fails_loudly_without_gpu(lib_path):
    import jubik_b200 as jb
    from jubik_b200 import _lib
    from jubik_b200.plan import make_desc, RawPlan
    with pytest.raises(jb.JubikError, match="no CUDA device"):
        jb.load_library(require_device=True)
    lib = _lib.load(require_device=False)
    with pytest.raises(jb.JubikError, match="no CPU fallback"):
        RawPlan(lib, make_desc(32, 1, 1, 4, 4, 1, False), 0)
    import torch
    with pytest.raises((TypeError, RuntimeError)):
        jb.ResponsePlan(torch.ones((1, 16, 1, 32, 32), dtype=torch.float64), 32, 4, 4)


def test_product_package_never_touches_oracle_or_emulation():
    pkg = os.path.join(ROOT, "jubik_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "libjubik_emu" not in txt, f
    code = "import sys; import jubik_b200; assert not any(m.startswith('oracle') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_integrate_mirror_known_answer():
    """jubik.convolve.integrate (convolve.py:179-184; reference test test_convolve.py:23-29) in the host mirror."""
    import numpy as np
    import torch
    import jubik_b200 as jb
    dom = jb.Domain(shape=(3, 4), distances=(0.5, 0.25))
    assert jb.integrate(np.ones((3, 4)), dom, axes=(0, 1)) == 3 * 4 * 0.5 * 0.25
    t = jb.integrate(torch.ones((2, 3, 4), dtype=torch.float64), jb.Domain((2, 3, 4), (1.0, 0.5, 0.25)), axes=(-2, -1))
    assert t.shape == (2,) and float(t[0]) == 3 * 4 * 0.5 * 0.25


def test_hot_kernels_use_tma_bulk_copies_and_fp64_pipe(lib_path):
    """SASS evidence (B200_PROFILING.md): the column pass is fed by TMA tensor loads and bulk copies on mbarriers
    (UTMALDG / UBLKCP / SYNCS), the r2c row kernels prefetch with cp.async (LDGSTS), and the arithmetic is FP64."""
    obj = os.path.join(ROOT, "jubik_b200", "_build", "fft_384.o")
    assert os.path.exists(obj)
    sass = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    for mnemonic in ("UTMALDG.3D", "UBLKCP", "SYNCS.ARRIVE.TRANS64", "LDGSTS.E.BYPASS.128", "DFMA", "DADD"):
        assert mnemonic in sass, mnemonic
    assert "HMMA" not in sass and "UTCHMMA" not in sass      # no tensor-core path: FP64 FFT / stencil work


def test_factor_orders_of_fft_list_are_consistent():
    """csrc/fft_list.h: every three-stage length factors exactly (column pass: X(... Q1, Q2, Q3); c2r rows: JB_ROWS_F3_LIST),
    the row list only names compiled pipeline lengths, and the module-sharing kernels' register paths (R3 <= 7) are what
    BASELINE's multi-module configurations (C3: L = 336, C5: L = 672) run."""
    import re
    from jubik_b200.build import fft_list, CSRC
    cols = {L: (pipe, q1, q2, q3) for (L, r1, r2, cb, pipe, q1, q2, q3) in fft_list()}
    for L, (pipe, q1, q2, q3) in cols.items():
        if pipe == 2:
            assert q1 * q2 * q3 == L, L
    txt = open(os.path.join(CSRC, "fft_list.h")).read()
    rows = [tuple(int(v) for v in m.groups()) for m in re.finditer(r"Y\((\d+),\s*(\d+),\s*(\d+),\s*(\d+)\)", txt)]
    assert rows, "JB_ROWS_F3_LIST not found"
    for (L, a, b, c) in rows:
        assert L in cols and cols[L][0] == 2 and a * b * c == L, L
    for L in (336, 672):
        assert cols[L][3] == 7 and cols[L][3] % 2 == 1      # odd last radix: conflict-free stage 3, register accumulator


def test_module_sharing_kernels_use_named_group_barriers(lib_path):
    """SASS evidence: the k1-group barriers of the module-sharing column kernels are hardware named barriers with
    immediate ids (BAR.SYNC with a non-zero id), and ptxas reserved only the groups' barriers, not all 16."""
    log = os.path.join(ROOT, "jubik_b200", "_build", "fft_672.o.log")
    assert os.path.exists(log)
    txt = open(log).read()
    i = txt.index("Function properties for _Z14k_cols3_adj_mt")
    used = txt[i:txt.index("\n", txt.index("Used", i))]
    assert "used 5 barriers" in used, used
