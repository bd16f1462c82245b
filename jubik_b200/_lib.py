"""ctypes binding of the C ABI in include/jubik_b200.h.

`load()` returns the bound CUDA library or raises: there is no CPU fallback and no alternative
implementation behind this module.  `bind(cdll)` only attaches the prototypes and is reused by
tests/emu to drive the host-thread emulation build of the same sources (test infrastructure).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# JUBIK_B200_LIB: development hook to load a differently tuned build of the same CUDA library (tools/build_variant.py)
LIB_PATH = os.environ.get("JUBIK_B200_LIB") or os.path.join(HERE, "libjubik_b200.so")

KERNEL_FULL, KERNEL_CROPPED = 0, 1
DATA_F64, DATA_I64 = 0, 1
PLAN_HOST_CALL, PLAN_METRIC, PLAN_SHARDED = 1, 2, 4

EXPORTS = [
    "jubik_last_error", "jubik_version", "jubik_fft_length_supported", "jubik_device_count",
    "jubik_plan_create", "jubik_plan_destroy", "jubik_plan_get_info", "jubik_plan_set_data", "jubik_plan_set_kernel",
    "jubik_plan_ensure_kernel", "jubik_stream_synchronize",
    "jubik_poisson_set_point", "jubik_poisson_metric_apply_at_point", "jubik_poisson_left_sqrt_metric_at_point",
    "jubik_plan_launch_count", "jubik_psf_apply", "jubik_psf_apply_T", "jubik_response_apply",
    "jubik_response_apply_T", "jubik_poisson_energy_and_grad", "jubik_poisson_metric_apply",
    "jubik_poisson_energy_and_grad_host", "jubik_poisson_profile", "jubik_measure_fp64_tflops",
    "jubik_poisson_left_sqrt_metric", "jubik_poisson_energy_and_grad_logsky",
    "jubik_poisson_energy_and_grad_logsky_padded",
    "jubik_slice_patches", "jubik_slice_patches_T", "jubik_convolve", "jubik_psf_interpolate",
    "jubik_comm_unique_id", "jubik_comm_create", "jubik_comm_adopt", "jubik_comm_destroy", "jubik_comm_rank",
    "jubik_comm_size", "jubik_sharded_flat_size", "jubik_poisson_energy_and_grad_sharded",
    "jubik_poisson_energy_and_grad_sharded_host",
]
SHARD_SAMPLES, SHARD_MODULES = 0, 1


class PlanDesc(C.Structure):
    _fields_ = [
        ("n0", C.c_int32), ("n1", C.c_int32), ("n_energy", C.c_int32), ("n_modules", C.c_int32),
        ("max_samples", C.c_int32), ("n_patches", C.c_int32), ("margin", C.c_int32),
        ("kernel_layout", C.c_int32), ("data_dtype", C.c_int32), ("has_exposure_cut", C.c_int32),
        ("fft_length", C.c_int32), ("flags", C.c_int32),
        ("d0", C.c_double), ("d1", C.c_double), ("exposure_cut", C.c_double), ("pixel_area", C.c_double),
        ("kernel", C.c_void_p), ("exposure", C.c_void_p), ("eff_area", C.c_void_p), ("data", C.c_void_p),
    ]


class PsfTables(C.Structure):
    _fields_ = [("n_energy", C.c_int32), ("n_rad", C.c_int32), ("px", C.c_int32), ("py", C.c_int32),
                ("psfs", C.c_void_p), ("rs", C.c_void_p), ("center_ids", C.c_void_p),
                ("delta0", C.c_double), ("delta1", C.c_double), ("pointing0", C.c_double), ("pointing1", C.c_double)]


class PlanInfo(C.Structure):
    _fields_ = [(k, C.c_int32) for k in
                ("N", "n", "m", "dx", "w", "P", "L", "PH", "S_max", "T", "E", "exact_period",
                 "psf_half_support", "fft_r1", "fft_r2", "col_block", "alias_fold", "alias_ext", "fast_path", "flags")] + \
               [("bytes_khat", C.c_size_t), ("bytes_workspace", C.c_size_t), ("bytes_total", C.c_size_t)]


class JubikError(RuntimeError):
    pass


def bind(lib):
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    lib.jubik_last_error.restype = C.c_char_p
    lib.jubik_last_error.argtypes = []
    lib.jubik_version.restype = i
    lib.jubik_fft_length_supported.argtypes = [i]
    lib.jubik_device_count.restype = i
    lib.jubik_plan_create.argtypes = [C.POINTER(vp), C.POINTER(PlanDesc), vp]
    lib.jubik_plan_destroy.argtypes = [vp]
    lib.jubik_plan_get_info.argtypes = [vp, C.POINTER(PlanInfo)]
    lib.jubik_plan_set_data.argtypes = [vp, vp, i, vp]
    lib.jubik_plan_set_kernel.argtypes = [vp, vp, i, vp]
    lib.jubik_plan_ensure_kernel.argtypes = [vp, vp, i, C.POINTER(C.c_int), vp]
    lib.jubik_poisson_set_point.argtypes = [vp, vp, i, vp]
    lib.jubik_poisson_metric_apply_at_point.argtypes = [vp, vp, i, vp, vp]
    lib.jubik_poisson_left_sqrt_metric_at_point.argtypes = [vp, vp, i, vp, vp]
    lib.jubik_plan_launch_count.argtypes = [vp]
    lib.jubik_stream_synchronize.argtypes = [vp]
    lib.jubik_plan_launch_count.restype = C.c_longlong
    for name in ("jubik_psf_apply", "jubik_psf_apply_T", "jubik_response_apply", "jubik_response_apply_T"):
        getattr(lib, name).argtypes = [vp, vp, i, vp, vp]
    lib.jubik_poisson_energy_and_grad.argtypes = [vp, vp, i, vp, vp, vp]
    lib.jubik_poisson_metric_apply.argtypes = [vp, vp, vp, i, vp, vp]
    lib.jubik_poisson_energy_and_grad_host.argtypes = [vp, vp, i, vp, vp, vp]
    lib.jubik_poisson_profile.argtypes = [vp, vp, i, vp, vp, C.POINTER(C.c_float), C.POINTER(C.c_int), vp]
    lib.jubik_measure_fp64_tflops.argtypes = [C.POINTER(d), vp]
    lib.jubik_poisson_left_sqrt_metric.argtypes = [vp, vp, vp, i, vp, vp]
    lib.jubik_poisson_energy_and_grad_logsky.argtypes = [vp, vp, i, vp, vp, vp]
    lib.jubik_poisson_energy_and_grad_logsky_padded.argtypes = [vp, vp, i, i, i, vp, vp, vp]
    lib.jubik_slice_patches.argtypes = [vp, i, i, i, i, i, vp, vp]
    lib.jubik_slice_patches_T.argtypes = [vp, i, i, i, i, i, vp, vp]
    lib.jubik_convolve.argtypes = [vp, vp, i, i, i, i, d, vp, vp]
    lib.jubik_psf_interpolate.argtypes = [C.POINTER(PsfTables), i, i, d, d, i, vp, vp]
    lib.jubik_comm_unique_id.argtypes = [vp]
    lib.jubik_comm_create.argtypes = [C.POINTER(vp), vp, i, i]
    lib.jubik_comm_adopt.argtypes = [C.POINTER(vp), vp]
    lib.jubik_comm_destroy.argtypes = [vp]
    lib.jubik_comm_rank.argtypes = [vp]
    lib.jubik_comm_size.argtypes = [vp]
    lib.jubik_sharded_flat_size.argtypes = [vp, vp, i, i]
    lib.jubik_sharded_flat_size.restype = C.c_longlong
    lib.jubik_poisson_energy_and_grad_sharded.argtypes = [vp, vp, i, vp, i, vp, vp, vp]
    lib.jubik_poisson_energy_and_grad_sharded_host.argtypes = [vp, vp, vp, i, vp, vp]
    for name in EXPORTS:
        if name not in ("jubik_last_error", "jubik_plan_launch_count", "jubik_sharded_flat_size"):
            getattr(lib, name).restype = i
    return lib


def check(lib, rc):
    if rc != 0:
        raise JubikError("jubik_b200 error %d: %s" % (rc, lib.jubik_last_error().decode()))


_LIB = None


def load(require_device=True):
    """Load libjubik_b200.so (built in-tree by `python -m jubik_b200.build`).  Raises if the library
    is missing or (require_device) no CUDA device is visible."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise JubikError("libjubik_b200.so is not built (run `python -m jubik_b200.build`); "
                             "jubik_b200 has no CPU fallback")
        _LIB = bind(C.CDLL(LIB_PATH))
    if require_device and _LIB.jubik_device_count() < 1:
        raise JubikError("no CUDA device visible: jubik_b200 runs on sm_100a only and has no CPU fallback")
    return _LIB
