"""NumPy front end of the host-thread EMULATION build (tests only).  Same C ABI, same sources as the
CUDA library, but "device" pointers are host arrays.  Never imported by jubik_b200."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from jubik_b200 import _lib  # noqa: E402
from jubik_b200.plan import RawPlan, make_desc  # noqa: E402
from tests.emu import build_emu  # noqa: E402

_EMU = None


def emu_lib():
    global _EMU
    if _EMU is None:
        _EMU = _lib.bind(C.CDLL(build_emu.build()))
    return _EMU


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class EmuPlan:
    def __init__(self, kernel, N, n, m, exposure=None, exposure_cut=None, eff_area=None, pixel_area=1.0,
                 data=None, max_samples=1, cropped=False, fft_length=0, flags=None):
        kernel = np.ascontiguousarray(kernel, dtype=np.float64)
        self.T, _, self.E = kernel.shape[:3]
        self.N, self.S_max = N, max_samples
        exposure = None if exposure is None else np.ascontiguousarray(exposure, dtype=np.float64)
        eff_area = None if eff_area is None else np.ascontiguousarray(eff_area, dtype=np.float64)
        is_int = False
        if data is not None:
            is_int = np.issubdtype(np.asarray(data).dtype, np.integer)
            data = np.ascontiguousarray(data, dtype=np.int64 if is_int else np.float64)
        desc = make_desc(N, self.E, self.T, n, m, kernel.ctypes.data, cropped, max_samples,
                         None if exposure is None else exposure.ctypes.data, exposure_cut,
                         None if eff_area is None else eff_area.ctypes.data, pixel_area,
                         None if data is None else data.ctypes.data, is_int, fft_length=fft_length,
                         flags=(_lib.PLAN_METRIC | _lib.PLAN_SHARDED | _lib.PLAN_HOST_CALL) if flags is None else flags)
        self.raw = RawPlan(emu_lib(), desc, 0)
        self.info = self.raw.info

    def _fwd(self, name, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        sq = x.ndim == 3
        x4 = x[None] if sq else x
        out = np.empty((x4.shape[0], self.T, self.E, self.N, self.N))
        self.raw.call(name, _p(x4), x4.shape[0], _p(out), None)
        return out[0] if sq else out

    def _adj(self, name, g):
        g = np.ascontiguousarray(g, dtype=np.float64)
        sq = g.ndim == 4
        g5 = g[None] if sq else g
        out = np.empty((g5.shape[0], self.E, self.N, self.N))
        self.raw.call(name, _p(g5), g5.shape[0], _p(out), None)
        return out[0] if sq else out

    def psf(self, x):
        return self._fwd("jubik_psf_apply", x)

    def psf_T(self, g):
        return self._adj("jubik_psf_apply_T", g)

    def response(self, x):
        return self._fwd("jubik_response_apply", x)

    def response_T(self, g):
        return self._adj("jubik_response_apply_T", g)

    def energy_and_grad(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        sq = x.ndim == 3
        x4 = x[None] if sq else x
        e = np.empty((x4.shape[0],))
        g = np.empty_like(x4)
        self.raw.call("jubik_poisson_energy_and_grad", _p(x4), x4.shape[0], _p(e), _p(g), None)
        return (e[0], g[0]) if sq else (e, g)

    def metric(self, x, v):
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        sq = x.ndim == 3
        x4, v4 = (x[None], v[None]) if sq else (x, v)
        out = np.empty_like(x4)
        self.raw.call("jubik_poisson_metric_apply", _p(x4), _p(v4), x4.shape[0], _p(out), None)
        return out[0] if sq else out

    def energy_and_grad_logsky(self, ls):
        ls = np.ascontiguousarray(ls, dtype=np.float64)
        sq = ls.ndim == 3
        x4 = ls[None] if sq else ls
        e = np.empty((x4.shape[0],))
        g = np.empty_like(x4)
        if x4.shape[1:] == (self.E, self.N, self.N):
            self.raw.call("jubik_poisson_energy_and_grad_logsky", _p(x4), x4.shape[0], _p(e), _p(g), None)
        else:       # the sky model's padded array: exp(log_sky[:E, :N, :N])
            self.raw.call("jubik_poisson_energy_and_grad_logsky_padded", _p(x4), x4.shape[0], x4.shape[1], x4.shape[2], _p(e), _p(g), None)
        return (e[0], g[0]) if sq else (e, g)

    def left_sqrt_metric(self, x, t):
        x = np.ascontiguousarray(x, dtype=np.float64)
        t = np.ascontiguousarray(t, dtype=np.float64)
        out = np.empty_like(x)
        self.raw.call("jubik_poisson_left_sqrt_metric", _p(x), _p(t), x.shape[0], _p(out), None)
        return out

    def set_point(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.raw.call("jubik_poisson_set_point", _p(x), x.shape[0], None)

    def metric_at_point(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty_like(v)
        self.raw.call("jubik_poisson_metric_apply_at_point", _p(v), v.shape[0], _p(out), None)
        return out

    def set_kernel(self, kernel, cropped):
        kernel = np.ascontiguousarray(kernel, dtype=np.float64)
        self.raw.call("jubik_plan_set_kernel", _p(kernel), _lib.KERNEL_CROPPED if cropped else _lib.KERNEL_FULL, None)

    def energy_and_grad_host(self, x):
        x4 = np.ascontiguousarray(x, dtype=np.float64)
        e = np.empty((x4.shape[0],))
        g = np.empty_like(x4)
        self.raw.call("jubik_poisson_energy_and_grad_host", _p(x4), x4.shape[0], _p(e), _p(g), None)
        return e, g

    def energy_and_grad_sharded(self, x, shard_mode, want_local=False):
        """Single-rank path of jubik_poisson_energy_and_grad_sharded (comm = NULL: no collective)."""
        x4 = np.ascontiguousarray(x, dtype=np.float64)
        S = x4.shape[0]
        n = int(self.raw.lib.jubik_sharded_flat_size(self.raw.handle, None, int(shard_mode), S))
        flat = np.full((n,), np.nan)
        gl = np.empty_like(x4) if want_local else None
        self.raw.call("jubik_poisson_energy_and_grad_sharded", None, int(shard_mode), _p(x4), S, _p(gl), _p(flat), None)
        return flat, gl
