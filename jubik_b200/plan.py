"""ResponsePlan -- device-resident state of one instrument response (PSF spectra, effective exposure,
mask, counts) and the calls that run the CUDA hot path on torch CUDA tensors.

torch is used for device memory and streams only; every flop of the path runs in libjubik_b200.so.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import PlanDesc, PlanInfo, check


class RawPlan:
    """Pointer-level wrapper around jubik_plan* (no array library involved)."""

    def __init__(self, lib, desc: PlanDesc, stream=0):
        self.lib = lib
        self.handle = C.c_void_p()
        check(lib, lib.jubik_plan_create(C.byref(self.handle), C.byref(desc), C.c_void_p(stream)))
        info = PlanInfo()
        check(lib, lib.jubik_plan_get_info(self.handle, C.byref(info)))
        self.info = {k: getattr(info, k) for k, _ in PlanInfo._fields_}

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.jubik_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name, *args):
        if not self.handle:
            raise _lib.JubikError("plan is closed")
        check(self.lib, getattr(self.lib, name)(self.handle, *args))

    @property
    def launch_count(self):
        return int(self.lib.jubik_plan_launch_count(self.handle))


def make_desc(N, E, T, n_patches, margin, kernel_ptr, cropped, max_samples=1, exposure_ptr=None,
              exposure_cut=None, eff_area_ptr=None, pixel_area=1.0, data_ptr=None, data_is_int=False,
              distances=(1.0, 1.0), fft_length=0, N1=None, flags=0):
    d = PlanDesc()
    d.n0, d.n1 = int(N), int(N if N1 is None else N1)
    d.n_energy, d.n_modules, d.max_samples = int(E), int(T), int(max_samples)
    d.n_patches, d.margin = int(n_patches), int(margin)
    d.kernel_layout = _lib.KERNEL_CROPPED if cropped else _lib.KERNEL_FULL
    d.data_dtype = _lib.DATA_I64 if data_is_int else _lib.DATA_F64
    d.has_exposure_cut = 0 if exposure_cut is None else 1
    d.exposure_cut = 0.0 if exposure_cut is None else float(exposure_cut)
    d.fft_length = int(fft_length)
    d.flags = int(flags)
    d.d0, d.d1 = float(distances[0]), float(distances[1])
    d.pixel_area = float(pixel_area)
    d.kernel = kernel_ptr
    d.exposure = exposure_ptr
    d.eff_area = eff_area_ptr
    d.data = data_ptr
    return d


class ResponsePlan:
    """R(x) = mask(exposure(psf(x * pixel_area))) and the Poisson likelihood on it, on one GPU.

    kernel   : torch.float64 CUDA tensor (T, n*n, E, N, N) origin-centred  [reference layout,
               erosita_psf_utils.py:262-263] or, with cropped=True, (T, n*n, E, 2dx, 2dx) centred.
    exposure : (T, E, N, N) raw exposures or None (PSF-only plan); thresholded with exposure_cut
               exactly like response.py:37-40 / :85-87 (mask == thresholded exposure == 0).
    data     : (T, E, N, N) dense counts (int64 or float64) or None.
    host_call / metric / sharded : which optional staging the plan allocates at creation (JUBIK_PLAN_* flags of
               include/jubik_b200.h): the host-buffer entry point, the metric's cached linearisation point, and the
               sharded entry point's per-sample gradient staging + collective stream.  Nothing is allocated later.
    """

    def __init__(self, kernel, N, n_patches, margin, exposure=None, exposure_cut=None, eff_area=None,
                 pixel_area=1.0, data=None, max_samples=1, cropped=False, distances=(1.0, 1.0),
                 fft_length=0, host_call=False, metric=True, sharded=False):
        import torch
        lib = _lib.load()
        if not (kernel.is_cuda and kernel.dtype == torch.float64):
            raise TypeError("kernel must be a float64 CUDA tensor (jubik_b200 has no CPU path)")
        if kernel.dim() != 5:
            raise ValueError("kernel must have shape (T, n*n, E, H, W)")
        kernel = kernel.contiguous()
        T, nn2, E = kernel.shape[0], kernel.shape[1], kernel.shape[2]
        if nn2 != n_patches * n_patches:
            raise ValueError("kernel.shape[1] must be n_patches_per_axis**2")
        dx = int(N / n_patches)
        exp_hw = (2 * dx, 2 * dx) if cropped else (N, N)
        if tuple(kernel.shape[-2:]) != exp_hw:
            raise ValueError("kernel spatial shape %s does not match layout %s" % (tuple(kernel.shape[-2:]), exp_hw))
        self.device = kernel.device
        self.T, self.E, self.N, self.S_max = T, E, N, max_samples
        keep = []
        eptr = fptr = dptr = None
        if exposure is not None:
            exposure = exposure.to(self.device, torch.float64).contiguous()
            if tuple(exposure.shape) != (T, E, N, N):
                raise ValueError("exposure should have shape (T, E, N, N)")
            eptr = exposure.data_ptr()
            keep.append(exposure)
        if eff_area is not None:
            eff_area = eff_area.to(self.device, torch.float64).contiguous()
            if tuple(eff_area.shape) != (T, E):
                raise ValueError("eff_area should have shape (T, E)")
            fptr = eff_area.data_ptr()
            keep.append(eff_area)
        data_is_int = False
        if data is not None:
            data, data_is_int = self._prep_data(data)
            dptr = data.data_ptr()
            keep.append(data)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            flags = ((_lib.PLAN_HOST_CALL if host_call else 0) | (_lib.PLAN_METRIC if (metric and exposure is not None) else 0) |
                     (_lib.PLAN_SHARDED if sharded else 0))
            desc = make_desc(N, E, T, n_patches, margin, kernel.data_ptr(), cropped, max_samples, eptr,
                             exposure_cut, fptr, pixel_area, dptr, data_is_int, distances, fft_length, flags=flags)
            self.raw = RawPlan(lib, desc, stream)
        self.info = self.raw.info
        self.cropped = bool(cropped)
        self.n_patches, self.margin = int(n_patches), int(margin)
        self._remember_kernel(kernel)                        # identity short-cut of ensure_kernel
        del keep

    def _remember_kernel(self, kernel):
        """Weak reference (the plan must not keep tens of GB of PSF kernels alive) + version counter of the tensor the
        resident spectra were built from."""
        import weakref
        self._k_obj, self._k_ver = weakref.ref(kernel), kernel._version

    def _prep_data(self, data):
        import torch
        data = data.to(self.device)
        if data.dtype in (torch.int64,):
            is_int = True
        elif data.dtype in (torch.int32, torch.int16, torch.uint8, torch.int8):
            data, is_int = data.to(torch.int64), True
        else:
            data, is_int = data.to(torch.float64), False
        data = data.contiguous()
        if tuple(data.shape) != (self.T, self.E, self.N, self.N):
            raise ValueError("data should have shape (T, E, N, N)")
        return data, is_int

    # -------------------------------------------------------------------------------- helpers
    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, name, *args):
        """One C-ABI call with the plan's device current (a process may hold plans on several GPUs)."""
        import torch
        with torch.cuda.device(self.device):
            self.raw.call(name, *args, self._stream())

    def _on_device(self, t, what):
        import torch
        if not (t.is_cuda and t.dtype == torch.float64):
            raise TypeError("%s must be a float64 CUDA tensor" % what)
        if t.device != self.device:
            raise ValueError("%s lives on %s but the plan on %s" % (what, t.device, self.device))

    def _sky(self, x, chunked=False):
        self._on_device(x, "input")
        squeeze = x.dim() == 3
        if squeeze:
            x = x.unsqueeze(0)
        if x.dim() != 4 or tuple(x.shape[1:]) != (self.E, self.N, self.N):
            raise ValueError("sky should have shape ([S,] E, N, N)")
        if x.shape[0] > self.S_max and not chunked:
            raise ValueError("more samples than the plan's max_samples")
        return x.contiguous(), squeeze

    def _det(self, g):
        self._on_device(g, "input")
        squeeze = g.dim() == 4
        if squeeze:
            g = g.unsqueeze(0)
        if g.dim() != 5 or tuple(g.shape[1:]) != (self.T, self.E, self.N, self.N):
            raise ValueError("detector-space input should have shape ([S,] T, E, N, N)")
        if g.shape[0] > self.S_max:
            raise ValueError("more samples than the plan's max_samples")
        return g.contiguous(), squeeze

    def _fwd(self, name, x):
        import torch
        x, sq = self._sky(x)
        out = torch.empty((x.shape[0], self.T, self.E, self.N, self.N), dtype=torch.float64, device=self.device)
        self._call(name, C.c_void_p(x.data_ptr()), x.shape[0], C.c_void_p(out.data_ptr()))
        return out[0] if sq else out

    def _adj(self, name, g):
        import torch
        g, sq = self._det(g)
        out = torch.empty((g.shape[0], self.E, self.N, self.N), dtype=torch.float64, device=self.device)
        self._call(name, C.c_void_p(g.data_ptr()), g.shape[0], C.c_void_p(out.data_ptr()))
        return out[0] if sq else out

    # ---------------------------------------------------------------------------------- calls
    def psf(self, x):
        """vmap(linpatch_convolve) over modules: ([S,]E,N,N) -> ([S,]T,E,N,N)."""
        return self._fwd("jubik_psf_apply", x)

    def psf_T(self, g):
        return self._adj("jubik_psf_apply_T", g)

    def response(self, x):
        """Dense R(x): exposure * psf(x * pixel_area), masked pixels = 0."""
        return self._fwd("jubik_response_apply", x)

    def response_T(self, g):
        return self._adj("jubik_response_apply_T", g)

    def set_data(self, data):
        data, is_int = self._prep_data(data)
        self._call("jubik_plan_set_data", C.c_void_p(data.data_ptr()), _lib.DATA_I64 if is_int else _lib.DATA_F64)
        import torch
        torch.cuda.current_stream(self.device).synchronize()

    def set_kernel(self, kernel, cropped=None):
        """Replace the PSF kernels (same geometry) -- the dynamic-kernel call form R(x, k) of the eROSITA response
        (erosita_response.py:349-350).  Redoes the PSF preparation and its FFT into the resident spectra."""
        import torch
        cropped = self.cropped if cropped is None else bool(cropped)
        self._on_device(kernel, "kernel")
        kernel = kernel.contiguous()
        dx = int(self.N / self.n_patches)
        exp = (self.T, self.n_patches ** 2, self.E) + ((2 * dx, 2 * dx) if cropped else (self.N, self.N))
        if tuple(kernel.shape) != exp:
            raise ValueError("kernel should have shape %s" % (exp,))
        self._k_obj = None
        self._call("jubik_plan_set_kernel", C.c_void_p(kernel.data_ptr()), _lib.KERNEL_CROPPED if cropped else _lib.KERNEL_FULL)
        self._remember_kernel(kernel)
        info = PlanInfo()
        check(self.raw.lib, self.raw.lib.jubik_plan_get_info(self.raw.handle, C.byref(info)))
        self.info = self.raw.info = {k: getattr(info, k) for k, _ in PlanInfo._fields_}

    def ensure_kernel(self, kernel, cropped=None):
        """set_kernel only if `kernel` differs in CONTENT from the kernels the resident spectra were built from
        (jubik_plan_ensure_kernel: device-side 128-bit fingerprint).  The tensor object last seen (unmodified since) is
        recognised without touching the device.  Returns True if the spectra were rebuilt."""
        if self._k_obj is not None and self._k_obj() is kernel and kernel._version == self._k_ver:
            return False
        cropped = (tuple(kernel.shape[-2:]) != (self.N, self.N)) if cropped is None else bool(cropped)
        self._on_device(kernel, "kernel")
        src, kernel = kernel, kernel.contiguous()
        dx = int(self.N / self.n_patches)
        exp = (self.T, self.n_patches ** 2, self.E) + ((2 * dx, 2 * dx) if cropped else (self.N, self.N))
        if tuple(kernel.shape) != exp:
            raise ValueError("kernel should have shape %s" % (exp,))
        flag = C.c_int(0)
        self._k_obj = None
        self._call("jubik_plan_ensure_kernel", C.c_void_p(kernel.data_ptr()),
                   _lib.KERNEL_CROPPED if cropped else _lib.KERNEL_FULL, C.byref(flag))
        self._remember_kernel(src)
        if flag.value:
            info = PlanInfo()
            check(self.raw.lib, self.raw.lib.jubik_plan_get_info(self.raw.handle, C.byref(info)))
            self.info = self.raw.info = {k: getattr(info, k) for k, _ in PlanInfo._fields_}
        return bool(flag.value)

    def energy_and_grad(self, x, need_grad=True, out_energy=None, out_grad=None):
        """Poisson energy sum(lambda) - d.log(lambda) per sample and its gradient w.r.t. the sky.
        out_energy (S,) / out_grad (S,E,N,N): optional preallocated contiguous float64 CUDA outputs (e.g. views
        into one flat buffer that is then all-reduced in a single NCCL call)."""
        import torch
        x, sq = self._sky(x)
        S = x.shape[0]
        energy = out_energy if out_energy is not None else torch.empty((S,), dtype=torch.float64, device=self.device)
        grad = (out_grad if out_grad is not None else torch.empty_like(x)) if need_grad else None
        if energy.numel() != S or not energy.is_contiguous() or (need_grad and (grad.numel() != x.numel() or not grad.is_contiguous())):
            raise ValueError("output buffers must be contiguous and match (S,) / (S,E,N,N)")
        self._on_device(energy, "out_energy")
        if need_grad:
            self._on_device(grad, "out_grad")
        self._call("jubik_poisson_energy_and_grad", C.c_void_p(x.data_ptr()), S, C.c_void_p(energy.data_ptr()),
                   C.c_void_p(grad.data_ptr()) if need_grad else None)
        if sq:
            return energy[0], (grad[0] if need_grad else None)
        return energy, grad

    def energy(self, x):
        return self.energy_and_grad(x, need_grad=False)[0]

    def flat_size(self, comm, n_samples, shard_mode=_lib.SHARD_SAMPLES):
        """Element count of the flat all-reduce buffer of energy_and_grad_sharded (jubik_sharded_flat_size)."""
        n = int(self.raw.lib.jubik_sharded_flat_size(self.raw.handle, comm.handle if comm is not None else None,
                                                     int(shard_mode), int(n_samples)))
        if n < 0:
            raise ValueError("bad shard mode / sample count")
        return n

    def energy_and_grad_sharded(self, x, comm, shard_mode=_lib.SHARD_SAMPLES, flat=None, grad_local=None):
        """Multi-GPU fused evaluation: this rank's share of the work, then ONE ncclAllReduce of `flat` on the current
        stream (jubik_poisson_energy_and_grad_sharded).  comm: Communicator or None (single rank, no collective).
          SHARD_SAMPLES: x = this rank's (S,E,N,N) samples -> (energies (n_ranks,S), sample-summed gradient (E,N,N));
                         grad_local (S,E,N,N), if given, receives the per-sample gradients of this rank.
          SHARD_MODULES: the plan holds this rank's modules, x is replicated -> (energies (S,), gradient (S,E,N,N)).
        S may exceed max_samples: the library then evaluates the samples in chunks of max_samples.
        The returned tensors are views into `flat`."""
        import torch
        x, _ = self._sky(x, chunked=True)
        S = x.shape[0]
        n = self.flat_size(comm, S, shard_mode)
        if flat is None:
            flat = torch.empty((n,), dtype=torch.float64, device=self.device)
        if flat.numel() != n or not flat.is_contiguous() or flat.dtype != torch.float64 or not flat.is_cuda:
            raise ValueError("flat must be a contiguous float64 CUDA tensor of %d elements" % n)
        if grad_local is not None and (grad_local.numel() != x.numel() or not grad_local.is_contiguous()):
            raise ValueError("grad_local must be contiguous (S,E,N,N)")
        self._on_device(flat, "flat")
        self._call("jubik_poisson_energy_and_grad_sharded", comm.handle if comm is not None else None,
                   int(shard_mode), C.c_void_p(x.data_ptr()), S,
                   C.c_void_p(grad_local.data_ptr()) if grad_local is not None else None,
                   C.c_void_p(flat.data_ptr()))
        npx = self.E * self.N * self.N
        if shard_mode == _lib.SHARD_SAMPLES:
            R = comm.size if comm is not None else 1
            return flat[npx:].view(R, S), flat[:npx].view(self.E, self.N, self.N)
        return flat[S * npx:], flat[:S * npx].view(S, self.E, self.N, self.N)

    def energy_and_grad_sharded_host(self, x_host, comm, flat_host):
        """End-to-end form of the sample-sharded call on HOST tensors (jubik_poisson_energy_and_grad_sharded_host):
        x_host (S,E,N,N) float64 CPU (pinned for full PCIe speed; S may exceed max_samples), flat_host float64 CPU of
        E*N*N + n_ranks*S elements <- [sample-summed gradient over all ranks | energies (n_ranks, S)].  Synchronises."""
        import torch
        R = comm.size if comm is not None else 1
        if x_host.dim() != 4 or tuple(x_host.shape[1:]) != (self.E, self.N, self.N):
            raise ValueError("x_host should have shape (S, E, N, N)")
        S = x_host.shape[0]
        npx = self.E * self.N * self.N
        for name, t, numel in (("x_host", x_host, S * npx), ("flat_host", flat_host, npx + R * S)):
            if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous():
                raise TypeError("%s must be a contiguous float64 CPU tensor" % name)
            if t.numel() != numel:
                raise ValueError("%s has %d elements, expected %d" % (name, t.numel(), numel))
        self._call("jubik_poisson_energy_and_grad_sharded_host", comm.handle if comm is not None else None,
                   C.c_void_p(x_host.data_ptr()), S, C.c_void_p(flat_host.data_ptr()))
        return flat_host[npx:].view(R, S), flat_host[:npx].view(self.E, self.N, self.N)

    def metric(self, x, v):
        """Fisher metric pulled back through R:  R^T diag(1/(R x)) R v."""
        import torch
        x, sq = self._sky(x)
        v, _ = self._sky(v)
        if v.shape != x.shape:
            raise ValueError("x and v must have the same shape")
        out = torch.empty_like(x)
        self._call("jubik_poisson_metric_apply", C.c_void_p(x.data_ptr()), C.c_void_p(v.data_ptr()), x.shape[0],
                   C.c_void_p(out.data_ptr()))
        return out[0] if sq else out

    def left_sqrt_metric(self, x, t):
        """R^T (t / sqrt(R x)): the Poissonian left-sqrt-metric pulled back to the sky (t in data space (S,T,E,N,N))."""
        import torch
        x, sq = self._sky(x)
        t, _ = self._det(t)
        if t.shape[0] != x.shape[0]:
            raise ValueError("x and t must have the same number of samples")
        out = torch.empty_like(x)
        self._call("jubik_poisson_left_sqrt_metric", C.c_void_p(x.data_ptr()), C.c_void_p(t.data_ptr()), x.shape[0],
                   C.c_void_p(out.data_ptr()))
        return out[0] if sq else out

    def set_point(self, x):
        """Cache the linearisation point of the metric: one forward pass stores expo^2 / (R x).  NIFTy's CG applies the
        metric 60-70 times per sample at a fixed position (eROSITA_demo.yaml:161)."""
        x, _ = self._sky(x)
        self._call("jubik_poisson_set_point", C.c_void_p(x.data_ptr()), x.shape[0])
        self._point_S = x.shape[0]

    def metric_at_point(self, v):
        """R^T diag(1/lambda(x0)) R v at the point of the last set_point: one forward + one adjoint pass."""
        import torch
        v, sq = self._sky(v)
        out = torch.empty_like(v)
        self._call("jubik_poisson_metric_apply_at_point", C.c_void_p(v.data_ptr()), v.shape[0], C.c_void_p(out.data_ptr()))
        return out[0] if sq else out

    def left_sqrt_metric_at_point(self, t):
        """R^T (t / sqrt(lambda(x0))) at the point of the last set_point: one adjoint pass."""
        import torch
        t, sq = self._det(t)
        out = torch.empty((t.shape[0], self.E, self.N, self.N), dtype=torch.float64, device=self.device)
        self._call("jubik_poisson_left_sqrt_metric_at_point", C.c_void_p(t.data_ptr()), t.shape[0], C.c_void_p(out.data_ptr()))
        return out[0] if sq else out

    def energy_and_grad_logsky(self, log_sky):
        """Fused value + gradient with respect to log_sky, sky = exp(log_sky[:E, :N, :N]) (sky_models.py:340-341).
        log_sky: ([S,] E, N, N), or the sky model's PADDED array ([S,] e_pad, n_pad, n_pad): the crop is read in place
        (strided), the gradient comes back in the padded shape with zeros outside the crop."""
        import torch
        self._on_device(log_sky, "input")
        sq = log_sky.dim() == 3
        s = (log_sky.unsqueeze(0) if sq else log_sky).contiguous()
        if s.dim() != 4 or s.shape[0] > self.S_max:
            raise ValueError("log_sky should have shape ([S,] E, N, N) with S <= max_samples")
        energy = torch.empty((s.shape[0],), dtype=torch.float64, device=self.device)
        grad = torch.empty_like(s)
        if tuple(s.shape[1:]) == (self.E, self.N, self.N):
            self._call("jubik_poisson_energy_and_grad_logsky", C.c_void_p(s.data_ptr()), s.shape[0],
                       C.c_void_p(energy.data_ptr()), C.c_void_p(grad.data_ptr()))
        else:
            ep, np0, np1 = s.shape[1:]
            if np0 != np1 or ep < self.E or np0 < self.N:
                raise ValueError("a padded log_sky should have shape ([S,] e_pad >= E, n_pad >= N, n_pad)")
            self._call("jubik_poisson_energy_and_grad_logsky_padded", C.c_void_p(s.data_ptr()), s.shape[0], int(ep), int(np0),
                       C.c_void_p(energy.data_ptr()), C.c_void_p(grad.data_ptr()))
        return (energy[0], grad[0]) if sq else (energy, grad)

    def energy_and_grad_host(self, x_host, energy_host, grad_host):
        """End-to-end call on HOST tensors (pinned for full PCIe speed): H2D, kernels, D2H, sync.
        x_host ([S,]E,N,N), energy_host (S,), grad_host like x_host or None: float64 CPU, contiguous."""
        import torch
        S = x_host.shape[0] if x_host.dim() == 4 else 1
        if S > self.S_max:
            raise ValueError("more samples than the plan's max_samples")
        want = S * self.E * self.N * self.N
        for name, t, numel in (("x_host", x_host, want), ("energy_host", energy_host, S), ("grad_host", grad_host, want)):
            if t is None and name == "grad_host":
                continue
            if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous():
                raise TypeError("%s must be a contiguous float64 CPU tensor" % name)
            if t.numel() != numel:
                raise ValueError("%s has %d elements, expected %d" % (name, t.numel(), numel))
        self._call("jubik_poisson_energy_and_grad_host", C.c_void_p(x_host.data_ptr()), S,
                   C.c_void_p(energy_host.data_ptr()),
                   C.c_void_p(grad_host.data_ptr()) if grad_host is not None else None)

    KERNEL_NAMES = {7: ("rows_fwd", "cols_fwd", "rows_out", "reduce", "rows_adj_in", "cols_adj", "rows_adj_out")}

    def profile(self, x):
        """energy_and_grad with a CUDA event after each kernel -> (energy, grad, {kernel: ms})."""
        import torch
        x, sq = self._sky(x)
        S = x.shape[0]
        energy = torch.empty((S,), dtype=torch.float64, device=self.device)
        grad = torch.empty_like(x)
        ms = (C.c_float * 48)()
        n = C.c_int(0)
        with torch.cuda.device(self.device):
            self.raw.call("jubik_poisson_profile", C.c_void_p(x.data_ptr()), S, C.c_void_p(energy.data_ptr()),
                          C.c_void_p(grad.data_ptr()), ms, C.byref(n), self._stream())
        vals = [float(ms[i]) for i in range(n.value)]
        names = self.KERNEL_NAMES.get(n.value, tuple("k%d" % i for i in range(n.value)))
        return energy, grad, dict(zip(names, vals))

    @property
    def launch_count(self):
        return self.raw.launch_count

    def close(self):
        self.raw.close()


def measure_fp64_tflops():
    """DFMA micro-benchmark on the current device (second roofline denominator)."""
    import torch
    lib = _lib.load()
    out = C.c_double()
    check(lib, lib.jubik_measure_fp64_tflops(C.byref(out), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.value


class Communicator:
    """jubik_comm*: an NCCL communicator owned by libjubik_b200.so (one rank per process / GPU).

    `Communicator.from_torch_distributed()` creates the ncclUniqueId on rank 0, distributes it through the already
    initialised torch.distributed group (any backend -- plumbing only) and calls ncclCommInitRank on the current
    CUDA device.  The all-reduce itself is issued by the C library, on the stream of the fused evaluation."""

    def __init__(self, unique_id: bytes, rank: int, world: int):
        self.lib = _lib.load()
        self.handle = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        check(self.lib, self.lib.jubik_comm_create(C.byref(self.handle), buf, int(rank), int(world)))
        self.rank, self.size = int(rank), int(world)

    @staticmethod
    def unique_id() -> bytes:
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        check(lib, lib.jubik_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, group=None):
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(box[0], rank, world)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.jubik_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
