"""Golden vectors for the device PSF builder: executes the REFERENCE'S OWN
/root/reference/jubik/instruments/erosita/erosita_psf_utils.py (psf_interpolator, get_psf, :17-264) on synthetic PSF
tables under the NumPy stand-in for jax of make_golden.py, extended by the four extra entry points that file uses:

    jax.lax.cond                        -> Python branch
    jax.scipy.ndimage.map_coordinates   -> scipy.ndimage.map_coordinates (order=1, mode='nearest': same definition)
    x.at[i].set(v)                      -> copy with the item replaced (ndarray subclass returned by every jnp function)
    jax.vmap                            -> the loop of make_golden.py

Run in THIS container only (the GPU box has no /root/reference):  python tests/golden/make_golden_psf.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from tests import synth  # noqa: E402

REF_FILE = "/root/reference/jubik/instruments/erosita/erosita_psf_utils.py"


class JArr(np.ndarray):
    class _At:
        def __init__(self, arr):
            self.arr = arr

        def __getitem__(self, idx):
            arr = self.arr

            class _Setter:
                def set(self, v):
                    out = np.array(arr, copy=True).view(JArr)
                    out[idx] = v
                    return out
            return _Setter()

    @property
    def at(self):
        return JArr._At(self)


def _wrap(f):
    def g(*a, **k):
        r = f(*a, **k)
        return r.view(JArr) if isinstance(r, np.ndarray) else r
    return g


def install():
    Domain = mg._install_stand_ins()
    jnp = sys.modules["jax.numpy"]
    for name in ("zeros", "array", "stack", "moveaxis", "sqrt", "angle", "cos", "sin", "flip"):
        setattr(jnp, name, _wrap(getattr(np, name)))
    jnp.pi = np.pi
    lax = sys.modules["jax.lax"]
    lax.cond = lambda pred, t, f, operand: t(operand) if bool(pred) else f(operand)
    from scipy import ndimage
    jsp = types.ModuleType("jax.scipy")
    jnd = types.ModuleType("jax.scipy.ndimage")
    jnd.map_coordinates = lambda inp, coords, order, mode="constant": ndimage.map_coordinates(
        np.asarray(inp, dtype=float), np.asarray(coords), order=order, mode=mode)
    jsp.ndimage = jnd
    sys.modules["jax.scipy"] = jsp
    sys.modules["jax.scipy.ndimage"] = jnd
    sys.modules["jax"].scipy = jsp
    return Domain


def main():
    Domain = install()
    spec = importlib.util.spec_from_file_location("jubik_erosita_psf_utils", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name, (E, N, n, nrad, px, dist) in {"psf_interp_N48_n4": (2, 48, 4, 4, 33, 8.0),
                                             "psf_interp_N40_n5": (1, 40, 5, 5, 25, 10.0)}.items():
        infos = synth.psf_infos(E, nrad, px, seed=sum(map(ord, name)))
        dom = Domain((E, N, N), (1.0, dist, dist))
        out = np.asarray(mod.psf_interpolator(dom, n, infos))
        assert out.shape == (n * n, E, N, N)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), E=E, N=N, n=n, nrad=nrad, px=px, dist=dist,
                            seed=sum(map(ord, name)), kernel=out)
        print(name, out.shape, float(out.max()), float(out.sum()))


if __name__ == "__main__":
    main()
