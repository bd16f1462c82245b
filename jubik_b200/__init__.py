"""jubik_b200 -- B200-native (sm_100a CUDA) implementation of J-UBIK's instrument-response hot path:
spatially-variant PSF convolution (linpatch_convolve) fused with exposure, readout mask and the Poisson
log-likelihood, forward and exact adjoint, float64.  See DESIGN.md / INTEGRATION.md.

Importing the package does not need a GPU; every compute entry point does and raises otherwise.
"""
from ._lib import JubikError, load as load_library  # noqa: F401
from .api import (Domain, FullModel, PoissonLikelihood, Vector, build_exposure_function,  # noqa: F401
                  build_mask_adjoint, build_psf_operator, build_readout_function, build_response, calculate_nwr,
                  convolve, create_mock_data, integrate, linpatch_convolve, psf_interpolator, slice_patches,
                  slice_patches_T)
from .plan import Communicator, ResponsePlan  # noqa: F401
from ._lib import SHARD_MODULES, SHARD_SAMPLES  # noqa: F401
from . import sharding  # noqa: F401

__version__ = "0.1.0"
