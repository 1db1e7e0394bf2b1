"""gpsa -- B200-native drop-in for the `gpsa` package of andrewcharlesjones/spatial-alignment.

Same public names as the reference's gpsa/__init__.py:1-16; the variational ELBO hot path
(VariationalGPSA.forward / loss_fn / backward and the rbf / matern12 kernel functions) runs in
hand-written sm_100a CUDA kernels behind the C ABI in include/gpsa_b200.h.  CUDA-only: there is
no CPU fallback.
"""
from gpsa.models.gpsa import GPSA
from gpsa.models.vgpsa import VariationalGPSA
from gpsa.util.util import (
    rbf_kernel,
    matern12_kernel,
    matern32_kernel,
    polar_warp,
    get_st_coordinates,
    LossNotDecreasingChecker,
)
from gpsa.util.util import rbf_kernel_numpy
from gpsa.plotting import (
    callback_oned,
    callback_twod,
    callback_twod_aligned_only,
    callback_twod_multimodal,
)

__version__ = "0.6+b200"
