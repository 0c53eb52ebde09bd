"""rodeo_legacy_b200 -- B200-native batched backend for rodeo's probabilistic ODE solver.

Only the one data-parallel hot path of mlysy/rodeo-legacy is here: batches of independent
``KalmanODE`` solves (forward time-varying Kalman filter + backward smoother) as
hand-written sm_100a CUDA kernels behind a C-ABI shared library, with a Python front end
(``rodeo_legacy_b200.cuda.KalmanODE``) that keeps the reference's constructor and
``solve_mv`` / ``solve_sim`` / ``solve`` signatures, plus the host-side prior helpers
(``ibm_init``, ``car_init``, ``indep_init``, ``zero_pad``, ``rand_mat``).

Importing this package does not load the CUDA library; ``rodeo_legacy_b200.cuda`` does,
and fails loudly when it is missing -- there is no CPU fallback.
"""
from .ibm import ibm_init, ibm_state
from .car import car_init
from .utils import indep_init, zero_pad, rand_mat, mvncond, solveV, norm_sim

__all__ = ["ibm_init", "ibm_state", "car_init", "indep_init", "zero_pad", "rand_mat",
           "mvncond", "solveV", "norm_sim"]
__version__ = "0.1.0"
