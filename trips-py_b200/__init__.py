"""trips_b200 - the Krylov hot path of TRIPs-Py (mpasha3/trips-py) on NVIDIA B200 (sm_100a).

Import name: `trips_b200` (the directory is `trips-py_b200/`; `trips_b200.py` at the repository root maps one to
the other).  Everything numerical runs in hand-written CUDA kernels behind the C ABI of include/tripsb200.h;
there is no CPU fallback.
"""
from . import kernels, operators, decompositions, reg_param, solvers  # noqa: F401
from .decompositions import (ArnoldiState, GKState, arnoldi_update, golub_kahan, golub_kahan_device,  # noqa: F401
                             golub_kahan_update, release_host_buffers)
from .operators import (BlockDiagCT, CenteredDerivative2D, CSROperator, FanBeamCT, FirstDerivative1D, FrameletOperator,  # noqa: F401
                        FirstDerivative2D, Identity, LinearOperator, ParallelBeamCT, PSFBlur2D, SpaceTimeDerivative,
                        as_operator, gauss_psf)
from .solvers import CGLS, GKS, MMGKS, Hybrid_GMRES, Hybrid_LSQR  # noqa: F401
from . import test_problems  # noqa: F401
from .test_problems import Deblurring2D, Tomography  # noqa: F401

__version__ = "0.1.0"
