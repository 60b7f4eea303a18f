#!/usr/bin/env python
"""One MMGKS run each at configs[2] (deblurring 1024^2) and configs[4] (dynamic CT 256^2 x 64) for an ncu capture of the
kernels next to the projectors: basis_dots / basis_combine, gram_dd, correlate2d, fd_apply / fd_adjoint, vec_*."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import trips_b200 as tb  # noqa: E402
import trips_oracle as O  # noqa: E402

rng = np.random.default_rng(2022)
n = 1024
PSF = tb.gauss_psf((9, 9), (3, 3))
Ab = tb.PSFBlur2D(PSF, n, n)
xt = O.shepp_logan(n).reshape(-1, 1)
b, delta = O.add_noise(O.blur_data(xt, PSF, n, n), 0.01, rng)
tb.MMGKS(Ab, b, tb.FirstDerivative2D(n, n), pnorm=2, qnorm=1, projection_dim=3, n_iter=int(os.environ.get("N_ITER", 12)), regparam="dp",
         delta=float(delta))
if "--cfg5" in sys.argv:
    nx, nt, per = 256, 64, 12
    th = O.ct_angles(nt * per)
    Ad = tb.BlockDiagCT(nx, [th[t::nt] for t in range(nt)])
    base = O.shepp_logan(nx)
    xt = np.concatenate([(base * (1 + 0.3 * np.sin(2 * np.pi * t / nt))).ravel() for t in range(nt)]).reshape(-1, 1)
    b, delta = O.add_noise(Ad @ xt, 0.01, rng)
    tb.MMGKS(Ad, b, tb.SpaceTimeDerivative(nx, nx, nt), pnorm=2, qnorm=1, projection_dim=1, n_iter=12, regparam="dp",
             delta=float(delta), epsilon=0.1)
