#!/usr/bin/env python
"""Where an MMGKS iteration goes at configs[2] (deblurring 1024^2): wall clock of the stages of the loop with a device
synchronisation after each (so device and host time add up), averaged over the iterations of one warm run."""
import os
import sys
import time
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import trips_b200 as tb  # noqa: E402
import trips_oracle as O  # noqa: E402
from trips_b200 import kernels as K  # noqa: E402
M = sys.modules[tb.MMGKS.__module__]  # the solver's module (the package re-exports the function under the same name)
C = sys.modules[M.GKSBases.__module__]

acc = defaultdict(float)


def timed(name, fn, sync=True):
    def run(*a, **kw):
        t0 = time.perf_counter()
        r = fn(*a, **kw)
        if sync:
            torch.cuda.synchronize()
        acc[name] += time.perf_counter() - t0
        return r
    return run


def main():
    n = int(os.environ.get("N", 1024))
    rng = np.random.default_rng(2022)
    PSF = tb.gauss_psf((9, 9), (3, 3))
    Ab = tb.PSFBlur2D(PSF, n, n)
    xt = O.shepp_logan(n).reshape(-1, 1)
    b, delta = O.add_noise(O.blur_data(xt, PSF, n, n), 0.01, rng)
    L = tb.FirstDerivative2D(n, n)
    call = lambda: tb.MMGKS(Ab, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=50, regparam="dp", delta=float(delta), x_true=xt)
    call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    call()
    torch.cuda.synchronize()
    base = time.perf_counter() - t0
    # instrument
    K.weighted_gram = timed("gram: device pass + D2H", K.weighted_gram)
    K.gram_factor = timed("gram: host dd Cholesky", K.gram_factor, sync=False)
    M.choose_lambda = timed("choose_lambda (dp, host)", M.choose_lambda, sync=False)
    M.tikhonov_projected = timed("tikhonov_projected (host)", M.tikhonov_projected, sync=False)
    M.expand = timed("expand: CGS2 + A v, L v", M.expand)
    M.apply_L_with_weights = timed("L x + weights", M.apply_L_with_weights)
    M.adjoint_L_weighted = timed("L^T (w r)", M.adjoint_L_weighted)
    orig_combine = K.basis_combine
    K.basis_combine = timed("basis_combine (V y, AV y, LV y)", orig_combine)
    C.K.basis_combine = K.basis_combine
    t0 = time.perf_counter()
    call()
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print(f"warm run {base / 50 * 1e3:.2f} ms/it; instrumented (serialised) {tot / 50 * 1e3:.2f} ms/it")
    for k_, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print(f"  {k_:40s} {v / 50 * 1e3:6.3f} ms/it")
    print(f"  {'(everything else)':40s} {(tot - sum(acc.values())) / 50 * 1e3:6.3f} ms/it")


if __name__ == "__main__":
    main()
