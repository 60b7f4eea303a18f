#!/usr/bin/env python
"""cProfile of the host side of MMGKS at configs[2] (deblurring 1024^2) and configs[4] (dynamic CT 256^2 x 64): where the
per-iteration wall clock goes once the kernels are short (tottime table, host functions only)."""
import cProfile
import io
import os
import pstats
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import trips_b200 as tb  # noqa: E402
import trips_oracle as O  # noqa: E402


def profile(name, fn):
    fn()  # warm-up
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    fn()
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
    print(f"==== {name}\n{s.getvalue()}", flush=True)


def main():
    rng = np.random.default_rng(2022)
    n = 1024
    PSF = tb.gauss_psf((9, 9), (3, 3))
    Ab = tb.PSFBlur2D(PSF, n, n)
    xt = O.shepp_logan(n).reshape(-1, 1)
    b, delta = O.add_noise(O.blur_data(xt, PSF, n, n), 0.01, rng)
    L = tb.FirstDerivative2D(n, n)
    profile("cfg3 MMGKS anisoTV 1024^2 50 it dp",
            lambda: tb.MMGKS(Ab, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=50, regparam="dp", delta=float(delta), x_true=xt))
    profile("cfg3 GKS 1024^2 50 it gcv",
            lambda: tb.GKS(Ab, b, L, projection_dim=3, n_iter=50, regparam="gcv", x_true=xt))
    profile("cfg3 Hybrid_GMRES 1024^2 50 it dp",
            lambda: tb.Hybrid_GMRES(Ab, b, 50, regparam="dp", delta=float(delta), x_true=xt))


if __name__ == "__main__":
    main()
