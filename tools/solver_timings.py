#!/usr/bin/env python
"""Wall-clock of the full-size solver configurations of BASELINE.json that are not the bench headline:
configs[2] MMGKS l2-l1 (iso)TV deblurring 1024^2, configs[4] MMGKS space-time TV on dynamic CT 256^2 x 64 frames,
configs[1] Hybrid_LSQR + GCV on CT 256^2/180 views.  GPU only (the oracle takes minutes at these sizes); parity for the
same code paths is established at smaller sizes in tests/test_gpu_solvers.py."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import trips_b200 as tb  # noqa: E402
import trips_oracle as O  # noqa: E402
from trips_b200 import _lib  # noqa: E402


def run(name, fn, iters, warm=True):
    if warm:  # the first call of a configuration pays allocator growth and kernel attribute set-up
        fn()
    torch.cuda.synchronize()
    l0 = _lib.launch_count
    t0 = time.perf_counter()
    x, info = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    extra = f" RRE {info['relError'][-1]:.4f}" if "relError" in info else ""
    lam = info.get("regParam")
    lam = f"lambda {float(lam):.3e}" if np.isscalar(lam) else "lambda -"
    print(f"{name}: {dt:.2f} s total, {dt / iters * 1e3:.1f} ms/it, {(_lib.launch_count - l0) / iters:.0f} launches/it, "
          f"{lam}{extra}", flush=True)


def main():
    rng = np.random.default_rng(2022)
    only_cfg4 = "--cfg4" in sys.argv
    if not only_cfg4:
        small_and_mid(rng)
    cfg4(rng)


def small_and_mid(rng):
    # configs[1]
    A = tb.ParallelBeamCT(256, 180)
    xt = O.shepp_logan(256).reshape(-1, 1)
    bt = A @ xt
    b, delta = O.add_noise(bt, 0.01, rng)
    run("cfg2 Hybrid_LSQR gcv CT256/180 50 it", lambda: tb.Hybrid_LSQR(A, b, n_iter=50, regparam="gcv", x_true=xt), 50)
    run("cfg2 Hybrid_LSQR dp  CT256/180 50 it", lambda: tb.Hybrid_LSQR(A, b, n_iter=50, regparam="dp", delta=float(delta), x_true=xt), 50)
    # configs[2]
    n = 1024
    PSF = tb.gauss_psf((9, 9), (3, 3))
    Ab = tb.PSFBlur2D(PSF, n, n)
    xt = O.shepp_logan(n).reshape(-1, 1)
    b, delta = O.add_noise(O.blur_data(xt, PSF, n, n), 0.01, rng)
    for label, L, kw in (("anisoTV", tb.FirstDerivative2D(n, n), {}),
                         ("isoTV", tb.CenteredDerivative2D(n, n), {"isoTV": "isoTV", "prob_dims": (n, n, 1)})):
        run(f"cfg3 MMGKS {label} deblur 1024^2 50 it dp",
            lambda: tb.MMGKS(Ab, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=50, regparam="dp", delta=float(delta),
                             x_true=xt, **kw), 50)
    run("cfg3 Hybrid_GMRES deblur 1024^2 50 it dp", lambda: tb.Hybrid_GMRES(Ab, b, 50, regparam="dp", delta=float(delta), x_true=xt), 50)
    # configs[4]
    nx, nt, per = 256, 64, 12
    th = O.ct_angles(nt * per)
    frames = [th[t::nt] for t in range(nt)]
    Ad = tb.BlockDiagCT(nx, frames)
    base = O.shepp_logan(nx)
    xt = np.concatenate([(base * (1 + 0.3 * np.sin(2 * np.pi * t / nt))).ravel() for t in range(nt)]).reshape(-1, 1)
    b, delta = O.add_noise(Ad @ xt, 0.01, rng)
    L = tb.SpaceTimeDerivative(nx, nx, nt)
    print(f"cfg5 operator {Ad.shape}, nnz {Ad.nnz:.3e}, L rows {L.shape[0]}")
    run("cfg5 MMGKS space-time TV dynamic CT 256^2x64 50 it dp",
        lambda: tb.MMGKS(Ad, b, L, pnorm=2, qnorm=1, projection_dim=1, n_iter=50, regparam="dp", delta=float(delta),
                         epsilon=0.1, x_true=xt), 50)
    run("cfg5 Hybrid_LSQR gcv dynamic CT 50 it", lambda: tb.Hybrid_LSQR(Ad, b, n_iter=50, regparam="gcv", x_true=xt), 50)
    del Ad, L
    torch.cuda.empty_cache()


def cfg4(rng):
    # configs[3] = the headline geometry, whole solvers (matrix-free layout: 15.7 GB resident)
    nx, views = 2048, 720
    A4 = tb.ParallelBeamCT(nx, views, layout="implicit")
    xt = O.shepp_logan(nx).reshape(-1, 1)
    b, delta = O.add_noise(A4 @ xt, 0.01, rng)
    run("cfg4 Hybrid_LSQR dp  CT2048/720 50 it", lambda: tb.Hybrid_LSQR(A4, b, n_iter=50, regparam="dp", delta=float(delta), x_true=xt), 50)
    run("cfg4 Hybrid_LSQR gcv CT2048/720 50 it", lambda: tb.Hybrid_LSQR(A4, b, n_iter=50, regparam="gcv", x_true=xt), 50)
    run("cfg4 CGLS CT2048/720 50 it", lambda: tb.CGLS(A4, b, np.zeros((nx * nx, 1)), 50, 0.0, x_true=xt), 50)
    M4, rhs = A4.T @ A4, A4.T @ b
    for reorth in ("mgs", "cgs2"):
        run(f"cfg4 Hybrid_GMRES on A^T A ({reorth}) CT2048/720 50 it",
            lambda: tb.Hybrid_GMRES(M4, rhs, 50, regparam=1e-2, x_true=xt, b200_reorth=reorth), 50)
    L4 = tb.FirstDerivative2D(nx, nx)
    run("cfg4 MMGKS anisoTV CT2048/720 30 it dp",
        lambda: tb.MMGKS(A4, b, L4, pnorm=2, qnorm=1, projection_dim=3, n_iter=30, regparam="dp", delta=float(delta), x_true=xt), 30)


if __name__ == "__main__":
    main()
