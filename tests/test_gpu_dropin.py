"""GPU tests of the drop-in claim: the REAL reference's own solver loops
(oracle/_ref = unmodified pip install of mpasha3/trips-py, loaded through oracle/ref_loader.py) run with trips_b200
operators handed in as `A` / `L`, and must return what they return on the scipy matrices / scipy.ndimage operator the
reference normally gets - bit for bit, because every trips_b200 product is bit-identical to scipy's.

Reference loops exercised: trips/solvers/Hybrid_LSQR.py:25-114, CGLS.py:16-86, Hybrid_GMRES.py:23-87, GKS.py:27-105,
MMGKS.py:28-137, trips/utilities/decompositions.py:118-255.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

import trips_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import trips_b200

    return trips_b200


@pytest.fixture(scope="module")
def ref():
    import ref_loader

    if not ref_loader.available():
        script = os.path.join(ROOT, "oracle", "make_ref.sh")
        if os.path.isdir("/root/reference/trips"):
            subprocess.run(["bash", script], check=True)
        if not ref_loader.available():
            pytest.fail("oracle/_ref is missing: run __graft_entry__.build() (oracle/make_ref.sh) where /root/reference exists")
    return ref_loader.load()


def ct_problem(tb, nx, views, seed=2022):
    op = tb.ParallelBeamCT(nx, views)
    A = op.to_scipy()
    xt = O.shepp_logan(nx).reshape((-1, 1))
    b, delta = O.add_noise(A @ xt, 0.01, np.random.default_rng(seed))
    return op, A, xt, b, float(delta)


def test_reference_is_the_unmodified_package(ref):
    import ref_loader

    assert os.path.isdir(os.path.join(ref_loader.REFERENCE_ROOT, "trips", "solvers"))
    assert ref.decompositions.golub_kahan_update.__module__ == "trips.utilities.decompositions"


@pytest.mark.parametrize("layout", ["sell", "implicit"])
def test_reference_krylov_cores_on_our_operator(tb, ref, layout):
    """decompositions.py:230-255 / :118-205 / :207-228 of the reference, A = tb.ParallelBeamCT."""
    op, A, xt, b, delta = ct_problem(tb, 48, 40)
    op = tb.ParallelBeamCT(48, 40, layout=layout)
    U0 = b / np.linalg.norm(b)
    Ur, Sr, Vr = U0, np.empty(1), np.empty((A.shape[1], 1))
    Ug, Sg, Vg = U0, np.empty(1), np.empty((A.shape[1], 1))
    for _ in range(8):
        Ur, Sr, Vr = ref.decompositions.golub_kahan_update(A, Ur, Sr, Vr)
        Ug, Sg, Vg = ref.decompositions.golub_kahan_update(op, Ug, Sg, Vg)
    assert np.array_equal(Sr, Sg) and np.array_equal(Ur, Ug) and np.array_equal(Vr, Vg)
    Ur, Sr, Vr = ref.decompositions.golub_kahan(A, b, 6)
    Ug, Sg, Vg = ref.decompositions.golub_kahan(op, b, 6)
    assert np.array_equal(Sr, Sg) and np.array_equal(Ur, Ug) and np.array_equal(Vr, Vg)
    # Arnoldi on the square normal operator: scipy A.T @ A vs the product of our operators
    M, Mg = (A.T @ A).tocsr(), op.T @ op
    v0 = (A.T @ b) / np.linalg.norm(A.T @ b)
    Vr, Hr = v0, np.empty(1)
    Vg, Hg = v0, np.empty(1)
    for _ in range(5):
        Vr, Hr = ref.decompositions.arnoldi_update(M, Vr, Hr)
        Vg, Hg = ref.decompositions.arnoldi_update(Mg, Vg, Hg)
    # (A^T A) v as one scipy matrix sums in another order than A^T (A v): agreement to rounding, not bitwise
    assert np.allclose(Hr, Hg, rtol=1e-9, atol=1e-9 * np.abs(Hr).max())


def test_reference_solvers_on_our_ct_operator(tb, ref):
    """The reference's CGLS / Hybrid_LSQR / GKS / MMGKS loops, A = tb.ParallelBeamCT, L = tb.FirstDerivative2D."""
    nx = 48
    op, A, xt, b, delta = ct_problem(tb, nx, 40)
    mf = tb.ParallelBeamCT(nx, 40, layout="implicit")
    L, Lo = tb.FirstDerivative2D(nx, nx), O.first_derivative_2d(nx, nx)
    for name, our in (("stored", op), ("matrix-free", mf)):
        xr, ir = ref.CGLS(A, b, np.zeros((A.shape[1], 1)), 20, 0, x_true=xt)
        xg, ig = ref.CGLS(our, b, np.zeros((A.shape[1], 1)), 20, 0, x_true=xt)
        assert np.array_equal(xr, xg), name
        xr, ir = ref.Hybrid_LSQR(A, b, n_iter=20, regparam="dp", x_true=xt, delta=delta)
        xg, ig = ref.Hybrid_LSQR(our, b, n_iter=20, regparam="dp", x_true=xt, delta=delta)
        assert np.array_equal(xr, xg) and ir["regParam_history"] == ig["regParam_history"], name
        xr, ir = ref.Hybrid_LSQR(A, b, n_iter=12, regparam="gcv", x_true=xt)
        xg, ig = ref.Hybrid_LSQR(our, b, n_iter=12, regparam="gcv", x_true=xt)
        assert np.array_equal(xr, xg), name
        xr, ir = ref.GKS(A, b, Lo, projection_dim=3, n_iter=10, regparam="dp", x_true=xt, delta=delta)
        xg, ig = ref.GKS(our, b, L, projection_dim=3, n_iter=10, regparam="dp", x_true=xt, delta=delta)
        assert np.array_equal(xr, xg) and np.array_equal(ir["relError"], ig["relError"]), name
        xr, ir = ref.MMGKS(A, b, Lo, pnorm=2, qnorm=1, projection_dim=3, n_iter=10, regparam="dp", x_true=xt, delta=delta)
        xg, ig = ref.MMGKS(our, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=10, regparam="dp", x_true=xt, delta=delta)
        assert np.array_equal(xr, xg) and ir["regParam_history"] == ig["regParam_history"], name
    # and the result of the reference loop on our operator is what our own solver returns (to the BLAS-vs-exact norms)
    xg, _ = ref.MMGKS(mf, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=10, regparam="dp", x_true=xt, delta=delta)
    xo, _ = tb.MMGKS(mf, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=10, regparam="dp", x_true=xt, delta=delta)
    assert np.linalg.norm(xg - xo) <= 1e-10 * np.linalg.norm(xg)


def test_reference_solvers_on_our_blur_operator(tb, ref):
    """Deblurring: the reference's Deblurring2D.forward_Op (scipy.ndimage behind a pylops FunctionOperator,
    Deblurring2D.py:66-73) vs tb.PSFBlur2D inside the reference's Hybrid_GMRES / Hybrid_LSQR / MMGKS."""
    n = 64
    D = ref.Deblurring2D(CommitCrime=True)
    Aref = D.forward_Op((9, 9), (2, 2), n, n)
    PSF = O.gauss_psf((9, 9), (2, 2))
    assert np.array_equal(PSF, D.Gauss((9, 9), (2, 2))[0])
    op = tb.PSFBlur2D(PSF, n, n)
    xt = O.shepp_logan(n).reshape((-1, 1))
    b, delta = O.add_noise(Aref @ xt, 0.01, np.random.default_rng(3))
    delta = float(delta)
    xr, ir = ref.Hybrid_GMRES(Aref, b, n_iter=15, regparam="dp", x_true=xt, delta=delta)
    xg, ig = ref.Hybrid_GMRES(op, b, n_iter=15, regparam="dp", x_true=xt, delta=delta)
    assert np.array_equal(xr, xg)
    xr, ir = ref.Hybrid_LSQR(Aref, b, n_iter=15, regparam=1e-3, x_true=xt)
    xg, ig = ref.Hybrid_LSQR(op, b, n_iter=15, regparam=1e-3, x_true=xt)
    assert np.array_equal(xr, xg)
    L, Lo = tb.FirstDerivative2D(n, n), O.first_derivative_2d(n, n)
    xr, ir = ref.MMGKS(Aref, b, Lo, pnorm=2, qnorm=1, projection_dim=3, n_iter=8, regparam="dp", x_true=xt, delta=delta)
    xg, ig = ref.MMGKS(op, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=8, regparam="dp", x_true=xt, delta=delta)
    assert np.array_equal(xr, xg)


def test_golub_kahan_dp_stop(tb, ref):
    """decompositions.py:164-195: discrepancy-principle stop inside golub_kahan - same stopping step and factors as the
    oracle (pinned to the reference in tests/test_oracle.py) and as the reference itself on our operator."""
    op, A, xt, b, delta = ct_problem(tb, 32, 30)
    for gk_delta in (1e-3, 3.0):
        with O.reductions("exact"):
            Uo, So, Vo = O.golub_kahan(A, b, 25, True, gk_delta=gk_delta)
        U, S, V = tb.golub_kahan(op, b, 25, dp_stop=True, gk_delta=gk_delta)
        assert S.shape == So.shape, (S.shape, So.shape)
        assert np.array_equal(S, So) and np.array_equal(U, Uo) and np.array_equal(V, Vo)
        Ur, Sr, Vr = ref.decompositions.golub_kahan(op, b, 25, dp_stop=True, gk_delta=gk_delta)
        assert Sr.shape == S.shape
    assert So.shape[1] < 25  # gk_delta = 3.0 stops early
    L = tb.FirstDerivative2D(32, 32)
    with pytest.raises(TypeError, match="multiple values"):  # the reference forwards dp_stop twice (GKS.py:36)
        tb.GKS(op, b, L, projection_dim=3, n_iter=3, regparam=0.1, dp_stop=True, delta=delta)
    with pytest.raises(TypeError, match="multiple values"):
        ref.GKS(A, b, O.first_derivative_2d(32, 32), projection_dim=3, n_iter=3, regparam=0.1, dp_stop=True, delta=delta)
    with pytest.raises(TypeError, match="multiple values"):
        tb.MMGKS(op, b, L, projection_dim=3, n_iter=3, regparam=0.1, dp_stop=False)


def test_tensors_of_another_device_are_refused(tb):
    """kernels launch on the current device's stream; a tensor of another GPU must raise, not be dereferenced."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    with torch.cuda.device(1):
        op = tb.ParallelBeamCT(32, 10, device="cuda:1")
        x1 = torch.ones(op.shape[1], dtype=torch.float64, device="cuda:1")
        y = op.apply_dev(x1)
        assert y.device.index == 1
    with pytest.raises(RuntimeError, match="current CUDA device"):
        op.apply_dev(x1)
