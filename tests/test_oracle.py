"""The oracle (oracle/trips_oracle.py) against (a) the golden vectors produced by the real reference and
(b) the real reference itself when /root/reference is present (build container only).  CPU only."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

import ref_loader
import trips_oracle as O
from conftest import GOLDEN

warnings.filterwarnings("ignore", category=SyntaxWarning)


def load(name):
    return np.load(f"{GOLDEN}/{name}.npz")


def csr(g, p="A"):
    return sp.csr_matrix((g[p + "_data"], g[p + "_indices"], g[p + "_indptr"]), shape=tuple(g[p + "_shape"]))


def test_ct_matrix_matches_fixture_and_geometry():
    g = load("ct24")
    A = O.ct_matrix(int(g["nx"]), O.ct_angles(int(g["views"])))
    A0 = csr(g)
    assert (A != A0).nnz == 0
    # every ray's weights sum to its chord through the image square; all weights in (0, sqrt 2]
    assert A.data.min() > 0 and A.data.max() <= np.sqrt(2) + 1e-12
    # 45 degrees: ray d crosses the square [-12,12]^2 with chord sqrt(2)*(24 - |s_d|*sqrt(2)); s_d = d - 16
    n_det = O.ct_num_detectors(24)
    sums = np.asarray(A[4 * n_det:5 * n_det].sum(axis=1)).ravel()
    expect = np.sqrt(2) * (24 - np.abs(np.arange(n_det) - 16) * np.sqrt(2))
    assert np.allclose(sums, expect, atol=1e-12)
    # n_det odd + nx even: at 0 degrees every ray runs exactly along a pixel boundary (measure-zero
    # intersection with the open pixels) => those rows are empty; 2048/256/64 (n_det even) do not degenerate
    assert A[:n_det].nnz == 0


def test_golden_gk_cgls_hybrid():
    g = load("ct24")
    A, b, xt = csr(g), g["b"], g["x_true"]
    U = b / np.linalg.norm(b)
    B, V = np.empty(1), np.empty((A.shape[1], 1))
    for _ in range(8):
        U, B, V = O.golub_kahan_update(A, U, B, V)
    assert np.array_equal(U, g["gk_U"]) and np.array_equal(B, g["gk_B"]) and np.array_equal(V, g["gk_V"])
    Ub, Sb, Vb = O.golub_kahan(A, b, 8)
    assert np.array_equal(Ub, U) and np.array_equal(Sb, B) and np.array_equal(Vb, V)
    x, info = O.CGLS(A, b, np.zeros((A.shape[1], 1)), 20, 0, x_true=xt)
    assert np.array_equal(x, g["cgls_x"]) and np.array_equal(info["relResidual"], g["cgls_relres"])
    assert np.array_equal(info["relError"], g["cgls_relerr"])
    for tag, rp, kw in (("fix", 1e-2, {}), ("dp", "dp", {"delta": float(g["delta"])}), ("gcv", "gcv", {})):
        x, info = O.Hybrid_LSQR(A, b, n_iter=20, regparam=rp, x_true=xt, **kw)
        assert np.array_equal(x, g[f"hlsqr_{tag}_x"]), tag
        assert np.array_equal(np.array(info["regParam_history"], dtype=float), g[f"hlsqr_{tag}_lam"])
        assert np.array_equal(info["relError"], g[f"hlsqr_{tag}_relerr"])
    M, rhs = (A.T @ A).tocsr(), A.T @ b
    for tag, rp, kw in (("fix", 1e-2, {}), ("dp", "dp", {"delta": float(g["hgmres_dp_delta"])})):
        x, info = O.Hybrid_GMRES(M, rhs, 15, regparam=rp, **kw)
        assert np.array_equal(x, g[f"hgmres_{tag}_x"]), tag


def test_golden_gks_mmgks():
    g = load("ct24")
    A, b = csr(g), g["b"]
    L = O.first_derivative_2d(int(g["nx"]), int(g["nx"]))
    for tag, rp, kw in (("fix", 1e-1, {}), ("dp", "dp", {"delta": float(g["delta"])}), ("gcv", "gcv", {})):
        x, info = O.GKS(A, b, L, projection_dim=3, n_iter=15, regparam=rp, **kw)
        assert np.array_equal(x, g[f"gks_{tag}_x"]) and np.array_equal(info["Residual"], g[f"gks_{tag}_res"])
        x, info = O.MMGKS(A, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=15, regparam=rp, **kw)
        assert np.array_equal(x, g[f"mmgks_{tag}_x"]) and np.array_equal(info["Residual"], g[f"mmgks_{tag}_res"])
    x, info = O.MMGKS(A, b, L, pnorm=1.5, qnorm=0.8, projection_dim=2, n_iter=10, regparam=1e-1, epsilon=0.05)
    assert np.array_equal(x, g["mmgks_pq_x"])


def test_golden_deblurring():
    g = load("deblur32")
    n, PSF = int(g["n"]), g["PSF"]
    assert np.array_equal(O.gauss_psf((7, 7), (2, 2)), PSF)
    A = O.blur_operator(PSF, n, n)
    assert np.array_equal(A @ g["probe"], g["fwd_probe"]) and np.array_equal(A.T @ g["probe"], g["adj_probe"])
    assert np.array_equal(O.blur_data(g["x_true"], PSF, n, n), g["b_true"])
    x, info = O.Hybrid_LSQR(A, g["b"], n_iter=15, regparam="dp", delta=float(g["delta"]))
    assert np.array_equal(x, g["hlsqr_dp_x"])
    x, _ = O.Hybrid_GMRES(A, g["b"], 15, regparam=1e-3)
    assert np.array_equal(x, g["hgmres_fix_x"])
    L = O.first_derivative_2d(n, n)
    x, info = O.MMGKS(A, g["b"], L, pnorm=2, qnorm=1, projection_dim=3, n_iter=12, regparam="dp", delta=float(g["delta"]))
    assert np.array_equal(x, g["mmgks_dp_x"])
    x, _ = O.GKS(A, g["b"], L, projection_dim=3, n_iter=12, regparam=1e-2)
    assert np.array_equal(x, g["gks_fix_x"])


def test_golden_parameter_rules():
    g = load("regparam")
    B, bhat, k = g["B"], g["bhat"], g["B"].shape[1]
    Q, s, _ = np.linalg.svd(B, full_matrices=False)
    assert O.generalized_crossvalidation(Q, np.diag(s), np.eye(k), bhat) == float(g["lam_std"])
    assert O.generalized_crossvalidation(Q, np.diag(s), np.eye(k), bhat, variant="modified", fullsize=500) == float(g["lam_mod"])
    assert O.discrepancy_principle(g["Qf"], B, O.IdentityOp(k), g["bfull"], delta=0.5) == float(g["lam_dp"])
    assert O.discrepancy_principle(g["Qf"][:, :k], g["RA"], g["RL"], g["bfull"], delta=0.8) == float(g["lam_dp_L"])
    assert O.generalized_crossvalidation(g["Qf"][:, :k], g["RA"], g["RL"], g["bfull"]) == float(g["lam_gcv_L"])


def test_difference_operators_known_answers():
    """The reference's own builders raise under this scipy; check the CSR restatement by definition."""
    L = O.first_derivative_1d(5).toarray()
    assert np.array_equal(L, np.eye(5)[:4] - np.eye(5, k=1)[:4])
    n = 4
    L2 = O.first_derivative_2d(n, n)
    assert L2.shape == (2 * n * (n - 1), n * n)
    X = np.arange(16.0).reshape(4, 4) ** 2
    u = L2 @ X.ravel()
    assert np.array_equal(u[:12].reshape(4, 3), X[:, :-1] - X[:, 1:])
    assert np.array_equal(u[12:].reshape(3, 4), X[:-1, :] - X[1:, :])
    L3 = O.spacetime_derivative(n, n, 3)
    assert L3.shape == (3 * 24 + 2 * 16, 48)
    x = np.random.default_rng(0).standard_normal(48)
    u = L3 @ x
    F = x.reshape(3, 16)
    assert np.array_equal(u[72:].reshape(2, 16), F[:-1] - F[1:])
    assert np.array_equal(u[:24], L2 @ F[0])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_oracle_is_bit_identical_to_the_reference():
    ref = ref_loader.load()
    rng = np.random.default_rng(7)
    A = sp.random(220, 150, density=0.06, random_state=3, format="csr")
    b = rng.standard_normal((220, 1))
    xt = rng.standard_normal((150, 1))
    same = lambda a, c: np.array_equal(np.asarray(a), np.asarray(c))  # noqa: E731
    U = b / np.linalg.norm(b)
    S, V = np.empty(1), np.empty((150, 1))
    U2, S2, V2 = U.copy(), S.copy(), V.copy()
    for _ in range(7):
        U, S, V = ref.decompositions.golub_kahan_update(A, U, S, V)
        U2, S2, V2 = O.golub_kahan_update(A, U2, S2, V2)
    assert same(U, U2) and same(S, S2) and same(V, V2)
    assert all(same(p, q) for p, q in zip(ref.decompositions.golub_kahan(A, b, 4), O.golub_kahan(A, b, 4)))
    M, rhs = (A.T @ A).tocsr(), A.T @ b
    Vr, H = rhs / np.linalg.norm(rhs), np.empty(1)
    Vo, Ho = Vr.copy(), H.copy()
    for _ in range(7):
        Vr, H = ref.decompositions.arnoldi_update(M, Vr, H)
        Vo, Ho = O.arnoldi_update(M, Vo, Ho)
    assert same(Vr, Vo) and same(H, Ho)
    x, i = ref.CGLS(A, b, np.zeros((150, 1)), 12, 0, x_true=xt)
    x2, i2 = O.CGLS(A, b, np.zeros((150, 1)), 12, 0, x_true=xt)
    assert same(x, x2) and same(i["relResidual"], i2["relResidual"]) and same(i["relError"], i2["relError"])
    L = O.first_derivative_1d(150)
    for rp, kw in (("gcv", {}), ("dp", {"delta": 0.4}), (0.02, {})):
        x, i = ref.Hybrid_LSQR(A, b, n_iter=10, regparam=rp, x_true=xt, **kw)
        x2, i2 = O.Hybrid_LSQR(A, b, n_iter=10, regparam=rp, x_true=xt, **kw)
        assert same(x, x2) and same(i["regParam_history"], i2["regParam_history"]) and same(i["relError"], i2["relError"])
        x, i = ref.Hybrid_GMRES(M, rhs, 10, regparam=rp, **kw)
        x2, i2 = O.Hybrid_GMRES(M, rhs, 10, regparam=rp, **kw)
        assert same(x, x2) and same(i["regParam_history"], i2["regParam_history"])
        x, i = ref.GKS(A, b, L, projection_dim=2, n_iter=8, regparam=rp, **kw)
        x2, i2 = O.GKS(A, b, L, projection_dim=2, n_iter=8, regparam=rp, **kw)
        assert same(x, x2) and same(i["Residual"], i2["Residual"])
        for pn, qn in ((2, 1), (1.2, 0.7)):
            x, i = ref.MMGKS(A, b, L, pnorm=pn, qnorm=qn, projection_dim=2, n_iter=8, regparam=rp, **kw)
            x2, i2 = O.MMGKS(A, b, L, pnorm=pn, qnorm=qn, projection_dim=2, n_iter=8, regparam=rp, **kw)
            assert same(x, x2) and same(i["regParam_history"], i2["regParam_history"])
    assert same(ref.phantoms.shepp_logan(48), O.shepp_logan(48))
    D = ref.Deblurring2D(CommitCrime=False)
    Ar = D.forward_Op((5, 7), (1.5, 2.5), 20, 20)
    PSF = O.gauss_psf((5, 7), (1.5, 2.5))
    Ao = O.blur_operator(PSF, 20, 20)
    z = rng.standard_normal((400, 1))
    assert same(Ar @ z, Ao @ z) and same(Ar.T @ z, Ao.T @ z) and same(D.gen_data(z), O.blur_data(z, PSF, 20, 20))
    # the fp64 centred-difference statement vs the pylops stand-in evaluated in float64
    Lc = O.centered_derivative_2d(6, 6)
    pl = ref.pylops
    Dx = pl.FirstDerivative(6, dtype="float64")
    Ls = pl.VStack((pl.Kronecker(pl.Identity(6), Dx), pl.Kronecker(Dx, pl.Identity(6))))
    w = rng.standard_normal(36)
    assert np.allclose(Lc @ w, Ls.matvec(w), atol=1e-15) and np.allclose(Lc.T @ np.r_[w, w], Ls.rmatvec(np.r_[w, w]), atol=1e-15)


def _clip_length(c, s, rho, half_w, half_h):
    """Length of the line {p : (c, s).p = rho} inside the rectangle [-half_w, half_w] x [-half_h, half_h]
    (Liang-Barsky on the parametrisation p = rho*(c, s) + tau*(-s, c)): an independent, closed-form answer."""
    px, py, ex, ey = rho * c, rho * s, -s, c
    lo, hi = -np.inf, np.inf
    for p0, e, h in ((px, ex, half_w), (py, ey, half_h)):
        if abs(e) < 1e-300:
            if abs(p0) >= h:
                return 0.0
            continue
        t0, t1 = (-h - p0) / e, (h - p0) / e
        lo, hi = max(lo, min(t0, t1)), min(hi, max(t0, t1))
    return max(hi - lo, 0.0)


@pytest.mark.parametrize("fan", [None, "reference", (40.0, 0.0, 0.8)])
def test_ct_matrix_rows_sum_to_the_ray_length_inside_the_image(fan):
    """Physics check of the system matrix that does not depend on any implementation of it: the chord lengths of one
    ray through all unit pixels add up to the length of the ray inside the image rectangle.  Holds for the parallel beam
    and for the fan beam (every ray with its own normal and offset: ASTRA 'fanflat' conventions, Tomography.py:57-67),
    on a non-square image, including rays that clip corners and rays that miss the image."""
    nx, ny, views = 20, 14, 11
    theta = np.linspace(0, np.pi, views, endpoint=False) + 0.013
    n_det = 41
    geom = O.fan_geometry(nx) if fan == "reference" else fan
    A = O.ct_matrix(nx, theta, ny=ny, n_det=n_det, fan=geom)
    got = np.asarray(A @ np.ones(nx * ny)).ravel()
    want = np.empty_like(got)
    for a, th in enumerate(theta):
        cd, sd, rho, *_ = O._ray_tables(np.cos(th), np.sin(th), n_det, geom)
        for d in range(n_det):
            want[a * n_det + d] = _clip_length(cd[d], sd[d], rho[d], nx / 2, ny / 2)
    assert np.allclose(got, want, rtol=0, atol=1e-11)
    assert (want == 0).any() and want.max() > min(nx, ny)  # the fixture has missing rays and long central ones
    # and every entry is a genuine chord: positive, at most sqrt(2)
    assert A.data.min() > 0 and A.data.max() <= np.sqrt(2) + 1e-15


def test_golden_fan_beam_and_l_curve_inside_the_solvers():
    """Outputs of the REAL reference's solvers on the fan-beam matrix (its own ASTRA geometry, built by O.ct_matrix) with
    regparam='l_curve' (trips/utilities/reg_param/l_curve.py through Hybrid_LSQR.py:94-98, GKS.py:67-68,
    MMGKS.py:100-101): the oracle reproduces them bit for bit, and the matrix in the fixture is the oracle's."""
    g = load("ctfan20")
    A, b, xt = csr(g), g["b"], g["x_true"]
    nx, views = int(g["nx"]), int(g["views"])
    A0 = O.ct_matrix(nx, O.ct_angles(views), fan=O.fan_geometry(nx))
    assert np.array_equal(A0.indptr, A.indptr) and np.array_equal(A0.indices, A.indices) and np.array_equal(A0.data, A.data)
    x, info = O.CGLS(A, b, np.zeros((A.shape[1], 1)), 15, 0, x_true=xt)
    assert np.array_equal(x, g["cgls_x"]) and np.array_equal(info["relError"], g["cgls_relerr"])
    x, info = O.Hybrid_LSQR(A, b, n_iter=12, regparam="l_curve", x_true=xt)
    assert np.array_equal(np.array(info["regParam_history"], dtype=float), g["hlsqr_lc_lam"]) and np.array_equal(x, g["hlsqr_lc_x"])
    L = O.first_derivative_2d(nx, nx)
    x, info = O.GKS(A, b, L, projection_dim=3, n_iter=8, regparam="l_curve")
    assert np.array_equal(np.array(info["regParam_history"], dtype=float), g["gks_lc_lam"]) and np.array_equal(x, g["gks_lc_x"])
    x, info = O.MMGKS(A, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=8, regparam="l_curve")
    assert np.array_equal(np.array(info["regParam_history"], dtype=float), g["mmgks_lc_lam"]) and np.array_equal(x, g["mmgks_lc_x"])


def test_oracle_golub_kahan_dp_stop_is_bit_identical_to_the_reference():
    """decompositions.py:164-195 (discrepancy-principle stop inside golub_kahan): same stopping step, same factors; and
    the reference's GKS forwards dp_stop twice (GKS.py:36) - a TypeError the oracle reproduces."""
    import ref_loader

    if not ref_loader.available():
        pytest.skip("reference not available (neither oracle/_ref nor /root/reference)")
    R = ref_loader.load()
    A = O.ct_matrix(24, O.ct_angles(20))
    x = O.shepp_logan(24).reshape(-1, 1)
    b, delta = O.add_noise(A @ x, 0.01, np.random.default_rng(1))
    for gd in (0.001, 2.0):
        Ur, Sr, Vr = R.decompositions.golub_kahan(A, b, 30, dp_stop=True, gk_delta=gd)
        Uo, So, Vo = O.golub_kahan(A, b, 30, True, gk_delta=gd)
        assert Sr.shape == So.shape
        assert np.array_equal(Ur, Uo) and np.array_equal(Sr, So) and np.array_equal(Vr, Vo)
    assert So.shape[1] < 30
    L = O.first_derivative_2d(24, 24)
    for fn in (R.GKS, O.GKS, R.MMGKS, O.MMGKS):
        with pytest.raises(TypeError, match="multiple values"):
            fn(A, b, L, projection_dim=3, n_iter=2, regparam=0.01, dp_stop=True, delta=float(delta))
