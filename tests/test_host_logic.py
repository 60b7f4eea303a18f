"""Host-side logic of the product that needs no GPU: the NumPy parameter-choice rules against the golden values
of the reference, the reference's argument errors, operator bookkeeping."""
import numpy as np
import pytest

import trips_oracle as O
from conftest import GOLDEN


def test_solver_drivers_limit_the_host_blas_pool_only_for_the_call():
    """solvers/_common.single_threaded_host_blas: one BLAS thread inside the call (k x k LAPACK is slower, at some sizes
    pathologically, on a multi-threaded OpenBLAS), the caller's pool size restored afterwards, signature and name kept."""
    import inspect

    threadpoolctl = pytest.importorskip("threadpoolctl")
    import trips_b200 as tb
    from trips_b200.solvers._common import single_threaded_host_blas

    def blas_threads():
        return [p["num_threads"] for p in threadpoolctl.threadpool_info() if p["user_api"] == "blas"]

    before = blas_threads()
    seen = []

    @single_threaded_host_blas
    def probe(a, b=2):
        """doc"""
        seen.append(blas_threads())
        return a + b

    assert probe(1, b=3) == 4 and probe.__name__ == "probe" and probe.__doc__ == "doc"
    assert all(t == 1 for t in seen[0]) and blas_threads() == before
    with pytest.raises(ZeroDivisionError):  # restored after an exception too
        single_threaded_host_blas(lambda: 1 / 0)()
    assert blas_threads() == before
    for fn in (tb.Hybrid_LSQR, tb.Hybrid_GMRES, tb.GKS, tb.MMGKS):  # the reference signatures survive the decorator
        assert list(inspect.signature(fn).parameters)[:2] == ["A", "b"]


def test_incremental_gram_cache_semantics(monkeypatch):
    """kernels.IncrementalGram on the host side (the device pass replaced by NumPy): the panel pass is requested exactly
    when the same basis buffer has only gained columns and the extras are the same vectors; a shrunken k, another
    buffer or another extra vector fall back to the full pass; the merged matrix is always the full Gram matrix."""
    import torch
    import trips_b200  # noqa: F401
    from trips_b200 import kernels as K

    rng = np.random.default_rng(6)
    m, kmax = 50, 9
    data = torch.from_numpy(rng.standard_normal((kmax, m)))  # (kmax, m): column j of the basis is row j
    b1, b2 = torch.from_numpy(rng.standard_normal(m)), torch.from_numpy(rng.standard_normal(m))
    calls = []

    def fake_weighted_gram(B, k, w=None, extras=(), extra_weighted=(), comm=None, first_col=0):
        assert w is None
        calls.append(first_col)
        M = np.column_stack([B[:k].numpy().T] + [e.numpy() for e in extras])
        G = M.T @ M
        if first_col > 0:  # what the panel kernel leaves: only the columns >= first_col and their mirror rows
            P = np.zeros_like(G)
            P[:, first_col:] = G[:, first_col:]
            P[first_col:, :] = G[first_col:, :]
            G = P
        return G, np.zeros_like(G)

    monkeypatch.setattr(K, "weighted_gram", fake_weighted_gram)
    inc = K.IncrementalGram()

    def full(B, k, extras):
        M = np.column_stack([B[:k].numpy().T] + [e.numpy() for e in extras])
        return M.T @ M

    for k in (3, 4, 5, 7):
        hi, _ = inc.update(data, k, extras=(b1,), extra_weighted=(0,))
        assert np.array_equal(hi, full(data, k, (b1,)))
    assert calls == [0, 3, 4, 5]
    hi, _ = inc.update(data, 7, extras=(b1,), extra_weighted=(0,))  # nothing new: full pass (k did not grow)
    hi, _ = inc.update(data, 5, extras=(b1,), extra_weighted=(0,))  # truncated basis
    assert calls[-2:] == [0, 0] and np.array_equal(hi, full(data, 5, (b1,)))
    hi, _ = inc.update(data, 6, extras=(b2,), extra_weighted=(0,))  # another right-hand side
    assert calls[-1] == 0 and np.array_equal(hi, full(data, 6, (b2,)))
    hi, _ = inc.update(data, 8, extras=(b2,), extra_weighted=(0,))
    assert calls[-1] == 6 and np.array_equal(hi, full(data, 8, (b2,)))
    other = data.clone()
    hi, _ = inc.update(other, 9, extras=(b2,), extra_weighted=(0,))  # another buffer (a grown / re-allocated basis)
    assert calls[-1] == 0 and np.array_equal(hi, full(other, 9, (b2,)))


def test_householder_factor_matches_the_gram_route_conventions():
    """solvers/_gks_core.householder_factor (the fallback of the Gram / Cholesky route): R upper triangular with a
    non-negative diagonal, R^T R = M^T M, C = Q^T Z, residuals; a zero column gives a zero diagonal entry, not an error."""
    import trips_b200  # noqa: F401
    from trips_b200.solvers._gks_core import householder_factor

    rng = np.random.default_rng(4)
    M = rng.standard_normal((300, 7))
    M[:, 4] = 0.0
    Z = rng.standard_normal((300, 2))
    R, C, r2 = householder_factor(M, Z)
    assert np.array_equal(R, np.triu(R)) and np.all(np.diag(R) >= 0) and R[4, 4] == 0.0
    assert np.allclose(R.T @ R, M.T @ M)
    # (for a rank-deficient M, Householder's Q carries an arbitrary unit column for the zero pivot - as the reference's
    # Q_A does: the projection is onto k orthonormal columns either way)
    import scipy.linalg as la

    Q = la.qr(M, mode="economic")[0]
    assert np.allclose(r2, np.sum((Z - Q @ (Q.T @ Z)) ** 2, axis=0))
    assert np.allclose(np.sum(C ** 2, axis=0) + r2, np.sum(Z ** 2, axis=0))


def test_spectral_gcv_objective_equals_the_solve_based_form():
    """generalized_crossvalidation evaluates the objective from one SVD of R_A R_L^-1; gcv_value is the reference's
    two-solves-per-evaluation form (gcv.py:25-78).  Same function of lambda, square and (k+1) x k, L = I and triangular
    L, 'modified' trace size; a singular R_L falls back to the solve-based form."""
    from trips_b200.reg_param import gcv as G

    rng = np.random.default_rng(12)
    for k, rect, ident in ((6, False, False), (17, True, True), (40, False, False), (40, True, True), (9, True, False)):
        RA = np.triu(rng.standard_normal((k, k))) + np.diag(np.logspace(0.5, -2, k))
        if rect:
            RA = np.vstack((RA, np.zeros((1, k))))
            RA[k, k - 1] = 0.3
        RL = np.eye(k) if ident else np.triu(rng.standard_normal((k, k))) + 2 * np.eye(k)
        c = rng.standard_normal((RA.shape[0], 1))
        s2, ch2 = G._spectral_form(RA, RL, c)
        for trace_size in (RA.shape[0], 500):
            for lam in (1e-6, 1e-3, 0.2, 30.0):
                d = s2 + lam
                got = float(np.dot((lam / d) ** 2, ch2)) / (trace_size - float(np.sum(s2 / d))) ** 2
                assert got == pytest.approx(G.gcv_value(lam, RA, RL, c, trace_size), rel=1e-9)
    RL = np.triu(rng.standard_normal((5, 5)))
    RL[2, 2] = 0.0
    assert G._spectral_form(np.eye(5), RL, np.ones((5, 1))) is None


def test_product_gcv_matches_reference_values():
    from trips_b200.reg_param import generalized_crossvalidation
    from trips_b200.reg_param.gcv import gcv_value

    g = np.load(f"{GOLDEN}/regparam.npz")
    B, bhat, k = g["B"], g["bhat"], g["B"].shape[1]
    Q, s, _ = np.linalg.svd(B, full_matrices=False)
    # the noise-free fixture has its minimum at the lower bound 1e-9, where the 'standard' objective is rounding
    # noise (k - trace -> 0): compare the 'modified' minimiser, and objective VALUES for both variants
    lam = generalized_crossvalidation(Q, np.diag(s), np.eye(k), bhat, variant="modified", fullsize=500)
    assert lam == pytest.approx(float(g["lam_mod"]), rel=1e-6)
    c = (Q.T @ bhat).reshape(-1, 1)
    for lam in (1e-6, 1e-3, 1.0, 50.0):
        ref = O.gcv_numerator(lam, Q, np.diag(s), np.eye(k), bhat) / O.gcv_denominator(lam, np.diag(s), np.eye(k), bhat)
        assert gcv_value(lam, np.diag(s), np.eye(k), c, k) == pytest.approx(ref, rel=1e-9)
    # an interior minimum: decaying spectrum, noisy right-hand side (Brent resolves ~1e-8 relative: SURVEY.md F11)
    rng = np.random.default_rng(5)
    m = 300
    Qm = np.linalg.qr(rng.standard_normal((m, k)))[0]
    RA = np.diag(np.logspace(0, -3, k)) @ (np.eye(k) + 0.1 * np.triu(rng.standard_normal((k, k)), 1))
    bfull = Qm @ (RA @ np.ones((k, 1))) + 0.05 * rng.standard_normal((m, 1))
    sg = np.diag([1, -1, 1, -1, -1, 1.0])
    for RL in (np.eye(k), np.eye(k) + 0.3 * np.triu(np.ones((k, k)), 1)):
        want = O.generalized_crossvalidation(Qm, RA, RL, bfull)
        assert 1e-6 < want < 50  # interior minimum
        got = generalized_crossvalidation(None, RA, RL, Qm.T @ bfull)
        assert got == pytest.approx(want, rel=1e-6)
        # row-sign changes of the triangular factors (Cholesky vs Householder convention) cannot change the rule
        assert generalized_crossvalidation(None, sg @ RA, RL, sg @ (Qm.T @ bfull)) == pytest.approx(want, rel=1e-6)
    # 'modified' denominator with the 'standard' numerator (the reference's mix) sits at the lower bound here
    want = O.generalized_crossvalidation(Qm, RA, np.eye(k), bfull, variant="modified", fullsize=m)
    assert generalized_crossvalidation(None, RA, np.eye(k), Qm.T @ bfull, variant="modified", fullsize=m) == pytest.approx(want, rel=1e-6)


def test_product_discrepancy_principle_matches_reference_values():
    from trips_b200.reg_param import discrepancy_principle, discrepancy_principle_projected

    g = np.load(f"{GOLDEN}/regparam.npz")
    B, k = g["B"], g["B"].shape[1]
    assert discrepancy_principle(g["Qf"], B, None, g["bfull"], delta=0.5) == pytest.approx(float(g["lam_dp"]), rel=1e-11)
    Qk = g["Qf"][:, :k]
    lam = discrepancy_principle(Qk, g["RA"], g["RL"], g["bfull"], delta=0.8)
    assert lam == pytest.approx(float(g["lam_dp_L"]), rel=1e-11)
    bp = Qk.T @ g["bfull"]
    res = np.linalg.norm(g["bfull"] - Qk @ bp)
    assert discrepancy_principle_projected(g["RA"], g["RL"], bp, res, 0.8) == pytest.approx(float(g["lam_dp_L"]), rel=1e-11)
    # unreachable discrepancy => lambda = 0, like the reference
    assert discrepancy_principle_projected(g["RA"], g["RL"], bp, res, 1e-6) == 0
    with pytest.raises(Exception, match="noise level delta"):
        discrepancy_principle_projected(g["RA"], g["RL"], bp, res, None)


def test_product_l_curve_matches_reference_values():
    """trips/utilities/reg_param/l_curve.py: curvature values and the maximiser, against numbers produced by the real
    reference (tests/golden/regparam.npz), the oracle's restatement, and - in the build container - the live reference."""
    from trips_b200.reg_param import l_curve
    from trips_b200.reg_param.l_curve import l_curve_curvature

    g = np.load(f"{GOLDEN}/regparam.npz")
    RA, RL, k = g["RA"], g["RL"], g["RA"].shape[0]
    bp = g["Qf"][:, :k].T @ g["bfull"]
    for lam, want in zip(g["lc_grid"], g["lc_kappa"]):
        assert l_curve_curvature(lam, RA, RL, bp) == want  # same lstsq calls on the same matrices: same bits
        assert O.l_curve_curvature(lam, RA, RL, bp) == want
    assert l_curve(RA, RL, bp) == float(g["lam_lc_L"]) == O.l_curve(RA, RL, bp)
    Q, s, _ = np.linalg.svd(g["B"], full_matrices=False)
    c = Q.T @ g["bhat"].reshape(-1, 1)
    assert l_curve(np.diag(s), np.eye(k), c) == float(g["lam_lc_I"])
    # Cholesky vs Householder sign conventions of the triangular factor cannot change the rule
    sg = np.diag([1, -1, 1, -1, -1, 1.0])
    assert l_curve(sg @ RA, RL, sg @ bp) == pytest.approx(float(g["lam_lc_L"]), rel=1e-9)
    import ref_loader

    if ref_loader.available():
        import importlib

        ref_loader.load()
        lc = importlib.import_module("trips.utilities.reg_param.l_curve")
        rng = np.random.default_rng(9)
        A2, L2, b2 = np.triu(rng.standard_normal((5, 5))) + 3 * np.eye(5), np.eye(5), rng.standard_normal((5, 1))
        assert l_curve(A2, L2, b2) == lc.l_curve(A2, L2, b2)
        assert l_curve_curvature(0.05, A2, L2, b2) == lc.curvature(0.05, A2, L2, b2)


def test_framelet_analysis_matrices_match_oracle_and_reference():
    """trips/utilities/operators.py:50-101: the product's and the oracle's analysis matrices are the same matrix, equal -
    in the build container - to the reference's create_analysis_operator entry for entry; level 1 is a tight frame."""
    import warnings

    import scipy.sparse as sp
    from trips_b200.operators import framelet_analysis

    def canon(M):
        M = sp.csr_matrix(M)
        M.sort_indices()
        M.eliminate_zeros()
        return M

    import ref_loader

    ref_ops = None
    if ref_loader.available():
        import importlib

        ref_loader.load()
        ref_ops = importlib.import_module("trips.utilities.operators")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for n, levels in ((8, 1), (9, 2), (16, 3), (7, 2)):
            P, Q = canon(framelet_analysis(n, levels)), canon(O.framelet_analysis(n, levels))
            assert P.shape == ((2 * levels + 1) * n, n) and abs(P - Q).max() == 0
            if ref_ops is not None:
                assert abs(P - canon(ref_ops.create_analysis_operator(n, levels))).max() == 0
    W1 = framelet_analysis(12, 1)
    assert abs(W1.T @ W1 - sp.identity(12)).max() < 1e-15


def test_argument_errors_mirror_the_reference():
    """Raised before any device work (Hybrid_LSQR.py:59-61, Hybrid_GMRES.py:29-31, GKS.py:32-34)."""
    from trips_b200 import GKS, Hybrid_GMRES, Hybrid_LSQR

    A = np.eye(4)
    b = np.ones((4, 1))
    for call in (lambda: Hybrid_LSQR(A, b, 5, regparam="dp"), lambda: Hybrid_GMRES(A, b, 5, regparam="dp"),
                 lambda: GKS(A, b, A, regparam="dp")):
        with pytest.raises(Exception, match="noise level delta"):
            call()


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from trips_b200 import CGLS, as_operator

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        as_operator(np.eye(3))
    with pytest.raises(RuntimeError):
        CGLS(np.eye(3), np.ones(3), np.zeros((3, 1)), 2, 0)


def test_product_never_imports_the_oracle():
    import os
    import re

    from conftest import ROOT

    pkg = os.path.join(ROOT, "trips-py_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+(trips_oracle|ref_loader|oracle)\b", src, flags=re.M), f


def test_gauss_psf_and_geometry_helpers_match_oracle():
    from trips_b200.operators import ct_angles, ct_num_detectors, gauss_psf

    for dim, spread in (((9, 9), (3, 3)), ((5, 7), (1.5, 2.5)), ((4, 6), 2)):
        assert np.array_equal(gauss_psf(dim, spread), O.gauss_psf(dim, spread))
    assert np.array_equal(ct_angles(90), O.ct_angles(90)) and ct_num_detectors(2048) == 2896 == O.ct_num_detectors(2048)


def test_parameter_rules_agree_with_the_oracle_on_random_projected_problems():
    """Seeded sweep: the product's host-side rules (fed the PROJECTED quantities the GPU solvers hand them) against the
    oracle's statements of the reference rules (fed the full-length ones), k = 2..9, triangular and diagonal factors."""
    from trips_b200.reg_param import discrepancy_principle_projected, generalized_crossvalidation, l_curve

    rng = np.random.default_rng(31)
    checked = 0
    for trial in range(12):
        k, m = int(rng.integers(2, 10)), 60
        Q = np.linalg.qr(rng.standard_normal((m, k)))[0]
        RA = np.triu(rng.standard_normal((k, k))) + np.diag(np.logspace(0.5, -2, k))
        RL = np.eye(k) if trial % 3 == 0 else np.triu(rng.standard_normal((k, k))) + 2 * np.eye(k)
        x = rng.standard_normal((k, 1))
        noise = rng.standard_normal((m, 1))
        b = Q @ (RA @ x) + 0.05 * noise
        c = Q.T @ b
        resid = float(np.linalg.norm(b - Q @ c))
        # GCV (the rule sees b only through Q^T b and k)
        want = O.generalized_crossvalidation(Q, RA, RL, b)
        got = generalized_crossvalidation(None, RA, RL, c)
        # Brent's method on a flat objective fixes lambda only loosely: compare the objective VALUES at
        # the two minimisers, with the oracle's own objective
        f = lambda lam: O.gcv_numerator(lam, Q, RA, RL, b) / O.gcv_denominator(lam, RA, RL, b)  # noqa: E731
        assert f(got) == pytest.approx(f(want), rel=1e-5), (trial, "gcv")
        # discrepancy principle on the projected problem
        delta = 1.02 * resid + 0.02
        want = O.discrepancy_principle(Q, RA, RL, b, delta=delta)
        got = discrepancy_principle_projected(RA, RL, c, resid, delta)
        assert got == pytest.approx(want, rel=1e-9, abs=1e-14), (trial, "dp")
        # L-curve: same function of (R_A, R_L, Q^T b)
        assert l_curve(RA, RL, c) == O.l_curve(RA, RL, c), (trial, "l_curve")
        checked += 1
    assert checked == 12


def test_host_basis_appends_in_place_only_behind_the_newest_view():
    """golub_kahan_update / arnoldi_update grow the caller's basis inside a capacity buffer.  Writing column
    k in place is only allowed when the array handed in is the newest view (k == high-water mark); an older or
    truncated view gets a fresh copy, so arrays already returned are never mutated (the reference returns fresh arrays
    from np.hstack, decompositions.py:243-247)."""
    from trips_b200.decompositions import _HostBasis, release_host_buffers

    release_host_buffers()
    U0 = np.arange(6.0).reshape(6, 1)
    hb, k = _HostBasis.adopt(U0)
    assert k == 1 and np.array_equal(hb.arr[:, :1], U0)
    hb.arr[:, 1] = 10.0
    U1 = hb.view(2)
    hb2, k2 = _HostBasis.adopt(U1)          # newest view: same buffer, append in place
    assert hb2 is hb and k2 == 2
    hb.arr[:, 2] = 20.0
    U2 = hb.view(3)
    hb3, k3 = _HostBasis.adopt(U1)          # an OLDER view (restart / branch): must not overwrite U2's column 2
    assert hb3 is not hb and k3 == 2
    hb3.arr[:, 2] = -1.0
    assert np.array_equal(U2[:, 2], np.full(6, 20.0))
    hb4, _ = _HostBasis.adopt(U2[:, :2])    # a truncated view shares the data pointer: also copied
    assert hb4 is not hb
    # eviction: the registry never pins more than MAX_LIVE buffers, and release drops them all
    for i in range(8):
        _HostBasis.adopt(np.full((3, 1), float(i)))
    assert len(_HostBasis.registry) <= _HostBasis.MAX_LIVE
    release_host_buffers()
    assert not _HostBasis.registry
    assert np.array_equal(U2[:, 2], np.full(6, 20.0))  # returned arrays stay valid


def test_forward_cta_order_is_a_permutation_heaviest_first():
    """kernels.forward_cta_order (scheduling list of the ray-driven forward projector for small launches): a permutation of
    all (angle, block) pairs, central blocks of near-diagonal angles first, empty edge blocks last."""
    import torch
    from trips_b200 import kernels as K

    nx, views, n_det = 512, 24, 724
    th = np.linspace(0, np.pi, views, endpoint=False)
    order = K.forward_cta_order(nx, nx, n_det, torch.from_numpy(np.cos(th)), torch.from_numpy(np.sin(th)), "cpu", force=True)
    o = order.numpy()
    nblk = o.size // views
    assert o.size == views * nblk and sorted(o.tolist()) == list(range(o.size))
    first_blocks, last_blocks = o[:views] % nblk, o[-views:] % nblk
    assert np.all(np.abs(first_blocks - (nblk - 1) / 2) <= nblk / 4 + 1)   # heavy = near the detector centre
    assert np.all(np.minimum(last_blocks, nblk - 1 - last_blocks) <= 1)    # light = the detector's ends


def test_band_rows_partition_the_image():
    from trips_b200.dist import band_rows, shard_angles

    for ny, world in ((2048, 8), (52, 2), (30, 4), (7, 3), (4096, 16)):
        edges = [band_rows(ny, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == ny
        assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
        assert all(lo % 4 == 0 for lo, _ in edges)
    assert sorted(np.concatenate([shard_angles(45, 4, r) for r in range(4)]).tolist()) == list(range(45))
