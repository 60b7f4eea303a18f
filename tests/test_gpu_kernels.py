"""Kernel-level parity on the B200: every CUDA kernel, called through the C ABI (trips_b200.kernels -> ctypes ->
libtripsb200.so), against NumPy/SciPy (the arithmetic the reference delegates to) on the same seeded inputs.
Integer/index results and the element-wise / stencil kernels must be bit-exact; reductions agree to rounding."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.ndimage import convolve

import trips_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import torch
    import trips_b200

    assert torch.cuda.is_available()
    return trips_b200


def dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def rel(a, b):
    return np.linalg.norm(np.asarray(a).ravel() - np.asarray(b).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300)


def test_library_is_loaded_and_device_is_sm100(tb):
    from trips_b200 import _lib

    _lib.require_device()
    import ctypes

    sm, cc = ctypes.c_int(), ctypes.c_int()
    l2, mem = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(_lib.lib().tb200_device_info(ctypes.byref(sm), ctypes.byref(l2), ctypes.byref(mem), ctypes.byref(cc)))
    assert cc.value // 10 == 10 and sm.value >= 100


@pytest.mark.parametrize("shape,density", [((700, 500), 0.2), ((300, 400), 0.01), ((64, 64), 0.6), ((1000, 3), 0.5)])
def test_spmv_matches_scipy(tb, shape, density):
    K = tb.kernels
    rng = np.random.default_rng(0)
    A = sp.random(*shape, density=density, random_state=1, format="csr")
    A.data[:] = rng.standard_normal(A.nnz)
    op = tb.CSROperator.from_scipy(A)
    x = rng.standard_normal(shape[1])
    u = rng.standard_normal(shape[0])
    z = rng.standard_normal(shape[0])
    # default order = scipy's summation order: bit-identical to csr_matvec, and to csc_matvec for the transpose,
    # in both device layouts (SELL-32-4 = the default fast path, plain CSR)
    assert op.A_sell is not None and op.A is not None
    for lay in (op, op.with_layout("sell"), op.with_layout("csr")):
        assert np.array_equal(host(lay.apply_dev(dev(x))), A @ x)
        assert np.array_equal(host(lay.adjoint_dev(dev(u))), A.T @ u)
    # the SELL container converts back to exactly the CSR it was made from
    back = op.with_layout("sell").to_scipy()
    assert np.array_equal(back.indptr, A.indptr) and np.array_equal(back.indices, A.indices) and np.array_equal(back.data, A.data)
    tree = op.with_order("tree")
    assert rel(host(tree.apply_dev(dev(x))), A @ x) < 1e-14 and rel(host(tree.adjoint_dev(dev(u))), A.T @ u) < 1e-14
    # fused recurrence epilogue + fused norm, scalar on the device
    import torch

    coef = torch.tensor([0.37], dtype=torch.float64, device="cuda")
    nrm = torch.zeros(2, dtype=torch.float64, device="cuda")
    y = op.apply_dev(dev(x), coef=coef, z=dev(z), norm_out=nrm)
    want = A @ x - 0.37 * z
    assert np.array_equal(host(y), want)
    # the fused norm is the correctly rounded exact sum of squares
    assert host(nrm)[0] == O.exact_dot(want, want) and host(nrm)[1] == np.sqrt(O.exact_dot(want, want))
    yt = tree.apply_dev(dev(x), coef=coef, z=dev(z), norm_out=nrm)
    assert rel(host(yt), want) < 1e-14 and host(nrm)[0] == O.exact_dot(host(yt), host(yt))
    # deterministic: two runs are bitwise identical
    y2 = op.apply_dev(dev(x), coef=coef, z=dev(z), norm_out=nrm)
    assert np.array_equal(host(y), host(y2))
    # numpy in -> numpy out through the operator surface, all input ranks the reference uses
    assert rel(op @ x, A @ x) < 1e-14 and (op @ x.reshape(-1, 1)).shape == (shape[0], 1)
    X = rng.standard_normal((shape[1], 3))
    assert rel(op @ X, A @ X) < 1e-14 and rel(op.T @ u, A.T @ u) < 1e-14
    assert rel(op * x.reshape(-1, 1), (A @ x).reshape(-1, 1)) < 1e-14


def test_spmv_edge_cases_empty_rows_and_ragged(tb):
    rng = np.random.default_rng(3)
    # ragged: empty rows, rows of length 1..300 so every head/body/tail combination of the 4-wide loads occurs
    lens = np.r_[0, 0, np.arange(1, 301), 0, 5, 0]
    indptr = np.r_[0, np.cumsum(lens)]
    n = 977
    indices = np.concatenate([np.sort(rng.choice(n, size=l, replace=False)) for l in lens]).astype(np.int32)
    data = rng.standard_normal(indptr[-1])
    A = sp.csr_matrix((data, indices, indptr), shape=(len(lens), n))
    op = tb.CSROperator.from_scipy(A)
    x = rng.standard_normal(n)
    got = host(op.apply_dev(dev(x)))
    assert np.array_equal(got, A @ x) and got[0] == 0.0 and got[1] == 0.0
    u = rng.standard_normal(len(lens))
    assert np.array_equal(host(op.adjoint_dev(dev(u))), A.T @ u)
    tree = op.with_order("tree")
    assert rel(host(tree.apply_dev(dev(x))), A @ x) < 1e-14 and rel(host(tree.adjoint_dev(dev(u))), A.T @ u) < 1e-14
    # every tile configuration of the sequential kernel gives the same bits
    from trips_b200 import _lib

    # (tile kernels 0-3, direct register-streaming kernels 4-6, forced gather mappings 8*mode)
    csr_only, sell_only = op.with_layout("csr"), op.with_layout("sell")
    for variant in (1, 2, 3, 4, 5, 6, 8 + 1, 16 + 1, 256, 512, 768, 0):
        _lib.check(_lib.lib().tb200_spmv_set_variant(variant))
        for lay in (csr_only, sell_only):
            assert np.array_equal(host(lay.apply_dev(dev(x))), A @ x), variant
            assert np.array_equal(host(lay.adjoint_dev(dev(u))), A.T @ u), variant
    # fp32-storage / fp64-accumulate variant: exact on the rounded values
    op32 = op.with_f32_storage()
    A32 = A.copy()
    A32.data = A.data.astype(np.float32).astype(np.float64)
    assert np.array_equal(host(op32.apply_dev(dev(x))), A32 @ x)
    assert rel(host(op32.with_order("tree").apply_dev(dev(x))), A32 @ x) < 1e-14
    # all-empty matrix
    E = tb.CSROperator.from_scipy(sp.csr_matrix((5, 7)))
    assert np.array_equal(host(E.apply_dev(dev(np.ones(7)))), np.zeros(5))


def test_vector_kernels_round_like_numpy(tb):
    K = tb.kernels
    rng = np.random.default_rng(1)
    n = 100_003
    x, y, w = rng.standard_normal(n), rng.standard_normal(n), rng.uniform(0.1, 3, n)
    a = 0.7310585786300049
    assert np.array_equal(host(K.vec_div(dev(x), a)), x / a)
    assert np.array_equal(host(K.vec_axpy(a, dev(x), dev(y))), y + a * x)
    assert np.array_equal(host(K.vec_axpy(a, dev(x), dev(y), sign=-1.0)), y - a * x)
    assert np.array_equal(host(K.vec_mul(dev(x), dev(w))), x * w)
    assert np.array_equal(host(K.vec_sub(dev(x), dev(y))), x - y)
    assert np.array_equal(host(K.vec_wsub(dev(w), dev(x), dev(y))), w * (x - y))
    assert np.array_equal(host(K.vec_add(dev(x), dev(y))), x + y)
    for expo in (-0.5, -0.25, 0.0, -0.3):
        got = host(K.irls_weights(dev(x), 0.1, expo))
        assert np.allclose(got, (x ** 2 + 0.1 ** 2) ** expo, rtol=1e-15, atol=0)
    assert np.array_equal(host(K.irls_weights(dev(x), 0.1, 0.0)), np.ones(n))
    # reductions return the correctly rounded exact sums (double-double accumulation), whatever the grid
    nrm = host(K.vec_norm2(dev(x)))
    assert nrm[0] == O.exact_dot(x, x) and nrm[1] == np.sqrt(O.exact_dot(x, x))
    assert host(K.vec_dot(dev(x), dev(y)))[0] == O.exact_dot(x, y)
    assert host(K.vec_diffnorm2(dev(x), dev(y)))[0] == O.exact_dot(x - y, x - y)
    for nn in (1, 31, 257, 5000):
        assert host(K.vec_norm2(dev(x[:nn])))[0] == O.exact_dot(x[:nn], x[:nn])
    # (double-double carries ~106 bits: the guarantee holds unless cancellation exceeds ~1e15, which norms never do)
    # fused norm of the axpy result; device-resident scalar
    import torch

    pair = torch.zeros(2, dtype=torch.float64, device="cuda")
    ad = torch.tensor([a], dtype=torch.float64, device="cuda")
    out = K.vec_axpy(ad, dev(x), dev(y), norm_out=pair)
    assert np.array_equal(host(out), y + a * x)
    assert abs(host(pair)[1] - np.linalg.norm(y + a * x)) < 1e-14 * np.linalg.norm(y + a * x)
    # empty vectors
    assert host(K.vec_norm2(dev(np.zeros(0))))[0] == 0.0


@pytest.mark.parametrize("n,k", [(50_001, 1), (50_001, 7), (200_000, 19), (4096, 40)])
def test_basis_kernels(tb, n, k):
    K = tb.kernels
    rng = np.random.default_rng(2)
    V = rng.standard_normal((n, k))
    w = rng.standard_normal(n)
    basis = K.Basis(n, k + 2, "cuda")
    for j in range(k):
        basis.next_col().copy_(dev(V[:, j]))
        basis.push()
    h = host(K.basis_dots(basis, k, dev(w)))[:k]
    assert np.allclose(h, V.T @ w, rtol=0, atol=1e-12 * np.linalg.norm(w) * np.sqrt(n))
    hh = rng.standard_normal(k)
    import torch

    pair = torch.zeros(2, dtype=torch.float64, device="cuda")
    out = host(K.basis_combine(basis, k, dev(hh), w=dev(w), sign=-1.0, norm_out=pair))
    want = w - V @ hh
    assert rel(out, want) < 1e-14 and abs(host(pair)[1] - np.linalg.norm(want)) < 1e-13 * np.linalg.norm(want)
    assert rel(host(K.basis_combine(basis, k, dev(hh))), V @ hh) < 1e-14
    assert np.array_equal(basis.to_numpy(), V)
    # growth keeps the columns
    for _ in range(4):
        basis.next_col().zero_()
        basis.push()
    assert np.array_equal(basis.to_numpy()[:, :k], V)


@pytest.mark.parametrize("n,k", [(100_000, 11), (65_538, 4), (4098, 21), (1000, 1)])
def test_basis_kernels_two_rows_per_access_equal_one_row(tb, n, k):
    """basis_combine accumulates every row in the same order whether a thread takes one row or two per access: same
    bits (incl. the fused norm); basis_dots only regroups its partial sums."""
    K = tb.kernels
    rng = np.random.default_rng(8)
    V = rng.standard_normal((n, k))
    basis = K.Basis(n, k, "cuda")
    for j in range(k):
        basis.next_col().copy_(dev(V[:, j]))
        basis.push()
    w, h = dev(rng.standard_normal(n)), dev(rng.standard_normal(k))
    got = []
    try:
        for v in (1, 0):
            tb._lib.lib().tb200_basis_set_vec2(v)
            pair = K.new_pair("cuda")
            o1 = host(K.basis_combine(basis, k, h, w=w, sign=-1.0, norm_out=pair))
            o2 = host(K.basis_combine(basis, k, h))
            got.append((o1, o2, host(pair), host(K.basis_dots(basis, k, w))))
    finally:
        tb._lib.lib().tb200_basis_set_vec2(1)
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])
    assert np.array_equal(got[0][2], got[1][2])
    assert rel(got[0][3], got[1][3]) < 1e-13 and rel(got[0][3], V.T @ host(w)) < 1e-13
    assert rel(got[0][0], host(w) - V @ host(h)) < 1e-13


@pytest.mark.parametrize("m,k,weighted", [(5000, 3, False), (20_001, 10, True), (3000, 37, True)])
def test_weighted_gram_and_factor_match_householder(tb, m, k, weighted):
    K = tb.kernels
    rng = np.random.default_rng(4)
    # ill-conditioned basis: kappa ~ 1e6, so a plain-double Gram/Cholesky would lose ~12 digits
    Uo = np.linalg.qr(rng.standard_normal((m, k)))[0]
    Vo = np.linalg.qr(rng.standard_normal((k, k)))[0]
    B = Uo @ np.diag(np.logspace(0, -6, k)) @ Vo.T
    b = rng.standard_normal(m)
    w = rng.uniform(0.5, 10, m) if weighted else None
    basis = K.Basis(m, k, "cuda")
    for j in range(k):
        basis.next_col().copy_(dev(B[:, j]))
        basis.push()
    bd = dev(b)
    Ghi, Glo = K.weighted_gram(basis, k, dev(w) if weighted else None, extras=(bd, bd), extra_weighted=(0, 1))
    R, C, res2 = K.gram_factor(Ghi, Glo, k)
    M = B * w[:, None] if weighted else B
    Q, Rr = np.linalg.qr(M)
    sg = np.sign(np.diag(Rr))
    Q, Rr = Q * sg[None, :], Rr * sg[:, None]
    assert np.allclose(R, Rr, rtol=0, atol=1e-9 * np.abs(Rr).max())
    for i in range(k):
        assert np.linalg.norm(R[i] - Rr[i]) <= 1e-8 * np.linalg.norm(Rr[i])
    wb = b * w if weighted else b
    assert np.allclose(C[:, 0], Q.T @ b, atol=1e-8 * np.linalg.norm(b))
    assert np.allclose(C[:, 1], Q.T @ wb, atol=1e-8 * np.linalg.norm(wb))
    assert abs(res2[1] - np.linalg.norm(wb - Q @ (Q.T @ wb)) ** 2) < 1e-8 * (wb @ wb)
    # the quantity the solvers consume: y = argmin ||R y - c||^2 + lam ||y||^2 agrees with the Householder route
    lam = 1e-3
    y1 = np.linalg.lstsq(np.vstack((R, np.sqrt(lam) * np.eye(k))), np.r_[C[:, 0], np.zeros(k)], rcond=None)[0]
    y2 = np.linalg.lstsq(np.vstack((Rr, np.sqrt(lam) * np.eye(k))), np.r_[Q.T @ b, np.zeros(k)], rcond=None)[0]
    assert rel(y1, y2) < 1e-9


@pytest.mark.parametrize("m,k,weighted", [(100_002, 7, True), (100_002, 29, False), (77_778, 55, True), (100_003, 12, True),
                                          (513, 5, True), (40, 3, False)])
def test_weighted_gram_bulk_copy_pipeline_equals_plain_staging(tb, m, k, weighted):
    """The double-buffered cp.async.bulk staging and the ordinary-load staging feed the same accumulation: same bits; and
    both are the Gram matrix to double-double accuracy (checked against a long-double product).  Odd m (odd leading
    dimension) cannot be bulk-copied and takes the plain path by itself; m % tile != 0 exercises the partial tile."""
    K = tb.kernels
    rng = np.random.default_rng(11)
    B = rng.standard_normal((m, k)) * np.logspace(0, -5, k)[None, :]
    b = rng.standard_normal(m)
    w = rng.uniform(1e-3, 30, m) if weighted else None
    basis = K.Basis(m, k + 2, "cuda")
    for j in range(k):
        basis.next_col().copy_(dev(B[:, j]))
        basis.push()
    bd, wd = dev(b), (dev(w) if weighted else None)
    out = []
    try:
        for block in (2, 4):
            tb._lib.lib().tb200_gram_set_block(block)
            for bulk in (1, 0):
                tb._lib.lib().tb200_gram_set_bulk(bulk)
                out.append(K.weighted_gram(basis, k, wd, extras=(bd, bd), extra_weighted=(0, 1)))
    finally:
        tb._lib.lib().tb200_gram_set_bulk(1)
        tb._lib.lib().tb200_gram_set_block(0)
    for a in (0, 2):  # same block shape: same accumulation order whatever the staging
        assert np.array_equal(out[a][0], out[a + 1][0]) and np.array_equal(out[a][1], out[a + 1][1])
    d = (out[0][0].astype(np.longdouble) + out[0][1]) - (out[2][0].astype(np.longdouble) + out[2][1])
    assert np.max(np.abs(d.astype(np.float64)) / np.sqrt(np.outer(np.diag(out[0][0]), np.diag(out[0][0])))) < 1e-17
    M = np.column_stack((B * w[:, None] if weighted else B, b, b * w if weighted else b)).astype(np.longdouble)
    G = M.T @ M
    scale = np.sqrt(np.outer(np.diag(G), np.diag(G))).astype(np.float64)
    got = out[0][0].astype(np.longdouble) + out[0][1].astype(np.longdouble)
    assert np.max(np.abs((got - G).astype(np.float64)) / scale) < 1e-15


@pytest.mark.parametrize("m", [60_000, 4099])
def test_incremental_gram_panel_equals_the_full_pass(tb, m):
    """An append-only, unweighted basis: the panel pass (new columns of G only) merged into the cached matrix equals the
    full double-double pass to double-double accuracy at every size, including panels that straddle 4-column blocks,
    several new columns at once, and a stale cache (another extra vector) that must fall back to the full pass."""
    K = tb.kernels
    rng = np.random.default_rng(5)
    kmax = 23
    B = rng.standard_normal((m, kmax)) * np.logspace(0, -4, kmax)[None, :]
    b = dev(rng.standard_normal(m))
    basis = K.Basis(m, kmax, "cuda")
    inc = K.IncrementalGram()
    k = 0
    for grow in (3, 1, 1, 1, 2, 1, 1, 5, 1, 1, 1, 1, 1, 1, 1, 1):
        for _ in range(grow):
            basis.next_col().copy_(dev(B[:, k]))
            basis.push()
            k += 1
        hi, lo = inc.update(basis, k, extras=(b,), extra_weighted=(0,))
        fhi, flo = K.weighted_gram(basis, k, None, extras=(b,), extra_weighted=(0,))
        got = hi.astype(np.longdouble) + lo.astype(np.longdouble)
        want = fhi.astype(np.longdouble) + flo.astype(np.longdouble)
        scale = np.sqrt(np.outer(np.diag(fhi), np.diag(fhi)))
        assert np.max(np.abs((got - want).astype(np.float64)) / scale) < 1e-17, k   # (long double carries 64 bits)
        assert np.array_equal(hi, hi.T) and np.array_equal(lo, lo.T)
        R1, C1, r1 = K.gram_factor(hi, lo, k)
        R2, C2, r2 = K.gram_factor(fhi, flo, k)
        assert np.allclose(R1, R2, rtol=1e-13, atol=0) and np.allclose(C1, C2, rtol=1e-12, atol=1e-300)
    assert k == kmax
    b2 = dev(rng.standard_normal(m))
    hi, lo = inc.update(basis, k, extras=(b2,), extra_weighted=(0,))  # different extra: cache key changes -> full pass
    fhi, flo = K.weighted_gram(basis, k, None, extras=(b2,), extra_weighted=(0,))
    assert np.array_equal(hi, fhi) and np.array_equal(lo, flo)


def test_gram_route_falls_back_to_householder_when_not_positive_definite(tb):
    """A basis vector in the null space of L (a zero column of LV) makes the Gram matrix singular: the double-double
    Cholesky refuses it and GKS / MMGKS factor that iteration the reference's way (Householder QR on the host) instead
    of stopping - the reference's la.qr walks through the same situation (GKS.py:54-58, MMGKS.py:94-95)."""
    import sys

    K = tb.kernels
    core = sys.modules[sys.modules[tb.GKS.__module__].GKSBases.__module__]
    rng = np.random.default_rng(21)
    m, k = 5000, 5
    B = rng.standard_normal((m, k))
    B[:, 2] = 0.0
    basis = K.Basis(m, k, "cuda")
    for j in range(k):
        basis.next_col().copy_(dev(B[:, j]))
        basis.push()
    b, w = rng.standard_normal(m), rng.uniform(0.5, 2.0, m)
    with pytest.warns(RuntimeWarning, match="Householder"):
        R, C, r2 = core._factor(basis, k, dev(w), (dev(b), dev(b)), (0, 1), None, None)
    Rw, Cw, rw = core.householder_factor(B * w[:, None], np.stack((b, b * w), axis=1))
    assert R.shape == (k, k) and C.shape == (k, 2) and R[2, 2] == 0.0
    assert np.allclose(R, Rw, rtol=0, atol=1e-12 * np.abs(Rw).max()) and np.allclose(C, Cw, atol=1e-10) and np.allclose(r2, rw)
    with pytest.warns(RuntimeWarning, match="Householder"):
        R, C, r2 = core._factor(basis, k, None, (dev(b),), (0,), K.IncrementalGram(), None)
    Ru, Cu, ru = core.householder_factor(B, b[:, None])
    assert np.allclose(R, Ru, rtol=0, atol=1e-12 * np.abs(Ru).max()) and np.allclose(C, Cu, atol=1e-10) and np.allclose(r2, ru)
    # a healthy basis does not take the fallback
    B[:, 2] = rng.standard_normal(m)
    basis.col(2).copy_(dev(B[:, 2]))
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("error", RuntimeWarning)
        R, C, r2 = core._factor(basis, k, dev(w), (dev(b), dev(b)), (0, 1), None, None)
    Rw, Cw, rw = core.householder_factor(B * w[:, None], np.stack((b, b * w), axis=1))
    assert np.allclose(R, Rw, rtol=1e-11, atol=1e-12 * np.abs(Rw).max()) and np.allclose(C, Cw, atol=1e-10)


@pytest.mark.parametrize("nx,views", [(24, 16), (64, 90), (33, 7)])
def test_ct_builder_is_bit_identical_to_the_numpy_statement(tb, nx, views):
    op = tb.ParallelBeamCT(nx, views)
    A = op.to_scipy()
    AT = op.transpose_to_scipy()
    A0 = O.ct_matrix(nx, O.ct_angles(views))
    assert A.shape == A0.shape and A.nnz == A0.nnz
    assert np.array_equal(A.indptr, A0.indptr) and np.array_equal(A.indices, A0.indices)
    assert np.array_equal(A.data, A0.data)
    # the stored transpose is the exact transpose (same values, same pattern, sorted indices)
    T0 = A0.T.tocsr()
    T0.sort_indices()
    assert np.array_equal(AT.indptr, T0.indptr) and np.array_equal(AT.indices, T0.indices)
    assert np.array_equal(AT.data, T0.data)
    # the natively built SELL-32-4 matrices hold exactly the same rows (and little padding)
    sell = tb.ParallelBeamCT(nx, views, layout="sell")
    assert sell.A is None and sell.AT is None
    As, ATs = sell.to_scipy(), sell.transpose_to_scipy()
    assert np.array_equal(As.indptr, A0.indptr) and np.array_equal(As.indices, A0.indices) and np.array_equal(As.data, A0.data)
    assert np.array_equal(ATs.indptr, T0.indptr) and np.array_equal(ATs.indices, T0.indices) and np.array_equal(ATs.data, T0.data)
    assert sell.AT_sell.stored <= 1.15 * A0.nnz + 4096
    # adjoint identity through the kernels
    rng = np.random.default_rng(0)
    x, u = rng.standard_normal(A.shape[1]), rng.standard_normal(A.shape[0])
    assert np.array_equal(host(sell.apply_dev(dev(x))), A0 @ x) and np.array_equal(host(sell.adjoint_dev(dev(u))), A0.T @ u)
    lhs = u @ host(op.apply_dev(dev(x)))
    rhs = host(op.adjoint_dev(dev(u))) @ x
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)


AWKWARD = sorted(set(t + d for t in (0.0, np.pi / 2, np.pi / 4, 3 * np.pi / 4, np.arctan(3.0), np.pi - np.arctan(3.0),
                                      np.arctan(1.0 / 3.0), np.pi / 3, np.pi / 6)
                     for d in (0.0, 1e-7, -1e-7, 3e-6, -3e-6, 1e-12) if 0 <= t + d < np.pi))


@pytest.mark.parametrize("forward", ["rays", "index"])
@pytest.mark.parametrize("nx,ny,views,n_det,angles", [
    (24, None, 16, None, None), (64, None, 90, None, None), (33, 20, 7, None, None), (40, 56, 9, 70, None),
    (48, None, 5, None, [0.0, np.pi / 2, np.pi / 4, 3 * np.pi / 4, 3.0]), (37, None, 11, 30, None), (128, None, 13, None, None),
    (20, None, len(AWKWARD), 29, AWKWARD), (130, 70, 40, None, None), (96, 200, 24, 333, None),
    (8, 7000, 5, 40, None), (8, 13000, 5, 40, None)])  # tall images: > 48 KB / no shared-memory row table
def test_matrix_free_projectors_are_bit_identical_to_the_stored_matrix(tb, nx, ny, views, n_det, angles, forward):
    """layout='implicit': the ray-driven forward projector enumerates each ray's pixels (forward='rays'; 'index' is round
    1's, which streams A's column indices), back-projection is pixel-driven.  Both must give the SAME BITS as scipy on
    the stored matrix (same entries, same summation order), including the fused recurrence / norm epilogue.  Covers
    non-square images, detectors narrower / wider than the image (clipped footprints), axis-aligned, 30 / 45 / 60-degree
    angles (degenerate trapezoids, rays through pixel corners) and angles a hair off the kernel's class boundaries."""
    kw = dict(ny=ny, n_det=n_det, angles=None if angles is None else np.array(angles))
    mf = tb.ParallelBeamCT(nx, views, layout="implicit", forward=forward, **kw)
    A0 = tb.ParallelBeamCT(nx, views, layout="csr", **kw).to_scipy()
    assert mf.shape == A0.shape and mf.nnz == A0.nnz
    if forward == "rays":
        assert mf.projector.nbytes <= 64 * len(mf.theta) + 16  # a per-angle table and nothing else
    elif (ny or nx) < 1000:  # (not the tall, nearly empty test images: their per-row tables dominate)
        assert mf.projector.nbytes < 0.3 * 24 * A0.nnz + 65536 + 16 * A0.shape[0] + 8 * A0.shape[1]  # column indices + per-row tables
    rng = np.random.default_rng(3)
    for trial in range(2):
        x, u = rng.standard_normal(A0.shape[1]), rng.standard_normal(A0.shape[0])
        assert np.array_equal(host(mf.apply_dev(dev(x))), A0 @ x)
        assert np.array_equal(host(mf.adjoint_dev(dev(u))), A0.T @ u)
    # fused epilogues: y = A x - coef z with ||y||, coefficient on the device
    import torch

    z_m, z_n = rng.standard_normal(A0.shape[0]), rng.standard_normal(A0.shape[1])
    coef = torch.tensor([0.37], dtype=torch.float64, device="cuda")
    pair = torch.zeros(2, dtype=torch.float64, device="cuda")
    y = host(mf.apply_dev(dev(x), coef=coef, z=dev(z_m), norm_out=pair))
    want = A0 @ x - 0.37 * z_m
    assert np.array_equal(y, want) and float(pair[1]) == float(np.sqrt(O.exact_dot(want, want)))
    v = host(mf.adjoint_dev(dev(u), coef=0.37, z=dev(z_n), norm_out=pair))
    want = A0.T @ u - 0.37 * z_n
    assert np.array_equal(v, want) and float(pair[1]) == float(np.sqrt(O.exact_dot(want, want)))
    assert (mf.to_scipy() != A0).nnz == 0  # explicit() materialises the same matrix on demand


@pytest.mark.parametrize("nx,ny,views,n_det,angles", [
    (24, None, 16, None, None), (64, None, 90, None, None), (33, 20, 7, None, None), (40, 56, 9, 70, None),
    (48, None, 5, None, [0.0, np.pi / 2, np.pi / 4, 3 * np.pi / 4, 3.0]), (128, None, 13, None, None),
    (20, None, len(AWKWARD), 29, AWKWARD), (130, 70, 40, None, None), (96, 200, 24, 333, None)])
def test_row_aligned_stored_sell_layout_is_bit_identical_to_the_plain_one(tb, nx, ny, views, n_det, angles, monkeypatch):
    """ParallelBeamCT(layout='sell') stores A in the row-aligned layout of the index-only projector with the values next
    to the indices (leading padding per ray, shallow rays addressing the transposed image): the same matrix (to_scipy),
    the same bits of A x with and without the fused epilogue, in fp32 storage too, and the same Golub-Kahan factors as
    the plain SELL layout (TB200_SELL_ALIGNED=0)."""
    import torch

    kw = dict(ny=ny, n_det=n_det, angles=None if angles is None else np.array(angles))
    op = tb.ParallelBeamCT(nx, views, layout="sell", **kw)
    assert isinstance(op.A_sell, tb.kernels.CtSellDevice)
    monkeypatch.setenv("TB200_SELL_ALIGNED", "0")
    plain = tb.ParallelBeamCT(nx, views, layout="sell", **kw)
    monkeypatch.delenv("TB200_SELL_ALIGNED")
    assert not isinstance(plain.A_sell, tb.kernels.CtSellDevice)
    A0 = tb.ParallelBeamCT(nx, views, layout="csr", **kw).to_scipy()
    As = op.to_scipy()
    assert np.array_equal(As.indptr, A0.indptr) and np.array_equal(As.indices, A0.indices) and np.array_equal(As.data, A0.data)
    rng = np.random.default_rng(9)
    x, z = rng.standard_normal(A0.shape[1]), rng.standard_normal(A0.shape[0])
    assert np.array_equal(host(op.apply_dev(dev(x))), A0 @ x)
    coef = torch.tensor([0.37], dtype=torch.float64, device="cuda")
    pair = torch.zeros(2, dtype=torch.float64, device="cuda")
    y = host(op.apply_dev(dev(x), coef=coef, z=dev(z), norm_out=pair))
    want = A0 @ x - 0.37 * z
    assert np.array_equal(y, want) and float(pair[1]) == float(np.sqrt(O.exact_dot(want, want)))
    assert np.array_equal(host(op.with_f32_storage().apply_dev(dev(x))), host(plain.with_f32_storage().apply_dev(dev(x))))
    b = rng.standard_normal(A0.shape[0])
    got, ref = tb.golub_kahan_device(op, b, 6), tb.golub_kahan_device(plain, b, 6)
    assert np.array_equal(got.B_host(), ref.B_host())
    assert np.array_equal(got.U.to_numpy(), ref.U.to_numpy()) and np.array_equal(got.V.to_numpy(), ref.V.to_numpy())


@pytest.mark.parametrize("forward", ["rays", "index"])
def test_matrix_free_golub_kahan_equals_stored_matrix_golub_kahan(tb, forward):
    nx, views, steps = 48, 36, 12
    b = np.random.default_rng(1).standard_normal(views * O.ct_num_detectors(nx))
    ref = tb.golub_kahan_device(tb.ParallelBeamCT(nx, views, layout="sell"), b, steps)
    got = tb.golub_kahan_device(tb.ParallelBeamCT(nx, views, layout="implicit", forward=forward), b, steps)
    assert np.array_equal(got.B_host(), ref.B_host())
    assert np.array_equal(got.U.to_numpy(), ref.U.to_numpy()) and np.array_equal(got.V.to_numpy(), ref.V.to_numpy())


@pytest.mark.parametrize("nx,ny,views,n_det,geom", [(24, None, 16, None, None), (64, None, 45, None, None),
                                                    (33, 20, 7, None, None), (40, 40, 9, 70, (90.0, 25.0, 1.1)),
                                                    (48, None, 5, None, (60.0, 0.0, 1.0))])
def test_fan_beam_builder_is_bit_identical_to_the_numpy_statement(tb, nx, ny, views, n_det, geom):
    """Flat-detector fan beam (the reference's ASTRA geometry, Tomography.py:57-67): A in CSR and SELL-32-4 and the
    stored transpose against the oracle's NumPy statement of the same per-ray arithmetic; SpMV = scipy bit for bit."""
    kw = {} if geom is None else dict(source_origin=geom[0], detector_origin=geom[1], detector_pixel_size=geom[2])
    op = tb.FanBeamCT(nx, views, ny=ny, n_det=n_det, layout="both", **kw)
    fan = O.fan_geometry(nx) if geom is None else geom
    assert op.fan == tuple(fan)
    A0 = O.ct_matrix(nx, O.ct_angles(views), ny=ny, n_det=n_det, fan=fan)
    A, AT = op.to_scipy(), op.transpose_to_scipy()
    assert A.shape == A0.shape and A.nnz == A0.nnz
    assert np.array_equal(A.indptr, A0.indptr) and np.array_equal(A.indices, A0.indices) and np.array_equal(A.data, A0.data)
    T0 = A0.T.tocsr()
    T0.sort_indices()
    assert np.array_equal(AT.indptr, T0.indptr) and np.array_equal(AT.indices, T0.indices) and np.array_equal(AT.data, T0.data)
    rng = np.random.default_rng(0)
    x, u = rng.standard_normal(A.shape[1]), rng.standard_normal(A.shape[0])
    assert np.array_equal(host(op.apply_dev(dev(x))), A0 @ x) and np.array_equal(host(op.adjoint_dev(dev(u))), A0.T @ u)
    # physics: every ray that crosses the image centrally has a path length close to the image width / |cos| bound
    assert 0.9 * min(nx, ny if ny else nx) < (A0 @ np.ones(A.shape[1])).max() <= np.hypot(nx, ny if ny else nx) + 1e-9


def test_ct_builder_angle_subset_and_block_diagonal(tb):
    nx, views = 32, 12
    full = tb.ParallelBeamCT(nx, views).to_scipy()
    n_det = O.ct_num_detectors(nx)
    sub = tb.ParallelBeamCT(nx, views, angle_subset=np.arange(1, views, 2)).to_scipy()
    rows = np.concatenate([np.arange(a * n_det, (a + 1) * n_det) for a in range(1, views, 2)])
    assert (sub != full[rows]).nnz == 0
    th = O.ct_angles(views)
    frames = [th[0:4], th[4:8] + 0.01, th[8:12] + 0.02]
    bd = tb.BlockDiagCT(nx, frames)
    blocks = [O.ct_matrix(nx, f) for f in frames]
    want = sp.block_diag(blocks, format="csr")
    got = bd.to_scipy()
    assert got.shape == want.shape and (got != want).nnz == 0
    assert (bd.transpose_to_scipy() != want.T.tocsr()).nnz == 0


@pytest.mark.parametrize("shape,psf_dim,spread", [((32, 32), (7, 7), (2, 2)), ((50, 37), (9, 5), (3, 1.5)),
                                                   ((40, 40), (4, 6), (1.0, 2.0)), ((5, 6), (9, 9), (3, 3))])
def test_psf_blur_is_bit_identical_to_ndimage(tb, shape, psf_dim, spread):
    rng = np.random.default_rng(0)
    nx, ny = shape
    PSF = O.gauss_psf(psf_dim, spread)
    PSF = PSF + 0.01 * rng.uniform(size=PSF.shape)  # not symmetric: tap ORDER and flips must be right
    op = tb.PSFBlur2D(PSF, nx, ny)
    X = rng.standard_normal((nx, ny))
    got = host(op.apply_dev(dev(X.ravel()))).reshape(nx, ny)
    assert np.array_equal(got, convolve(X, PSF, mode="reflect"))
    got_t = host(op.adjoint_dev(dev(X.ravel()))).reshape(nx, ny)
    assert np.array_equal(got_t, convolve(X, np.flipud(np.fliplr(PSF)), mode="reflect"))
    opc = tb.PSFBlur2D(PSF, nx, ny, mode="constant")
    assert np.array_equal(host(opc.apply_dev(dev(X.ravel()))).reshape(nx, ny), convolve(X, PSF, mode="constant"))


def test_psf_blur_golden_and_adjointness(tb, golden_dir):
    g = np.load(f"{golden_dir}/deblur32.npz")
    n = int(g["n"])
    op = tb.PSFBlur2D(g["PSF"], n, n)
    assert np.array_equal(op @ g["probe"], g["fwd_probe"]) and np.array_equal(op.T @ g["probe"], g["adj_probe"])
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(n * n), rng.standard_normal(n * n)
    assert abs(y @ (op @ x) - (op.T @ y) @ x) < 1e-12  # symmetric Gaussian PSF: the reference's A^T is the true adjoint


@pytest.mark.parametrize("nx,nt", [(8, 1), (31, 1), (16, 5), (7, 2)])
def test_difference_operators_bit_identical_to_the_sparse_matrices(tb, nx, nt):
    rng = np.random.default_rng(0)
    L = O.first_derivative_2d(nx, nx) if nt == 1 else O.spacetime_derivative(nx, nx, nt)
    op = tb.SpaceTimeDerivative(nx, nx, nt)
    assert op.shape == L.shape
    x = rng.standard_normal(L.shape[1])
    r = rng.standard_normal(L.shape[0])
    assert np.array_equal(op @ x, L @ x)
    assert np.array_equal(op.T @ r, L.T @ r)
    # fused IRLS weights and weighted adjoint
    import torch

    wout = torch.empty(L.shape[0], dtype=torch.float64, device="cuda")
    u = op.apply_dev(dev(x), wout=wout, eps=0.1, expo=-0.5)
    assert np.array_equal(host(u), L @ x)
    assert np.allclose(host(wout), ((L @ x) ** 2 + 0.01) ** (-0.5), rtol=1e-15, atol=0)
    w = rng.uniform(0.5, 2, L.shape[0])
    assert np.array_equal(host(op.adjoint_dev(dev(r), w=dev(w))), L.T @ (w * r))
    if nt == 1:
        L1 = O.first_derivative_1d(nx)
        o1 = tb.FirstDerivative1D(nx)
        z = rng.standard_normal(nx)
        assert np.array_equal(o1 @ z, L1 @ z) and np.array_equal(o1.T @ z[:-1], L1.T @ z[:-1])


@pytest.mark.parametrize("nx,ny", [(6, 6), (17, 17), (9, 14)])
def test_centred_gradient_and_iso_weights_bit_identical(tb, nx, ny):
    import torch

    rng = np.random.default_rng(0)
    if nx == ny:
        L = O.centered_derivative_2d(nx, ny)
    else:  # general rectangle: [I_nx (x) D(ny) ; D(nx) (x) I_ny]
        def D(n):
            M = sp.lil_matrix((n, n))
            for i in range(1, n - 1):
                M[i, i + 1], M[i, i - 1] = 0.5, -0.5
            return M.tocsr()
        L = sp.vstack((sp.kron(sp.identity(nx), D(ny)), sp.kron(D(nx), sp.identity(ny)))).tocsr()
    op = tb.CenteredDerivative2D(nx, ny)
    assert op.shape == L.shape
    x, r = rng.standard_normal(nx * ny), rng.standard_normal(2 * nx * ny)
    assert np.array_equal(op @ x, L @ x) and np.array_equal(op.T @ r, L.T @ r)
    w = rng.uniform(0.5, 2, 2 * nx * ny)
    assert np.array_equal(host(op.adjoint_dev(dev(r), w=dev(w))), L.T @ (w * r))
    u = L @ x
    N = nx * ny
    want = (u[:N] ** 2 + u[N:] ** 2 + 0.1 ** 2) ** ((1 - 2) / 4)
    got = host(op.iso_weights(dev(x), 0.1, (1 - 2) / 4))
    assert np.allclose(got[:N], want, rtol=1e-15, atol=0) and np.array_equal(got[:N], got[N:])


def test_difference_operator_frame_sharding_with_halos(tb):
    """Two 'ranks' each owning half of the frames reproduce the global operator given one-frame halos."""
    import torch

    nx, nt = 6, 6
    N = nx * nx
    rng = np.random.default_rng(0)
    L = O.spacetime_derivative(nx, nx, nt)
    x = rng.standard_normal(nt * N)
    u_glob = L @ x
    p2 = 2 * nx * (nx - 1)
    lo = tb.SpaceTimeDerivative(nx, nx, 3, has_next=True)
    hi = tb.SpaceTimeDerivative(nx, nx, 3)
    x0, x1 = dev(x[:3 * N]), dev(x[3 * N:])
    u0 = host(lo.apply_dev(x0, x_next=x1[:N].contiguous()))
    u1 = host(hi.apply_dev(x1))
    spatial = np.r_[u0[:3 * p2], u1[:3 * p2]]
    temporal = np.r_[u0[3 * p2:], u1[3 * p2:]]
    assert np.array_equal(spatial, u_glob[:nt * p2]) and np.array_equal(temporal, u_glob[nt * p2:])
    r = rng.standard_normal(L.shape[0])
    g_glob = L.T @ r
    rs, rt = r[:nt * p2], r[nt * p2:].reshape(nt - 1, N)
    r0 = dev(np.r_[rs[:3 * p2], rt[:3].ravel()])
    r1 = dev(np.r_[rs[3 * p2:], rt[3:].ravel()])
    g0 = host(lo.adjoint_dev(r0))
    g1 = host(hi.adjoint_dev(r1, rt_prev=dev(rt[2])))
    assert np.array_equal(np.r_[g0, g1], g_glob)


@pytest.mark.parametrize("nx,ny,views,n_det,geom", [(24, None, 16, None, None), (64, None, 45, None, None),
                                                    (33, 20, 7, None, None), (40, 40, 9, 70, (90.0, 25.0, 1.1)),
                                                    (48, None, 5, None, (60.0, 0.0, 1.0)), (256, None, 24, None, None)])
def test_fan_beam_matrix_free_forward_is_bit_identical_to_the_stored_matrix(tb, nx, ny, views, n_det, geom):
    """FanBeamCT(layout='implicit'): the ray-driven forward projector with per-ray geometry (the reference's own ASTRA
    geometry, Tomography.py:57-67) gives the bits of the stored product; A^T stays a stored SpMV; Golub-Kahan on it equals
    Golub-Kahan on the stored pair."""
    import torch

    kw = {} if geom is None else dict(source_origin=geom[0], detector_origin=geom[1], detector_pixel_size=geom[2])
    st = tb.FanBeamCT(nx, views, ny=ny, n_det=n_det, layout="sell", **kw)
    mf = tb.FanBeamCT(nx, views, ny=ny, n_det=n_det, layout="implicit", **kw)
    assert mf.shape == st.shape and mf.nnz == st.nnz and mf.A_sell is None and mf.A is None
    rng = np.random.default_rng(9)
    x, u, z = rng.standard_normal(st.shape[1]), rng.standard_normal(st.shape[0]), rng.standard_normal(st.shape[0])
    assert torch.equal(mf.apply_dev(dev(x)), st.apply_dev(dev(x)))
    assert torch.equal(mf.adjoint_dev(dev(u)), st.adjoint_dev(dev(u)))
    pair_a = torch.zeros(2, dtype=torch.float64, device="cuda")
    pair_b = torch.zeros(2, dtype=torch.float64, device="cuda")
    ya = mf.apply_dev(dev(x), coef=0.25, z=dev(z), norm_out=pair_a)
    yb = st.apply_dev(dev(x), coef=0.25, z=dev(z), norm_out=pair_b)
    assert torch.equal(ya, yb) and torch.equal(pair_a, pair_b)
    b = rng.standard_normal(st.shape[0])
    ga, gb = tb.golub_kahan_device(mf, b, 6), tb.golub_kahan_device(st, b, 6)
    assert np.array_equal(ga.B_host(), gb.B_host()) and np.array_equal(ga.V.to_numpy(), gb.V.to_numpy())
    assert (mf.to_scipy() != st.to_scipy()).nnz == 0
