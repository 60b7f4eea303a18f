"""The C-ABI library loads and exports exactly what include/tripsb200.h declares (no compute calls: CPU only)."""
import ctypes
import os
import re

import numpy as np

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tripsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from trips_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "build libtripsb200.so first (__graft_entry__.build())"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in tripsb200.h but not exported"


def test_ctypes_table_matches_header():
    from trips_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()
    assert _lib.lib().tb200_version() >= 100


def test_workspace_queries_and_error_channel():
    from trips_b200 import _lib

    L = _lib.lib()
    assert L.tb200_spmv_workspace_len(1000) >= 1000 // 16
    assert L.tb200_reduce_workspace_len() > 0
    assert L.tb200_basis_workspace_len(10) >= 10
    assert L.tb200_gram_workspace_len(8) > 0
    assert L.tb200_fd_rows(1, 4, 4, 0) == 2 * 4 * 3
    assert L.tb200_fd_rows(3, 4, 4, 0) == 3 * 24 + 2 * 16
    # argument errors come back as status codes with a message, never as a crash
    rc = L.tb200_spmv_csr_f64(0, -1, 1, 0, None, None, None, None, None, 0.0, None, None, None, None, None)
    assert rc == 1001 and b"negative" in L.tb200_last_error()
    rc = L.tb200_spmv_csr_f64(7, 1, 1, 0, None, None, None, None, None, 0.0, None, None, None, None, None)
    assert rc == 1001 and b"order" in L.tb200_last_error()


def test_gram_factor_dd_matches_householder_qr():
    """Host double-double Cholesky of a Gram matrix reproduces QR's R / Q^T z to working precision even when the
    Gram matrix is numerically singular in plain double (kappa^2 ~ 1e18)."""
    from trips_b200 import kernels as K
    from fractions import Fraction

    rng = np.random.default_rng(1)
    m, k = 60, 6
    Uo = np.linalg.qr(rng.standard_normal((m, k)))[0]
    Vo = np.linalg.qr(rng.standard_normal((k, k)))[0]
    B = Uo @ np.diag(np.logspace(0, -9, k)) @ Vo.T
    z = rng.standard_normal(m)
    M = np.column_stack((B, z))
    # exact Gram matrix in rational arithmetic, split into hi + lo doubles
    K_ = k + 1
    Ghi, Glo = np.zeros((K_, K_)), np.zeros((K_, K_))
    for i in range(K_):
        for j in range(K_):
            s = sum(Fraction(float(M[r, i])) * Fraction(float(M[r, j])) for r in range(m))
            hi = float(s)
            Ghi[i, j], Glo[i, j] = hi, float(s - Fraction(hi))
    R, C, res2 = K.gram_factor(Ghi, Glo, k)
    Qr, Rr = np.linalg.qr(B)
    sgn = np.sign(np.diag(Rr))
    Rr, Qr = Rr * sgn[:, None], Qr * sgn[None, :]
    assert np.allclose(R, Rr, rtol=0, atol=1e-13 * np.abs(Rr).max())
    # relative accuracy row by row (small singular directions included)
    for i in range(k):
        assert np.linalg.norm(R[i] - Rr[i]) <= 1e-6 * np.linalg.norm(Rr[i])
    assert np.allclose(C[:, 0], Qr.T @ z, atol=1e-7)
    assert abs(res2[0] - np.linalg.norm(z - Qr @ (Qr.T @ z)) ** 2) < 1e-7  # Householder itself carries eps*kappa here


def test_library_is_not_older_than_its_sources():
    """A stale libtripsb200.so (sources edited, `make` not re-run) silently mis-binds arguments: refuse it."""
    import glob

    from trips_b200 import _lib

    srcs = glob.glob(os.path.join(ROOT, "trips-py_b200", "csrc", "*.cu*"))
    newest = max(os.path.getmtime(p) for p in srcs)
    assert os.path.getmtime(_lib.LIB_PATH) >= newest, "rebuild: make -C trips-py_b200/csrc"
