"""Discrepancy principle for the projected problem (host, NumPy - stays on the CPU as in the reference).

Follows trips/utilities/reg_param/discrepancy_principle.py:19-99 (dptype 'tikhonov'): transform to standard form
when L is not the identity (:37-61), SVD of the projected matrix (:63-65), the "is the discrepancy reachable"
test (:66-74) and Newton's iteration on beta = 1/lambda started at 1e-8, stopped when the update is below
1e-12*beta or after 30 steps (:77-97).  The value returned is the reference's: 1/beta of the last *accepted*
iterate (the reference leaves `alpha` one step behind when the tolerance test fires).

Structural change only: the two quantities that involve the long vectors,  Q^T b  and  ||b - Q Q^T b||,  are inputs
(the device computes them in one streaming pass); everything here is k x k.
"""
import numpy as np
import scipy.linalg as la


def _is_identity(L):
    if L is None:
        return True
    if type(L).__name__ == "Identity":
        return True
    if isinstance(L, np.ndarray) and L.ndim == 2 and L.shape[0] == L.shape[1]:
        return bool(np.allclose(L, np.eye(L.shape[0])))
    return False


def _standard_form(A, L, b):
    """(A L_A^+, b - A x_null) as in the reference (:40-61)."""
    UL, SL, VL = la.svd(L)
    rows, cols = L.shape
    if rows >= cols and SL[-1] != 0:
        return A @ (VL.T @ np.diag(SL ** (-1))), b
    if rows >= cols:
        W = VL[np.where(SL == 0), :].reshape((-1, 1))
    else:
        W = VL[rows - cols:, :].T
    AW = A @ W
    Q_AW, R_AW = np.linalg.qr(AW, mode="reduced")
    Q_LT, R_LT = np.linalg.qr(L.T, mode="reduced")
    LAwpinv = (np.eye(cols) - (W @ np.linalg.inv(R_AW) @ Q_AW.T @ A)) @ Q_LT @ np.linalg.inv(R_LT.T)
    xnull = W @ np.linalg.inv(R_AW) @ Q_AW.T @ b
    return A @ LAwpinv, b - A @ xnull


def discrepancy_principle_projected(A, L, b_proj, resid_norm, delta, eta=1.01, explicitProj=False):
    """A: projected matrix ((k+1) x k bidiagonal / Hessenberg, or k x k R_A); L: identity/None or k x k R_L;
    b_proj = Q^T b (column); resid_norm = ||b - Q Q^T b||.  Returns lambda."""
    if not isinstance(delta, (float, int)):
        raise Exception("""A value for the noise level delta was not provided and the discrepancy principle cannot be applied.
                    Please supply a value of delta based on the estimated noise level of the problem, or choose the regularization parameter according to gcv.""")
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b_proj, dtype=np.float64).reshape(-1, 1)
    if _is_identity(L):
        Anew, bnew = A, b
    else:
        Anew, bnew = _standard_form(A, np.asarray(L, dtype=np.float64), b)
    U, S, _ = la.svd(Anew)
    sv2 = S ** 2
    bhat = U.T @ bnew
    rows, cols = Anew.shape
    target = (eta * delta) ** 2
    out_of_range2 = resid_norm ** 2
    if rows > cols:
        sv2 = np.append(sv2.reshape((-1, 1)), np.zeros((rows - cols, 1)))
        testzero = la.norm(bhat[cols - rows:, :]) ** 2 - target
        if explicitProj:
            testzero += out_of_range2
    else:
        testzero = out_of_range2 - target
    if not testzero < 0:
        return 0
    # Newton on beta = 1/lambda (:77-97), the same arithmetic on 1-D arrays and Python scalars: at k ~ 30 the reference's
    # (k, 1)-array form spends its time in NumPy call overhead (30 steps x a dozen calls), which is host time the
    # device waits for in every iteration of the hybrid solvers
    sv2 = np.ascontiguousarray(sv2, dtype=np.float64).ravel()
    bh = np.ascontiguousarray(bhat, dtype=np.float64).ravel()
    nrm2 = la.blas.dnrm2
    beta = 1e-8
    alpha = None
    iterations = 0
    while (iterations < 30) or ((iterations <= 100) and (abs(alpha) < 10 ** (-16))):
        d = 1.0 / (sv2 * beta + 1)
        zbeta = d * bh
        f = nrm2(zbeta) ** 2 - target
        if explicitProj:
            f += out_of_range2
        wbeta = d * zbeta
        f_prime = 2 / beta * float(zbeta @ (wbeta - zbeta))
        beta_new = beta - f / f_prime
        if abs(beta_new - beta) < 10 ** (-12) * beta:
            if alpha is None:
                alpha = 1 / beta_new
            break
        beta = beta_new
        alpha = 1 / beta_new
        iterations += 1
    return alpha


def discrepancy_principle(Q, A, L, b, delta=None, eta=1.01, **kwargs):
    """Reference signature (discrepancy_principle.py:19) for small host operands: Q, A, L, b NumPy arrays."""
    dptype = kwargs.get("dptype", "tikhonov")
    if dptype != "tikhonov":
        raise NotImplementedError("only dptype='tikhonov' is on the Krylov hot path")
    Q = np.asarray(Q)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 1)
    bp = Q.T @ b
    resid = la.norm(b - Q @ bp)
    return discrepancy_principle_projected(A, L, bp, resid, delta, eta, kwargs.get("explicitProj", False))
