"""TEST INFRASTRUCTURE - CPU oracle for the Krylov hot path of TRIPs-Py (mpasha3/trips-py).

A NumPy/SciPy restatement of the reference's algorithms, written so that every floating-point operation happens
in the same order as in the reference: on identical inputs it reproduces the reference BIT FOR BIT (pinned by
tests/test_oracle.py::test_oracle_is_bit_identical_to_the_reference against the real reference in the build container, and by the golden vectors under
tests/golden/ everywhere else).  It exists because /root/reference cannot travel to the GPU box and because
the reference needs pylops/astra/matplotlib to import.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
Nothing under trips-py_b200/ does: the product has no CPU path.

Each function cites the reference lines it follows (paths relative to the reference root).  Operators passed as
A / L are anything supporting `@` and `.T` (scipy.sparse matrices, ndarrays, FunctionOp below).
"""
import numpy as np
import scipy.linalg as la
import scipy.optimize as op
import scipy.sparse as sp
from scipy.ndimage import convolve


# ======================================================================================================
# reductions
# ======================================================================================================
# The reference takes norms and dot products with np.linalg.norm / np.dot, i.e. OpenBLAS ddot, whose summation order
# depends on the CPU micro-kernel and on the number of BLAS threads.  Golub-Kahan / CGLS without reorthogonalisation
# amplify that last-bit freedom to ~1e-3 in the iterate after 50 iterations on CT problems (measured: DESIGN.md), so
# "the reference's result" is itself only defined up to the BLAS build.  Two modes:
#   'blas'   np.linalg.norm / np.dot                -> bit-identical to the reference on the same machine (pinned)
#   'exact'  correctly rounded exact sum (TwoProduct + math.fsum) -> machine independent; the CUDA path computes the
#            same value with double-double accumulation, so CUDA and oracle agree bit for bit.
# Both are legitimate fp64 evaluations of the same formula and differ by at most one ulp per reduction.
_REDUCTIONS = "blas"


def set_reductions(mode):
    global _REDUCTIONS
    if mode not in ("blas", "exact"):
        raise ValueError(mode)
    prev, _REDUCTIONS = _REDUCTIONS, mode
    return prev


class reductions:
    """Context manager: `with reductions('exact'): ...`"""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = set_reductions(self.mode)

    def __exit__(self, *exc):
        set_reductions(self.prev)


def exact_dot(a, b):
    """Correctly rounded sum_i a_i*b_i: Dekker/Veltkamp TwoProduct (exact without FMA) + math.fsum."""
    import math

    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    b = np.ascontiguousarray(b, dtype=np.float64).ravel()
    p = a * b
    ca = 134217729.0 * a
    ah = ca - (ca - a)
    al = a - ah
    cb = 134217729.0 * b
    bh = cb - (cb - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return math.fsum(np.concatenate((p, e)).tolist())


def _norm(v):
    if _REDUCTIONS == "blas":
        return np.linalg.norm(v)
    return np.float64(np.sqrt(exact_dot(v, v)))


def _dot(a, b):
    if _REDUCTIONS == "blas":
        return np.dot(a, b)
    return np.float64(exact_dot(a, b))


# ======================================================================================================
# Krylov cores                                                   trips/utilities/decompositions.py
# ======================================================================================================

def golub_kahan_update(A, U, S, V):
    """decompositions.py:230-255."""
    k = S.shape[0]
    u_last = U[:, -1]
    if k == 1:
        v = A.T @ u_last
    else:
        v = A.T @ u_last - S[k - 1, k - 2] * V[:, k - 2]
    alpha = _norm(v)
    v = v / alpha
    u = A @ v - alpha * u_last
    beta = _norm(u)
    u = u / beta
    U = np.hstack((U, u.reshape((-1, 1))))
    V = v.reshape((-1, 1)) if k == 1 else np.hstack((V, v.reshape((-1, 1))))
    if k == 1:
        S = np.array([[alpha], [beta]])
    else:
        col = np.zeros(k)
        col[-1] = alpha
        row = np.zeros(k)
        row[-1] = beta
        S = np.vstack((np.hstack((S, col.reshape((-1, 1)))), row.reshape((1, -1))))
    return (U, S, V)


def golub_kahan(A, b, n_iter, dp_stop=False, **kwargs):
    """decompositions.py:118-205."""
    eta = kwargs["gk_eta"] if ("gk_eta" in kwargs) else 1.001
    delta = kwargs["gk_delta"] if ("gk_delta" in kwargs) else 0.001
    res_norm = np.inf
    rows, cols = A.shape
    betas = np.zeros(1)
    alphas = np.zeros(1)
    U = np.zeros((rows, 2))
    V = np.zeros((cols, 1))
    U[:, 0] = (b / _norm(b)).flatten()
    for it in range(n_iter):
        if (dp_stop == True) and (res_norm <= eta * delta):  # noqa: E712                         (:166-168)
            break
        if it != 0:
            U = np.pad(U, ((0, 0), (0, 1)))
            V = np.pad(V, ((0, 0), (0, 1)))
            betas = np.pad(betas, ((0, 1)))
            alphas = np.pad(alphas, ((0, 1)))
        V[:, it] = A.T @ U[:, it] - betas[it - 1] * V[:, it - 1]
        alphas[it] = _norm(V[:, it])
        V[:, it] = V[:, it] / alphas[it]
        U[:, it + 1] = A @ V[:, it] - alphas[it] * U[:, it]
        betas[it] = _norm(U[:, it + 1])
        U[:, it + 1] = U[:, it + 1] / betas[it]
        if dp_stop == True:  # noqa: E712                                                          (:185-195)
            S = np.pad(np.diag(alphas), ((0, 1), (0, 0))) + np.pad(np.diag(betas), ((1, 0), (0, 0)))
            bhat = U.T @ b
            y = np.linalg.lstsq(S, bhat, rcond=None)[0]
            x = V @ y
            res_norm = _norm(A @ x - b)
    k = alphas.shape[0]
    S = np.zeros((k + 1, k))
    S[range(0, k), range(0, k)] = alphas
    S[range(1, k + 1), range(0, k)] = betas
    return (U, S, V)


def arnoldi_update(A, V, H):
    """decompositions.py:207-228 (modified Gram-Schmidt, single pass)."""
    k = H.shape[0]
    w = A @ V[:, -1]
    h = np.zeros((k, 1))
    for j in range(k):
        h[j] = _dot(V[:, j], w)
        w = w - h[j] * V[:, j]
    H = h if k == 1 else np.hstack((H, h))
    last = np.zeros((1, k))
    last[:, -1] = _norm(w)
    H = np.vstack((H, last))
    V = np.hstack((V, w.reshape((-1, 1)) / H[-1, -1]))
    return (V, H)


def arnoldi_update_cgs2(A, V, H):
    """The north-star variant (NOT reference behaviour): classical Gram-Schmidt applied twice."""
    k = H.shape[0]
    w = A @ V[:, -1]
    h1 = V.T @ w
    w = w - V @ h1
    h2 = V.T @ w
    w = w - V @ h2
    h = (h1 + h2).reshape(-1, 1)
    H = h if k == 1 else np.hstack((H, h))
    last = np.zeros((1, k))
    last[:, -1] = _norm(w)
    H = np.vstack((H, last))
    V = np.hstack((V, w.reshape((-1, 1)) / H[-1, -1]))
    return (V, H)


# ======================================================================================================
# parameter-choice rules                                          trips/utilities/reg_param/
# ======================================================================================================

def _todense(M):
    return M.todense() if hasattr(M, "todense") and not isinstance(M, np.ndarray) else M


def gcv_numerator(lam, Q_A, R_A, R_L, b):
    """gcv.py:25-49, 'standard' variant (the reference never forwards kwargs to it: gcv.py:94)."""
    RA2 = _todense(R_A.T @ R_A)
    RL2 = _todense(R_L.T @ R_L)
    inv = la.solve((RA2 + lam * RL2), (R_A.T @ Q_A.T @ b))
    return (_norm(R_A @ inv - Q_A.T @ b)) ** 2


def gcv_denominator(lam, R_A, R_L, b, **kwargs):
    """gcv.py:51-78."""
    variant = kwargs["variant"] if ("variant" in kwargs) else "standard"
    RA2 = _todense(R_A.T @ R_A)
    RL2 = _todense(R_L.T @ R_L)
    inv = la.solve((RA2 + lam * RL2), R_A.T)
    if variant == "modified":
        trace_term = kwargs["fullsize"] - np.trace(R_A @ inv)
    else:
        trace_term = R_A.shape[0] - np.trace(R_A @ inv)
    return trace_term ** 2


def generalized_crossvalidation(Q_A, R_A, R_L, b, **kwargs):
    """gcv.py:80-95, gcvtype='tikhonov'."""
    f = lambda lam: gcv_numerator(lam, Q_A, R_A, R_L, b) / gcv_denominator(lam, R_A, R_L, b, **kwargs)  # noqa: E731
    return op.fminbound(func=f, x1=1e-09, x2=1e2, args=(), xtol=1e-12, maxfun=1000, full_output=0, disp=0)


def _lc_x(l, C, D, B, E, b, d):
    """l_curve.py:25-45: (C + l D) x = B^T b + l E^T d."""
    return np.linalg.lstsq(C + l * D, B.T @ b + l * E.T @ d, rcond=None)[0]


def _lc_dx(l, C, D, E, x_l, d):
    """l_curve.py:47-67."""
    return -np.linalg.lstsq(C + l * D, D @ x_l - E.T @ d, rcond=None)[0]


def _lc_d2x(l, C, D, E, x_l, dx_l, d):
    """l_curve.py:69-87."""
    lhs = C + l * D
    inv4 = np.linalg.lstsq(lhs, D @ x_l, rcond=None)[0]
    return 2 * np.linalg.lstsq(lhs, D @ dx_l - D @ inv4, rcond=None)[0]


def _lc_term(l, Op, A, L, b, c, d, order):
    """First (order 1) or second (order 2) derivative of ||Op x_l - c||^2   (l_curve.py:89-131, callers :133-169)."""
    C, D = A.T @ A, L.T @ L
    x_l = _lc_x(l, C, D, A, L, b, d)
    dx = _lc_dx(l, C, D, L, x_l, d)
    f, f1 = Op @ x_l, Op @ dx
    if order == 1:
        return (2 * (f - c).T @ f1).item()
    f2 = Op @ _lc_d2x(l, C, D, L, x_l, dx, d)
    return (2 * (f1.T @ f1 + (f - c).T @ f2)).item()


def l_curve_curvature(l, A, L, b, d=None):
    """l_curve.py:171-188."""
    b = np.asarray(b).reshape(-1, 1)
    d = np.zeros((L.shape[0], 1)) if d is None else d
    f1, f2 = _lc_term(l, A, A, L, b, b, d, 1), _lc_term(l, A, A, L, b, b, d, 2)
    g1, g2 = _lc_term(l, L, A, L, b, d, d, 1), _lc_term(l, L, A, L, b, d, d, 2)
    return (-g1 * f2 + f1 * g2) / (g1 ** 2 + f1 ** 2) ** 1.5


def l_curve(A, L, b, d=None):
    """l_curve.py:190-202."""
    return op.fminbound(func=lambda l: -1 * l_curve_curvature(l, A, L, b, d), x1=1e-9, x2=2, xtol=1e-12, maxfun=1000,
                        full_output=0, disp=0)


class IdentityOp:
    """Stand-in for pylops.Identity where the reference passes one as L / R_L (Hybrid_LSQR.py:76, Hybrid_GMRES.py:56)."""

    def __init__(self, n):
        self.shape = (n, n)

    @property
    def T(self):
        return self

    def __matmul__(self, x):
        return self if isinstance(x, IdentityOp) else x

    def todense(self):
        return np.eye(self.shape[0])


def is_identity(L):
    """trips/utilities/utils.py:47-62."""
    if isinstance(L, IdentityOp):
        return True
    if isinstance(L, np.ndarray) and (L.shape[0] == L.shape[1]) and np.allclose(L, np.eye(L.shape[0])):
        return True
    if sp.issparse(L) and (L.shape[0] == L.shape[1]) and (L - sp.eye(L.shape[0])).sum() < 10 ** (-6):
        return True
    return False


def discrepancy_principle(Q, A, L, b, delta=None, eta=1.01, **kwargs):
    """discrepancy_principle.py:19-99, dptype='tikhonov'."""
    if not (isinstance(delta, float) or isinstance(delta, int)):
        raise Exception("A value for the noise level delta was not provided and the discrepancy principle cannot be applied.")
    explicitProj = kwargs["explicitProj"] if ("explicitProj" in kwargs) else False
    bfull = b
    b = Q.T @ b
    if is_identity(L):
        Anew, bnew = A, b
    else:
        UL, SL, VL = la.svd(L)
        if L.shape[0] >= L.shape[1] and SL[-1] != 0:
            Anew = A @ (VL.T @ np.diag((SL) ** (-1)))
            bnew = b
        else:
            if L.shape[0] >= L.shape[1]:
                W = VL[np.where(SL == 0), :].reshape((-1, 1))
            else:
                W = VL[L.shape[0] - L.shape[1]:, :].T
            AW = A @ W
            Q_AW, R_AW = np.linalg.qr(AW, mode="reduced")
            Q_LT, R_LT = np.linalg.qr(L.T, mode="reduced")
            LAwpinv = (np.eye(L.shape[1]) - (W @ np.linalg.inv(R_AW) @ Q_AW.T @ A)) @ Q_LT @ np.linalg.inv(R_LT.T)
            Anew = A @ LAwpinv
            xnull = W @ np.linalg.inv(R_AW) @ Q_AW.T @ b
            bnew = b - A @ xnull
    U, S, V = la.svd(Anew)
    sv = S ** 2
    bhat = U.T @ bnew
    if Anew.shape[0] > Anew.shape[1]:
        sv = np.append(sv.reshape((-1, 1)), np.zeros((Anew.shape[0] - Anew.shape[1], 1)))
        if explicitProj:
            testzero = la.norm(bhat[Anew.shape[1] - Anew.shape[0]:, :]) ** 2 + la.norm(bfull - Q @ b) ** 2 - (eta * delta) ** 2
        else:
            testzero = la.norm(bhat[Anew.shape[1] - Anew.shape[0]:, :]) ** 2 - (eta * delta) ** 2
    else:
        testzero = la.norm(bfull - Q @ b) ** 2 - (eta * delta) ** 2
    sv.shape = (sv.shape[0], 1)
    beta = 1e-8
    iterations = 0
    if testzero < 0:
        while (iterations < 30) or ((iterations <= 100) and (np.abs(alpha) < 10 ** (-16))):
            zbeta = (((sv * beta + 1) ** (-1)) * bhat.reshape((-1, 1))).reshape((-1, 1))
            if explicitProj:
                f = la.norm(zbeta) ** 2 + la.norm(bfull - Q @ b) ** 2 - (eta * delta) ** 2
            else:
                f = la.norm(zbeta) ** 2 - (eta * delta) ** 2
            wbeta = (((sv * beta + 1) ** (-1)) * zbeta).reshape((-1, 1))
            f_prime = 2 / beta * zbeta.T @ (wbeta - zbeta)
            beta_new = beta - f / f_prime
            if abs(beta_new - beta) < 10 ** (-12) * beta:
                break
            beta = beta_new
            alpha = 1 / beta_new[0, 0]
            iterations += 1
    else:
        alpha = 0
    return alpha


# ======================================================================================================
# solvers                                                                   trips/solvers/
# ======================================================================================================

def _stack_solve(B, L, rhs_top, lam):
    return np.linalg.lstsq(np.vstack((B, np.sqrt(lam) * L)),
                           np.vstack((rhs_top.reshape((-1, 1)), np.zeros((B.shape[1], 1)))), rcond=None)[0]


def CGLS(A, b, x0, max_iter, tol, x_true=None):
    """CGLS.py:42-86 (with np.eps -> machine epsilon, never reached on the test problems)."""
    b = b.reshape((-1, 1))
    x = x0
    r = b - A @ x
    t = A.T @ r
    p = t
    x_history, rel_residual, rel_error = [], [], []
    norms_t0 = _norm(t)
    gamma, xmax = norms_t0 ** 2, _norm(x)
    k, check = 0, 0
    while (k < max_iter) and (check == 0):
        x_old = x
        k += 1
        w = A @ p
        delta = _norm(w) ** 2
        if delta == 0:
            delta = np.finfo(float).eps
        beta = gamma / delta
        x = x + beta * p
        x_history.append(x)
        r = r - beta * w
        t = A.T @ r
        gamma_old = gamma
        norm_t = _norm(t)
        gamma = norm_t ** 2
        p = t + (gamma / gamma_old) * p
        norm_x = _norm(x)
        xmax = max(xmax, norm_x)
        check = (norm_t <= norms_t0 * tol) or (norm_x * tol >= 1)
        rel_residual.append(_norm(x - x_old) / _norm(x))
        if x_true is not None:
            rel_error.append(_norm(x - x_true) / _norm(x))
    info = {"xHistory": x_history, "regParam": [], "relResidual": rel_residual, "its": k}
    if x_true is not None:
        info["relError"] = rel_error
    return (x, info)


def Hybrid_LSQR(A, b, n_iter=100, regparam="gcv", x_true=None, **kwargs):
    """Hybrid_LSQR.py:55-114 (dp_stop=False)."""
    n = A.shape[1]
    beta = _norm(b)
    U = b.reshape((-1, 1)) / beta
    B = np.empty(1)
    V = np.empty((n, 1))
    x_history, lambda_history = [], []
    bhat = np.zeros(1)
    bhat[0] = beta
    for ii in range(n_iter):
        (U, B, V) = golub_kahan_update(A, U, B, V)
        bhat = np.append(bhat, 0)
        k = B.shape[1]
        if ii == 0:
            lambdah = 0
            continue
        if regparam == "gcv":
            Q_A, s, _ = la.svd(B, full_matrices=False)
            lambdah = generalized_crossvalidation(Q_A, np.diag(s), np.eye(k), bhat, variant="modified",
                                                  fullsize=A.shape[0], **kwargs)
        elif regparam == "dp":
            lambdah = discrepancy_principle(U, B, IdentityOp(k), b, **kwargs)
        elif regparam == "l_curve":  # Hybrid_LSQR.py:94-98
            Q_A, s, _ = la.svd(B, full_matrices=False)
            lambdah = l_curve(np.diag(s), np.eye(k), Q_A.T @ bhat.reshape((-1, 1)))
        else:
            lambdah = regparam
        lambda_history.append(lambdah)
        y = _stack_solve(B, np.eye(k), bhat, lambdah)
        x = (V @ y).reshape((-1, 1))
        x_history.append(x)
    info = {"xHistory": x_history, "regParam": lambdah, "regParam_history": lambda_history, "relResidual": [], "its": ii,
            "B": B, "U": U, "V": V}
    if x_true is not None:
        xt = x_true.reshape(-1, 1)
        info["relError"] = [la.norm(xx - xt) / la.norm(xt) for xx in x_history]
    return (x, info)


def Hybrid_GMRES(A, b, n_iter, regparam="gcv", x_true=None, reorth="mgs", **kwargs):
    """Hybrid_GMRES.py:33-87 (dp_stop=False).  relResidual holds ||bhat - H y|| (the reference's :80 broadcasts)."""
    n = A.shape[1]
    if A.shape[0] != n:
        raise Exception("Please check the size of the matrx A: it should be square in order to apply hybrid GMRES")
    x_history, lambda_history, residual_history = [], [], []
    beta = _norm(b)
    V = b.reshape((-1, 1)) / beta
    H = np.empty(1)
    bhat = np.zeros(1)
    bhat[0] = beta
    step = arnoldi_update if reorth == "mgs" else arnoldi_update_cgs2
    for ii in range(n_iter):
        (V, H) = step(A, V, H)
        bhat = np.append(bhat, 0)
        k = H.shape[1]
        if ii == 0:
            lambdah = 0
        elif regparam == "gcv":
            Q_A, s, _ = la.svd(H, full_matrices=False)
            lambdah = generalized_crossvalidation(Q_A, np.diag(s), IdentityOp(k), bhat, **kwargs)
        elif regparam == "dp":
            lambdah = discrepancy_principle(V, H, IdentityOp(k), b, **kwargs)
        elif regparam == "l_curve":  # Hybrid_GMRES.py:67-71
            Q_A, s, _ = la.svd(H, full_matrices=False)
            lambdah = l_curve(np.diag(s), np.eye(k), Q_A.T @ bhat.reshape((-1, 1)))
        else:
            lambdah = regparam
        lambda_history.append(lambdah)
        y = _stack_solve(H, np.eye(k), bhat, lambdah)
        x = (V[:, :-1] @ y).reshape((-1, 1))
        x_history.append(x)
        residual_history.append(la.norm(bhat.reshape((-1, 1)) - H @ y))
    info = {"xHistory": x_history, "regParam": lambdah, "regParam_history": lambda_history,
            "relResidual": residual_history, "its": ii, "H": H, "V": V}
    if x_true is not None:
        xt = x_true.reshape(-1, 1)
        info["relError"] = [la.norm(xx - xt) / la.norm(xt) for xx in x_history]
    return (x, info)


def GKS(A, b, L, projection_dim=3, n_iter=50, regparam="gcv", x_true=None, **kwargs):
    """GKS.py:36-105, QR branch (L not the identity)."""
    dp_stop = kwargs["dp_stop"] if ("dp_stop" in kwargs) else False
    (U, B, V) = golub_kahan(A, b, projection_dim, dp_stop, **kwargs)  # TypeError when dp_stop is in kwargs, as the reference
    AV = A @ V
    LV = L @ V
    x_history, lambda_history, residual_history = [], [], []
    for ii in range(n_iter):
        (Q_A, R_A) = la.qr(AV, mode="economic")
        _, R_L = la.qr(LV, mode="economic")
        if regparam == "gcv":
            lambdah = generalized_crossvalidation(Q_A, R_A, R_L, b, **kwargs)
        elif regparam == "dp":
            lambdah = discrepancy_principle(Q_A, R_A, R_L, b, **kwargs)
        elif regparam == "l_curve":  # GKS.py:67-68
            lambdah = l_curve(R_A, R_L, Q_A.T @ b)
        else:
            lambdah = regparam
        lambda_history.append(lambdah)
        y, _, _, _ = np.linalg.lstsq(np.concatenate((R_A, np.sqrt(lambdah) * R_L)),
                                     np.concatenate((Q_A.T @ b, np.zeros((R_L.shape[0], 1)))), rcond=None)
        x = V @ y
        x_history.append(x)
        ra = AV @ y - b
        ra = A.T @ ra
        rb = (LV @ y)
        rb = L.T @ rb
        r = ra + lambdah * rb
        r = r - V @ (V.T @ r)
        r = r - V @ (V.T @ r)
        r = r - V @ (V.T @ r)
        residual_history.append(la.norm(r))
        vn = r / _norm(r)
        V = np.column_stack((V, vn))
        AV = np.column_stack((AV, A @ vn))
        LV = np.column_stack((LV, L @ vn))
    info = {"xHistory": x_history, "regParam": lambdah, "regParam_history": lambda_history,
            "Residual": residual_history, "its": ii}
    if x_true is not None:
        xt = x_true.reshape(-1, 1)
        info["relError"] = [la.norm(xx - xt) / la.norm(xt) for xx in x_history]
    return (x, info)


def smoothed_holder_weights(x, epsilon, p):
    """weights.py:66-68."""
    return (x ** 2 + epsilon ** 2) ** (p / 2 - 1)


def MMGKS(A, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=5, regparam="gcv", x_true=None, **kwargs):
    """MMGKS.py:37-137, default (anisotropic) weights; isoTV via kwargs isoTV='isoTV', Ls=<2N x N gradient>;
    group sparsity via GS='GS', prob_dims=(nx, ny, nt) (:45-52, 79-91: L is REPLACED by kron(I_nt, old 2-D differences))."""
    epsilon = kwargs["epsilon"] if ("epsilon" in kwargs) else 0.1
    iso_L = kwargs.pop("iso_Ls", None)
    gs = kwargs.pop("GS", False) in ["GS", "gs", "Gs"]
    prob_dims = kwargs.pop("prob_dims", False)
    dp_stop = kwargs["dp_stop"] if ("dp_stop" in kwargs) else False
    (U, B, V) = golub_kahan(A, b, projection_dim, dp_stop, **kwargs)  # TypeError when dp_stop is in kwargs, as the reference
    x_history, lambda_history, residual_history = [], [], []
    x = A.T @ b
    AV = A @ V
    if gs:
        nx, ny, nt = prob_dims
        Ls = first_derivative_old_2d(nx, ny)
        L = sp.kron(sp.identity(nt), Ls).tocsr()
    LV = L @ V
    for ii in range(n_iter):
        v = A @ x - b
        wf = (v ** 2 + epsilon ** 2) ** (pnorm / 2 - 1)
        AA = AV * wf
        (Q_A, R_A) = la.qr(AA, mode="economic")
        u = L @ x
        if iso_L is not None:
            # MMGKS.py:64-78 with nt = 1 and an fp64 statement of the gradient operator
            spacen = int(iso_L.shape[0] / 2)
            LsX = iso_L @ x.reshape(-1, 1)
            weightx = (LsX[:spacen, :] ** 2 + LsX[spacen:2 * spacen, :] ** 2 + epsilon ** 2) ** ((qnorm - 2) / 4)
            weightx = np.concatenate((weightx.flatten(), weightx.flatten()))
            weightt = (u[2 * spacen:] ** 2 + epsilon ** 2) ** ((qnorm - 2) / 4)
            wr = np.concatenate((weightx.reshape(-1, 1), weightt))
        elif gs:  # MMGKS.py:79-91 (C-order reshape of the frame-major x, exp(2) as the smoothing constant: as written)
            nt_ = int(x.reshape((-1, 1)).shape[0] / (nx * ny))
            Dutemp = Ls.dot(np.reshape(x, (nx * ny, nt_)))
            wr = np.exp(2) * np.ones((2 * nx * (ny - 1), 1))
            for i in range(2 * nx * (ny - 1)):
                wr[i] = (np.linalg.norm(Dutemp[i, :]) ** 2 + wr[i]) ** (qnorm / 2 - 1)
            wr = np.kron(np.ones((nt_, 1)), wr)
        else:
            wr = smoothed_holder_weights(u, epsilon=epsilon, p=qnorm).reshape((-1, 1))
        LL = LV * wr
        (Q_L, R_L) = la.qr(LL, mode="economic")
        if regparam == "gcv":
            lambdah = generalized_crossvalidation(Q_A, R_A, R_L, wf * b, **kwargs)
        elif regparam == "dp":
            lambdah = discrepancy_principle(Q_A, R_A, R_L, wf * b, **kwargs)
        elif regparam == "l_curve":  # MMGKS.py:100-101 (unweighted b)
            lambdah = l_curve(R_A, R_L, Q_A.T @ b)
        else:
            lambdah = regparam
        lambda_history.append(lambdah)
        y, _, _, _ = np.linalg.lstsq(np.concatenate((R_A, np.sqrt(lambdah) * R_L)),
                                     np.concatenate((Q_A.T @ b, np.zeros((R_L.shape[0], 1)))), rcond=None)
        x = V @ y
        x_history.append(x)
        if ii >= R_L.shape[0]:
            break
        ra = wf * (AV @ y - b)
        ra = A.T @ ra
        rb = wr * (LV @ y)
        rb = L.T @ rb
        r = ra + lambdah * rb
        r = r - V @ (V.T @ r)
        r = r - V @ (V.T @ r)
        vn = r / _norm(r)
        V = np.column_stack((V, vn))
        AV = np.column_stack((AV, A @ vn))
        LV = np.column_stack((LV, L @ vn))
        residual_history.append(la.norm(r))
    info = {"xHistory": x_history, "regParam": lambdah, "regParam_history": lambda_history,
            "Residual": residual_history, "its": ii}
    if x_true is not None:
        info["relError"] = [la.norm(xx - x_true) / la.norm(x_true) for xx in x_history]
    return (x, info)


# ======================================================================================================
# operators
# ======================================================================================================

def first_derivative_1d(n):
    """operators.py:24-28, built in CSR (the reference's DIA slicing raises under scipy >= 1.15: SURVEY.md F4):
    rows i = 0..n-2 of I - shift(+1), i.e. (L x)_i = x_i - x_{i+1}."""
    D = sp.diags(np.ones(n - 1), offsets=1, format="csr")
    return (sp.identity(n, format="csr") - D).tocsr()[0:-1, :]


def first_derivative_2d(nx, ny):
    """operators.py:30-36."""
    IDx = sp.kron(sp.identity(nx), first_derivative_1d(nx))
    DyI = sp.kron(first_derivative_1d(ny), sp.identity(ny))
    return sp.vstack((IDx, DyI)).tocsr()


def first_derivative_old_1d(n):
    """operators_old.py:66-72 (generate_first_derivative_operator_matrix): rows 0..n-2 of I - subdiag(1), i.e. row 0 is
    x_0 and row i is x_i - x_{i-1}.  (.tocsr(): the dia_matrix of the reference is not subscriptable under scipy 1.18.)"""
    D = sp.spdiags(data=np.ones(n - 1), diags=-1, m=n, n=n)
    return (sp.identity(n, format="csr") - D).tocsr()[0:-1, :]


def first_derivative_old_2d(nx, ny):
    """operators_old.py:75-85 (generate_first_derivative_operator_2d_matrix)."""
    IDx = sp.kron(sp.identity(nx), first_derivative_old_1d(nx))
    DyI = sp.kron(first_derivative_old_1d(ny), sp.identity(ny))
    return sp.vstack((IDx, DyI)).tocsr()


def framelet_filters(l, n):
    """operators.py:50-83 (construct_H)."""
    e = np.ones((n,))
    H_0 = (sp.spdiags(e, -l, n, n) + sp.spdiags(2 * e, 0, n, n) + sp.spdiags(e, l, n, n)).tolil()
    H_1 = (sp.spdiags(-e, -l, n, n) + sp.spdiags(e, l, n, n)).tolil()
    H_2 = (sp.spdiags(-e, -l, n, n) + sp.spdiags(2 * e, 0, n, n) + sp.spdiags(-e, l, n, n)).tolil()
    for jj in range(0, l):
        H_0[jj, l - jj - 1] += 1
        H_0[-jj - 1, -l + jj] += 1
        H_1[jj, l - jj - 1] -= 1
        H_1[-jj - 1, -l + jj] += 1
        H_2[jj, l - jj - 1] -= 1
        H_2[-jj - 1, -l + jj] -= 1
    return (H_0.tocsr() / 4).tocsr(), (H_1.tocsr() * (np.sqrt(2) / 4)).tocsr(), (H_2.tocsr() / 4).tocsr()


def framelet_analysis(n, l):
    """operators.py:86-101 (create_analysis_operator_rec / create_analysis_operator)."""
    def rec(level, w):
        if level == l:
            return sp.vstack(framelet_filters(level, n))
        (H_0, H_1, H_2) = framelet_filters(level, n)
        W_1 = rec(level + 1, H_0)
        return sp.vstack((W_1, H_1, H_2)) * w
    return sp.csr_matrix(rec(1, 1))


def framelet_operator(n, m, l):
    """operators.py:104-113 (create_framelet_operator; `.H` of a real sparse matrix written as `.T`, which is what
    scipy >= 1.14 still offers)."""
    W_n, W_m = framelet_analysis(n, l), framelet_analysis(m, l)
    k = 2 * l + 1
    fwd = lambda x: (W_n @ (np.asarray(x).reshape(n, m, order="F") @ W_m.T)).reshape(-1, 1, order="F")  # noqa: E731
    bwd = lambda x: (W_n.T @ (np.asarray(x).reshape(n * k, m * k, order="F") @ W_m)).reshape(-1, 1, order="F")  # noqa: E731
    return FunctionOp(fwd, bwd, (n * k * m * k, n * m))


def spacetime_derivative(nx, ny, nt):
    """operators.py:39-45."""
    ITLs = sp.kron(sp.identity(nt), first_derivative_2d(nx, ny))
    LTIN = sp.kron(first_derivative_1d(nt), sp.identity(nx ** 2))
    return sp.vstack((ITLs, LTIN)).tocsr()


def centered_derivative_2d(nx, ny):
    """fp64 statement of operators_old.first_derivative_operator_2d (operators_old.py:35-45): pylops' 3-point centred
    first derivative (zero rows at both ends) in VStack(Kronecker(I, D), Kronecker(D, I)) form; 2*nx*ny rows."""
    def D(n):
        M = sp.lil_matrix((n, n))
        for i in range(1, n - 1):
            M[i, i + 1] = 0.5
            M[i, i - 1] = -0.5
        return M.tocsr()
    return sp.vstack((sp.kron(sp.identity(nx), D(nx)), sp.kron(D(ny), sp.identity(ny)))).tocsr()


class FunctionOp:
    """Matrix-free operator with `@` and `.T` (stands for pylops.FunctionOperator, Deblurring2D.py:72)."""

    def __init__(self, f, fT, shape):
        self.f, self.fT, self.shape = f, fT, shape

    def _apply(self, fn, x):
        x = np.asarray(x)
        if x.ndim == 2 and x.shape[1] > 1:
            return np.stack([fn(x[:, j]).reshape(-1) for j in range(x.shape[1])], axis=1)
        y = fn(x).reshape(-1)
        return y.reshape(-1, 1) if x.ndim == 2 else y

    def __matmul__(self, x):
        return self._apply(self.f, x)

    @property
    def T(self):
        return FunctionOp(self.fT, self.f, (self.shape[1], self.shape[0]))


def gauss_psf(dim, spread):
    """Deblurring2D.py:48-64 (Gauss), without the centre lookup."""
    m, n = dim[0], dim[1]
    s1, s2 = (spread, spread) if type(spread) in [int] else (spread[0], spread[1])
    x = np.arange(-np.fix(n / 2), np.ceil(n / 2))
    y = np.arange(-np.fix(m / 2), np.ceil(m / 2))
    X, Y = np.meshgrid(x, y)
    PSF = np.exp(-0.5 * ((X ** 2) / (s1 ** 2) + (Y ** 2) / (s2 ** 2)))
    PSF /= PSF.sum()
    return PSF


def blur_operator(PSF, nx, ny):
    """Deblurring2D.py:66-73 (forward_Op)."""
    fwd = lambda X: convolve(X.reshape([nx, ny]), PSF, mode="reflect").reshape((-1, 1))  # noqa: E731
    bwd = lambda B: convolve(B.reshape([nx, ny]), np.flipud(np.fliplr(PSF)), mode="reflect").reshape((-1, 1))  # noqa: E731
    return FunctionOp(fwd, bwd, (nx * ny, nx * ny))


def blur_data(x, PSF, nx, ny):
    """Deblurring2D.py:119-133 (gen_data, CommitCrime=False): blur on a zero-padded 2x image, crop the centre."""
    big = np.zeros((2 * nx, 2 * ny))
    px, py = nx // 2, ny // 2
    big[px:px + nx, py:py + ny] = x.reshape((nx, ny))
    b = convolve(big, PSF, mode="constant")
    return b[px:px + nx, py:py + ny].reshape((-1, 1))


# ---- parallel-beam CT matrix (new synthetic problem; geometry conventions of Tomography.py:53-56) --------------

def ct_angles(views):
    return np.linspace(0, np.pi, views, endpoint=False)


def ct_num_detectors(nx):
    return int(np.sqrt(2) * nx)


def _ray_tables(c, s, n_det, fan):
    """Per-detector line parameters at one angle: unit normal (cd, sd), signed offset rho, and the trapezoid constants.
    Parallel beam: the same normal for every bin, rho = d - (n_det-1)/2.  Fan beam (so, dd, dps): ASTRA 'fanflat'
    conventions (Tomography.py:57-67) - source so*(sin, -cos), detector centre dd*(-sin, cos), axis (cos, sin).
    Same operations in the same order as tb200_ctgeom.cuh ray_geometry / make_geom."""
    k = np.arange(n_det) - 0.5 * (n_det - 1)
    if fan is None:
        cd, sd, rho = np.full(n_det, c), np.full(n_det, s), k
    else:
        so, dd, dps = (np.float64(v) for v in fan)
        sx, sy = so * s, -(so * c)
        px = -(dd * s) + k * (c * dps)
        py = (dd * c) + k * (s * dps)
        ex, ey = px - sx, py - sy
        ln = np.sqrt(ex * ex + ey * ey)
        cd, sd = ey / ln, -(ex / ln)
        rho = cd * sx + sd * sy
    hi, lo = np.maximum(np.abs(cd), np.abs(sd)), np.minimum(np.abs(cd), np.abs(sd))
    d2 = 0.5 * (hi + lo)
    with np.errstate(divide="ignore"):
        inv_hi, inv_hilo = 1.0 / hi, 1.0 / (hi * lo)  # lo == 0: inf, the min below returns the plateau
    return cd, sd, rho, d2, inv_hi, inv_hilo


def ct_matrix(nx, theta, ny=None, n_det=None, fan=None):
    """Dense-loop-free NumPy statement of the device builder (trips-py_b200/csrc/ct_builder.cu): entry = chord length
    of ray (angle a, detector d) through unit pixel (iy, ix); row = a*n_det + d, column = iy*nx + ix.
    Every arithmetic step is a separately rounded IEEE operation in the same order as the CUDA code.
    fan = (source_origin, detector_origin, detector_pixel_size) selects the flat-detector fan beam."""
    ny = nx if ny is None else ny
    n_det = ct_num_detectors(nx) if n_det is None else n_det
    theta = np.asarray(theta, dtype=np.float64)
    cx = np.arange(nx) - 0.5 * (nx - 1)
    cy = np.arange(ny) - 0.5 * (ny - 1)
    rows, cols, vals = [], [], []
    for a, (c, s) in enumerate(zip(np.cos(theta), np.sin(theta))):
        cd, sd, rho, d2, inv_hi, inv_hilo = _ray_tables(c, s, n_det, fan)
        # candidate detectors around each pixel's projection (any estimate within a bin or two will do: the exact
        # predicate below decides)
        if fan is None:
            centre = (cx[None, :] * c) + (cy[:, None] * s) + 0.5 * (n_det - 1)
            offsets = (-1, 0, 1, 2)
        else:
            so, dd, dps = fan
            qx, qy = cx[None, :] - so * s, cy[:, None] + so * c
            depth, lateral = -qx * s + qy * c, qx * c + qy * s
            centre = lateral * (so + dd) / (depth * dps) + 0.5 * (n_det - 1)
            offsets = (-2, -1, 0, 1, 2, 3)
        base = np.floor(np.clip(centre, -4, n_det + 4)).astype(np.int64)
        for off in offsets:
            d = base + off
            ok = (d >= 0) & (d < n_det)
            dd_ = np.clip(d, 0, n_det - 1)
            proj = (cx[None, :] * cd[dd_]) + (cy[:, None] * sd[dd_])      # cx*c + cy*s with the ray's own normal
            t = rho[dd_] - proj
            at = np.abs(t)
            hit = ok & (at < d2[dd_])
            with np.errstate(invalid="ignore"):
                w = np.minimum(inv_hi[dd_], (d2[dd_] - at) * inv_hilo[dd_])  # plateau 1/hi, slopes (d2-|t|)/(hi*lo)
            iy, ix = np.nonzero(hit)
            rows.append(a * n_det + dd_[iy, ix])
            cols.append(iy * nx + ix)
            vals.append(w[iy, ix])
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(len(theta) * n_det, nx * ny))
    A.sort_indices()
    return A


def fan_geometry(nx):
    """(source_origin, detector_origin, detector_pixel_size) of Tomography.define_proj_id (Tomography.py:57-59)."""
    so, dd = 3.0 * nx, 1.0 * nx
    return so, dd, (so + dd) / so


def shepp_logan(n):
    """Modified Shepp-Logan phantom: the standard ten-ellipse table on the grid [-1,1]^2 with spacing 2/(n-1),
    negatives clipped - the construction of trips/utilities/phantoms.py:18-60; n x n, float64."""
    ell = [(1.0, .69, .92, 0, 0, 0), (-.8, .6624, .8740, 0, -.0184, 0), (-.2, .1100, .3100, .22, 0, -18),
           (-.2, .1600, .4100, -.22, 0, 18), (.1, .2100, .2500, 0, .35, 0), (.1, .0460, .0460, 0, .1, 0),
           (.1, .0460, .0460, 0, -.1, 0), (.1, .0460, .0230, -.08, -.605, 0), (.1, .0230, .0230, 0, -.606, 0),
           (.1, .0230, .0460, .06, -.605, 0)]
    g = (np.arange(n) - (n - 1) / 2) / ((n - 1) / 2)
    X, Y = np.meshgrid(g, -g)
    img = np.zeros((n, n))
    for A_, a, b, x0, y0, phi in ell:
        ph = phi * np.pi / 180
        xr = (X - x0) * np.cos(ph) + (Y - y0) * np.sin(ph)
        yr = (Y - y0) * np.cos(ph) - (X - x0) * np.sin(ph)
        img[(xr ** 2) / a ** 2 + (yr ** 2) / b ** 2 <= 1] += A_
    img[img < 0] = 0
    return img


def add_noise(b_true, level, rng):
    """Tomography.py:203-212 / Deblurring2D.py:141-147 with a seeded generator: e = level*||b||/||n|| * n."""
    noise = rng.standard_normal(b_true.shape[0]).reshape((-1, 1))
    e = level * np.linalg.norm(b_true) / np.linalg.norm(noise) * noise
    return b_true.reshape((-1, 1)) + e, la.norm(e)
