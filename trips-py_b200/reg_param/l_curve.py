"""L-curve rule on the projected problem (NumPy, host) - trips/utilities/reg_param/l_curve.py:190-202.

lambda maximises the curvature of the curve (f, g)(lambda) = (||A x - b||^2, ||L x||^2), x = x(lambda) the Tikhonov
solution of the k x k projected problem; the derivatives of x come from differentiating the normal equations
(l_curve.py:25-87).  The reference recomputes x, x', x'' from scratch inside each of the four derivative terms
(:121-163); here they are formed once per lambda with the same `lstsq` calls on the same matrices, so every number
is the same.  Search: `fminbound` on [1e-9, 2], xtol 1e-12, as the reference (:201).
"""
import numpy as np
import scipy.optimize as op


class _Curve:
    def __init__(self, A, L, b, d=None):
        self.A, self.L = np.asarray(A, dtype=np.float64), np.asarray(L, dtype=np.float64)
        self.b = np.asarray(b, dtype=np.float64).reshape(-1, 1)
        self.d = np.zeros((self.L.shape[0], 1)) if d is None else np.asarray(d, dtype=np.float64).reshape(-1, 1)
        self.C, self.D = self.A.T @ self.A, self.L.T @ self.L  #                               (:141-142)

    def derivatives(self, lam):
        """x, dx/dlambda, d2x/dlambda2 at lambda                                              (:25-87)"""
        A, L, D, d = self.A, self.L, self.D, self.d
        M = self.C + lam * D
        solve = lambda rhs: np.linalg.lstsq(M, rhs, rcond=None)[0]  # noqa: E731
        x = solve(A.T @ self.b + lam * L.T @ d)
        Dx = D @ x
        x1 = -solve(Dx - L.T @ d)
        x2 = 2 * solve(D @ x1 - D @ solve(Dx))
        return x, x1, x2

    def curvature(self, lam):
        """kappa = (-g' f'' + f' g'') / (g'^2 + f'^2)^(3/2)                                    (:171-188)"""
        x, x1, x2 = self.derivatives(lam)

        def term(Op, c):  # first and second derivative of ||Op x - c||^2                     (:89-131)
            r, r1, r2 = Op @ x - c, Op @ x1, Op @ x2
            return (2 * r.T @ r1).item(), (2 * (r1.T @ r1 + r.T @ r2)).item()

        f1, f2 = term(self.A, self.b)
        g1, g2 = term(self.L, self.d)
        return (-g1 * f2 + f1 * g2) / (g1 ** 2 + f1 ** 2) ** 1.5


def l_curve_curvature(lam, A, L, b, d=None):
    return _Curve(A, L, b, d).curvature(lam)


def l_curve(A, L, b, d=None):
    curve = _Curve(A, L, b, d)
    return op.fminbound(func=lambda lam: -curve.curvature(lam), x1=1e-9, x2=2, xtol=1e-12, maxfun=1000, full_output=0, disp=0)
