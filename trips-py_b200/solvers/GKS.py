"""Generalised Krylov subspace method on the GPU - signature and returns of trips/solvers/GKS.py:27-105.

min ||A x - b||^2 + lambda ||L x||^2 over an expanding subspace span(V): Golub-Kahan start (projection_dim
steps), then per iteration: factors of AV and LV (here: one Gram pass each, section "Replacing the QRs" in
_gks_core.py), lambda by GCV / discrepancy principle / fixed, projected Tikhonov solve on the host, lift,
normal-equations residual r = A^T(AV y - b) + lambda L^T(LV y), three classical Gram-Schmidt sweeps against V
(GKS.py:86-88), and expansion of V, AV, LV by pointer bump.

Deviation (documented in DESIGN.md): when `L` is the identity the reference takes an SVD branch (GKS.py:44-50) that
drops the right singular vectors of AV before solving; this implementation always uses the triangular-factor
branch, which is the mathematically consistent one.
"""
import torch

from .. import kernels as K
from ..operators import as_operator, to_device_vector
from ._common import ErrorTracker, LazyHistory, dev_scalar, host_column, need_delta, tikhonov_projected, single_threaded_host_blas
from ._gks_core import GKSBases, adjoint_L_weighted, choose_lambda, expand, factor_pair


@single_threaded_host_blas
def GKS(A, b, L, projection_dim=3, n_iter=50, regparam="gcv", x_true=None, **kwargs):
    delta, dp_stop = need_delta(regparam, kwargs, "gcv or a different stopping criterion.")
    if "dp_stop" in kwargs:  # the reference forwards it twice: golub_kahan(A, b, projection_dim, dp_stop, **kwargs) (:36)
        raise TypeError("golub_kahan() got multiple values for argument 'dp_stop'")
    A = as_operator(A)
    L = as_operator(L, A.device)
    dev = A.device
    bd = to_device_vector(b, dev)
    comm = kwargs.get("b200_comm")  # dist.FrameComm: frame-sharded dynamic CT (A block diagonal over the ranks' frames)
    bases = GKSBases(A, L, bd, projection_dim, n_iter, comm=comm)
    x_history = LazyHistory()
    lambda_history, residuals = [], []
    err = ErrorTracker(x_true, dev, comm=kwargs.get("b200_comm"))
    keep = kwargs.get("b200_history", "lazy")
    rp_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("b200_")}
    m, n = A.shape
    tm = torch.empty(m, dtype=K.F64, device=dev)
    tp = torch.empty(L.shape[0], dtype=K.F64, device=dev)
    ra = torch.empty(n, dtype=K.F64, device=dev)
    rb = torch.empty(n, dtype=K.F64, device=dev)
    xd = None
    lambdah = 0
    ii = -1
    for ii in range(n_iter):
        k = bases.k
        R_A, R_L, c, _, resid = factor_pair(bases, bd)  # la.qr(AV), la.qr(LV), Q_A.T@b       (GKS.py:54-58)
        lambdah = choose_lambda(regparam, R_A, R_L, c, resid, delta, rp_kwargs)
        lambda_history.append(lambdah)
        y = tikhonov_projected(R_A, R_L, c, lambdah)  #                                       (:74)
        yd = dev_scalar(y, dev)
        xd = K.basis_combine(bases.V, k, yd, out=xd)  # x = V @ y                               (:76)
        if keep != "none":
            x_history.append_lift(bases.V, k, y)
        err.add(xd)
        K.basis_combine(bases.AV, k, yd, out=tm)  # AV @ y
        K.vec_sub(tm, bd, out=tm)  # ra = AV@y - b                                               (:81)
        A.adjoint_dev(tm, out=ra)  # ra = A.T @ ra                                               (:82)
        K.basis_combine(bases.LV, k, yd, out=tp)  # rb = LV @ y                                  (:83)
        adjoint_L_weighted(L, tp, None, out=rb, comm=comm)  # rb = L.T @ rb                                 (:84)
        K.vec_axpy(float(lambdah), rb, ra, out=ra)  # r = ra + lambdah*rb                        (:85)
        expand(bases, ra, 3, residuals)  #                                                       (:86-96)
    info = {"xHistory": x_history, "regParam": lambdah, "regParam_history": lambda_history,
            "Residual": [float(v) for v in torch.stack(residuals)[:, 1].cpu().numpy()] if residuals else [], "its": ii}
    if x_true is not None:
        info["relError"] = err.values()
    return (host_column(xd), info)
