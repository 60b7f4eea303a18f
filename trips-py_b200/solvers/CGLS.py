"""CGLS on the GPU - same signature, returns and stopping rule as trips/solvers/CGLS.py:16-86.

Per iteration: w = A p (fused ||w||^2), x += beta p, r -= beta w, t = A^T r (fused ||t||^2), p = t + (gamma/gamma_old) p.
The scalars beta, gamma are formed on the host in IEEE double exactly as the reference does
(`np.linalg.norm(w)**2`: the square of the rounded norm, not the sum of squares), because the stopping test
`norm_t <= norms_t0*tol or norm_x*tol >= 1` (:75) needs them on the host every iteration anyway.

Reference quirks kept: relResidual is ||x - x_old||/||x|| (:76); relError divides by ||x|| (:79).
Reference bug not reproduced: `np.eps` (:63, AttributeError when ||A p|| == 0) - machine epsilon is used.
"""
import numpy as np
import torch

from .. import kernels as K
from ..decompositions import apply_fused
from ..kernels import F64
from ..operators import as_operator, to_device_vector
from ._common import ErrorTracker, LazyHistory, host_column


def CGLS(A, b, x0, max_iter, tol, x_true=None, **kwargs):
    A = as_operator(A)
    dev = A.device
    m, n = A.shape
    bd = to_device_vector(b, dev)
    x = to_device_vector(x0, dev).clone()
    keep = kwargs.get("b200_history", "lazy")
    comm = kwargs.get("b200_comm")  # dist.RowComm / dist.FrameComm: norms over split spaces are summed across ranks

    def sync(pair, space):
        if comm is not None:
            comm.sync_norm_(pair, space)

    sc = torch.zeros(8, dtype=F64, device=dev)  # [0:2] ||w||, [2:4] ||t||, [4:6] ||x||, [6:8] ||x - x_old||
    r = K.vec_sub(bd, A.apply_dev(x))  # r = b - A x                                   (CGLS.py:45)
    t = torch.empty(n, dtype=F64, device=dev)
    apply_fused(A, r, t, adjoint=True, norm_out=sc[2:4])  # t = A^T r                              (:46)
    sync(sc[2:4], "model")
    p = t.clone()
    K.vec_norm2(x, out=sc[4:6])
    sync(sc[4:6], "model")
    host = sc.cpu().numpy()
    norms_t0, normx = float(host[3]), float(host[5])
    gamma, xmax = norms_t0 ** 2, normx
    k, check = 0, 0
    x_history = LazyHistory()
    rel_residual, norms_x = [], []
    err = ErrorTracker(x_true, dev, comm=comm)
    w = torch.empty(m, dtype=F64, device=dev)
    x_old = torch.empty_like(x)
    norm_x = normx
    while (k < max_iter) and (check == 0):
        x_old.copy_(x)
        k += 1
        apply_fused(A, p, w, norm_out=sc[0:2])  # w = A p                               (:60)
        sync(sc[0:2], "data")
        delta = float(sc[1].item()) ** 2  # np.linalg.norm(w)**2                          (:61)
        if delta == 0:
            delta = np.finfo(np.float64).eps
        beta = gamma / delta
        K.vec_axpy(beta, p, x, out=x)  # x = x + beta*p                                   (:65)
        K.vec_axpy(beta, w, r, out=r, sign=-1.0)  # r = r - beta*w                        (:67)
        apply_fused(A, r, t, adjoint=True, norm_out=sc[2:4])  # t = A^T r                            (:68)
        K.vec_norm2(x, out=sc[4:6])
        K.vec_diffnorm2(x, x_old, out=sc[6:8])
        for q in (2, 4, 6):
            sync(sc[q:q + 2], "model")
        host = sc.cpu().numpy()
        gamma_old = gamma
        norm_t = float(host[3])
        gamma = norm_t ** 2
        K.vec_axpy(gamma / gamma_old, p, t, out=p)  # p = t + (gamma/gamma_old)*p          (:72)
        norm_x = float(host[5])
        xmax = max(xmax, norm_x)
        check = (norm_t <= norms_t0 * tol) or (norm_x * tol >= 1)
        rel_residual.append(float(host[7]) / norm_x)
        norms_x.append(norm_x)
        if keep != "none":
            x_history.append_device(x.clone())
        err.add(x)
    info = {"xHistory": x_history, "regParam": [], "relResidual": rel_residual, "its": k}
    if x_true is not None:
        info["relError"] = err.values(denominators=norms_x)
    return (host_column(x), info)
