"""Hybrid GMRES on the GPU - signature and returns of trips/solvers/Hybrid_GMRES.py:23-87.

Arnoldi step on the device (`b200_reorth='mgs'` reproduces the reference's modified Gram-Schmidt sweep,
decompositions.py:216-218; `'cgs2'` is the bandwidth-optimal two-pass block variant), projected Tikhonov problem
and parameter rule on the host, lift x = V_k y on the device.

A must be square (:33-35).  For tomography pass the normal-equations operator `A.T @ A` with right-hand side
`A.T @ b` (operators compose: `CSROperator.T @ CSROperator` is a GPU operator).

Kept: at ii == 0 lambda = 0 but an iterate IS formed (:52-53,74-79, unlike Hybrid_LSQR); GCV is called with the
caller's kwargs only (:57).  Not reproduced: `la.norm(bhat - H@y)` at :80 subtracts a (k+1,1) column from a (k+1,)
vector and so reports the norm of a broadcast (k+1)x(k+1) matrix; info['relResidual'] here is ||bhat - H y||.
dp_stop=True is refused (same defect as in Hybrid_LSQR).
"""
import numpy as np
from scipy import linalg as la

from .. import kernels as K
from ..decompositions import ArnoldiState
from ..operators import as_operator, to_device_vector
from ..reg_param.discrepancy_principle import discrepancy_principle_projected
from ..reg_param.gcv import generalized_crossvalidation
from ..reg_param.l_curve import l_curve
from ._common import ErrorTracker, LazyHistory, dev_scalar, host_column, need_delta, tikhonov_projected, single_threaded_host_blas


@single_threaded_host_blas
def Hybrid_GMRES(A, b, n_iter, regparam="gcv", x_true=None, **kwargs):
    delta, dp_stop = need_delta(regparam, kwargs, "gcv or a different stopping criterion.")
    if dp_stop is not False:
        raise NotImplementedError("dp_stop=True is not supported (reference branch Hybrid_GMRES.py:61-67 is defective)")
    A = as_operator(A)
    n = A.shape[1]
    if A.shape[0] != n:
        raise Exception("Please check the size of the matrx A: it should be square in order to apply hybrid GMRES")
    dev = A.device
    bd = to_device_vector(b, dev)
    comm = kwargs.get("b200_comm")  # dist.BandComm / FrameComm: model space split over the ranks
    st = ArnoldiState(A, bd, n_iter, reorth=kwargs.get("b200_reorth", "mgs"), comm=comm)
    beta0 = float(st.beta0.cpu()[1])
    x_history = LazyHistory()
    lambda_history, residual_history = [], []
    err = ErrorTracker(x_true, dev, comm=comm if (comm is not None and comm.is_sharded("model")) else None)
    keep = kwargs.get("b200_history", "lazy")
    rp_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("b200_")}
    xd = None
    lambdah = 0
    ii = -1
    for ii in range(n_iter):
        st.step()  # (V, H) = arnoldi_update(A, V, H)                                     (Hybrid_GMRES.py:47)
        H = st.H_host()
        k = H.shape[1]
        bhat = np.zeros(k + 1)
        bhat[0] = beta0
        eye = np.eye(k)
        if ii == 0:
            lambdah = 0
        elif isinstance(regparam, str) and regparam == "gcv":
            Q_A, s, _ = la.svd(H, full_matrices=False)
            lambdah = generalized_crossvalidation(Q_A, np.diag(s), eye, bhat, **rp_kwargs)
        elif isinstance(regparam, str) and regparam == "dp":
            h = K.basis_dots(st.V, k + 1, bd)  # V^T b                              (discrepancy_principle.py:34)
            st._sum(h)
            explicit = rp_kwargs.get("explicitProj", False)
            resid = 0.0
            if explicit:
                res = K.new_pair(dev)
                K.basis_combine(st.V, k + 1, h, w=bd, sign=-1.0, norm_out=res)
                st._sync_norm(res)
                resid = float(res.cpu()[1])
            lambdah = discrepancy_principle_projected(H, None, h.cpu().numpy()[:k + 1], resid, delta,
                                                      rp_kwargs.get("eta", 1.01), explicit)
        elif isinstance(regparam, str) and regparam == "l_curve":
            Q_A, s, _ = la.svd(H, full_matrices=False)  #                                       (Hybrid_GMRES.py:67-71)
            lambdah = l_curve(np.diag(s), eye, Q_A.T @ bhat.reshape((-1, 1)))
        elif isinstance(regparam, str):
            raise NotImplementedError(f"regparam={regparam!r}: 'gcv', 'dp', 'l_curve' or a number")
        else:
            lambdah = regparam
        lambda_history.append(lambdah)
        y = tikhonov_projected(H, eye, bhat, lambdah)
        xd = K.basis_combine(st.V, k, dev_scalar(y, dev), out=xd)  # x = V[:, :-1] @ y          (:77)
        if keep != "none":
            x_history.append_lift(st.V, k, y)
        residual_history.append(float(la.norm(bhat.reshape(-1, 1) - H @ y)))
        err.add(xd)
    if xd is None:
        raise ValueError("Hybrid_GMRES needs n_iter >= 1")
    info = {"xHistory": x_history, "regParam": lambdah, "regParam_history": lambda_history,
            "relResidual": residual_history, "its": ii}
    if x_true is not None:
        info["relError"] = err.values()
    return (host_column(xd), info)
