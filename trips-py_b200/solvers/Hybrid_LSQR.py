"""Hybrid LSQR on the GPU - signature and returns of trips/solvers/Hybrid_LSQR.py:25-114.

Each iteration: one device Golub-Kahan step (2 fused SpMVs + 2 scalings, no host traffic), one small D2H of the
bidiagonal entries, the (k+1) x k Tikhonov problem and its parameter rule on the host in NumPy (as in the
reference), and the lift x = V y as one stream of the device basis.

Reference behaviour kept because it changes results: at ii == 0 lambda = 0 and NO iterate is formed (:77-78), so
n_iter = 1 fails like the reference (UnboundLocalError there, ValueError here); GCV is called with
variant='modified', fullsize=m (:84) while its numerator stays 'standard' (gcv.py:94); info['its'] is the last
loop index.  Refused: dp_stop=True (the reference path at :88-94 raises a shape error before it can stop).
"""
import numpy as np
import torch
from scipy import linalg as la

from .. import kernels as K
from ..decompositions import GKState
from ..operators import as_operator, to_device_vector
from ..reg_param.discrepancy_principle import discrepancy_principle_projected
from ..reg_param.gcv import generalized_crossvalidation
from ..reg_param.l_curve import l_curve
from ._common import ErrorTracker, LazyHistory, dev_scalar, host_column, need_delta, tikhonov_projected, single_threaded_host_blas


@single_threaded_host_blas
def Hybrid_LSQR(A, b, n_iter=100, regparam="gcv", x_true=None, **kwargs):
    delta, dp_stop = need_delta(regparam, kwargs, "gcv.")
    if dp_stop is not False:
        raise NotImplementedError("dp_stop=True is not supported: the reference implementation of that branch "
                                  "(Hybrid_LSQR.py:88-94) fails with a shape error")
    A = as_operator(A)
    dev = A.device
    m, n = A.shape
    bd = to_device_vector(b, dev)
    comm = kwargs.get("b200_comm")  # dist.RowComm (static CT, rows by angle) or dist.FrameComm (block-diagonal A)
    m_total = m if comm is None else comm.total(m, "data")
    # (a band-sharded matrix-free operator brings its own fused recurrence: dist.ShardedGKState)
    st = A.gk_state(bd, n_iter) if hasattr(A, "gk_state") else GKState(A, bd, n_iter, comm=comm)
    x_history = LazyHistory()
    lambda_history, residual_history = [], []
    err = ErrorTracker(x_true, dev, comm=comm)
    keep = kwargs.get("b200_history", "lazy")
    rp_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("b200_")}
    xd = None
    lambdah = 0
    ii = -1
    use_dp = isinstance(regparam, str) and regparam == "dp"
    explicit = rp_kwargs.get("explicitProj", False)
    # Pipelined by one iteration: the projected problem of iteration ii (host: SVD / root finding / least squares on
    # (k+1) x k matrices, as in the reference) is solved WHILE the device runs Golub-Kahan step ii + 1, which does not
    # depend on it.  What the host needs from iteration ii - the bidiagonal entries and, for 'dp', U^T b - is reduced on the
    # device right after step ii and copied to pinned memory asynchronously; the lift x = V y follows step ii + 1 on the stream
    # (it only reads columns < k of V).  Same values as the sequential loop, in the same order.
    pending = None
    pinned = torch.empty((2, 3 * n_iter + 8), dtype=torch.float64).pin_memory()  # double buffer of the per-iteration D2H
    flip = [0]

    def snapshot(k):
        """Enqueue the device-side reductions of the current iterate and their async D2H; returns what finish() needs."""
        parts = [st.beta0[1:2], st.alpha[:k, 1], st.beta[:k, 1]]
        resid_pair = None
        if use_dp:
            h = K.basis_dots(st.U, k + 1, bd)  # U^T b                                   (discrepancy_principle.py:34)
            if comm is not None:
                comm.sum_(h, "data")
            parts.append(h[:k + 1])
            if explicit:  # ||b - U U^T b|| is only consulted by the explicitProj variant (discrepancy_principle.py:69,82)
                resid_pair = K.new_pair(dev)
                K.basis_combine(st.U, k + 1, h, w=bd, sign=-1.0, norm_out=resid_pair)
                if comm is not None:
                    comm.sync_norm_(resid_pair, "data")
                parts.append(resid_pair[1:2])
        packed = torch.cat(parts)
        host = pinned[flip[0]][:packed.numel()]
        flip[0] ^= 1
        host.copy_(packed, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return k, host, ev

    def finish(snap):
        nonlocal xd, lambdah
        k, host, ev = snap
        ev.synchronize()
        v = host.numpy()
        beta0, al, be = v[0], v[1:1 + k], v[1 + k:1 + 2 * k]
        B = np.zeros((k + 1, k))
        B[np.arange(k), np.arange(k)] = al
        B[np.arange(1, k + 1), np.arange(k)] = be
        bhat = np.zeros(k + 1)
        bhat[0] = beta0
        eye = np.eye(k)
        if isinstance(regparam, str) and regparam == "gcv":
            Q_A, s, _ = la.svd(B, full_matrices=False)
            lambdah = generalized_crossvalidation(Q_A, np.diag(s), eye, bhat, variant="modified", fullsize=m_total, **rp_kwargs)
        elif use_dp:
            # discrepancy_principle(U, B, L, b): U^T b and ||b - U U^T b|| come from the device basis
            h = v[1 + 2 * k:2 + 3 * k]
            resid = float(v[2 + 3 * k]) if explicit else 0.0
            lambdah = discrepancy_principle_projected(B, None, h, resid, delta, rp_kwargs.get("eta", 1.01), explicit)
        elif isinstance(regparam, str) and regparam == "l_curve":
            Q_A, s, _ = la.svd(B, full_matrices=False)  #                                       (Hybrid_LSQR.py:94-98)
            lambdah = l_curve(np.diag(s), eye, Q_A.T @ bhat.reshape((-1, 1)))
        elif isinstance(regparam, str):
            raise NotImplementedError(f"regparam={regparam!r}: 'gcv', 'dp', 'l_curve' or a number")
        else:
            lambdah = regparam
        lambda_history.append(lambdah)
        y = tikhonov_projected(B, eye, bhat, lambdah)
        xd = K.basis_combine(st.V, k, dev_scalar(y, dev), out=xd)  # x = V @ y                (:105)
        if keep != "none":
            x_history.append_lift(st.V, k, y)
        err.add(xd)

    if isinstance(regparam, str) and regparam not in ("gcv", "dp", "l_curve"):
        raise NotImplementedError(f"regparam={regparam!r}: 'gcv', 'dp', 'l_curve' or a number")
    for ii in range(n_iter):
        st.step()  # (U, B, V) = golub_kahan_update(A, U, B, V)                          (Hybrid_LSQR.py:74)
        snap = snapshot(st.k) if ii > 0 else None  # ii == 0: lambda = 0 and no iterate   (Hybrid_LSQR.py:77-78)
        if pending is not None:
            finish(pending)  # host work of iteration ii - 1, overlapped with step ii on the device
        pending = snap
        if ii == 0:
            lambdah = 0
    if pending is not None:
        finish(pending)
    if xd is None:
        raise ValueError("Hybrid_LSQR forms no iterate when n_iter < 2 (the reference raises UnboundLocalError)")
    info = {"xHistory": x_history, "regParam": lambdah, "regParam_history": lambda_history,
            "relResidual": residual_history, "its": ii}
    if x_true is not None:
        info["relError"] = err.values()
    return (host_column(xd), info)
