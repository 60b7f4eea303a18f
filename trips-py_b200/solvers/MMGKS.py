"""Majorisation-minimisation generalised Krylov subspace method (l_p - l_q) on the GPU.
Signature and returns of trips/solvers/MMGKS.py:28-137.

Per iteration (reference line numbers):
  v = A x - b; wf = (v^2+eps^2)^(p/2-1)                    :56-57   (elided when pnorm == 2: wf == 1 exactly)
  factors of AV*wf                                         :58-59   one weighted Gram pass + k x k Cholesky
  u = L x; wr = (u^2+eps^2)^(q/2-1) | isoTV weights        :60-93   fused into the difference stencil
  factors of LV*wr                                         :94-95   one weighted Gram pass
  lambda (gcv | dp on wf*b | fixed)                        :96-103  host, k x k
  y = lstsq([R_A; sqrt(lambda) R_L], [Q_A^T b; 0]); x = V y :106-107
  r = A^T(wf*(AV y - b)) + lambda L^T(wr*(LV y))           :111-118 weights fused into the stencil adjoint
  r -= V(V^T r) twice; vn = r/||r||; append V, AV, LV      :119-129

The weights follow the CODE, not the paper: both `AV*wf` and `LV*wr` use the full IRLS weight with exponent
(p/2-1), (q/2-1) (:57,93), except the isoTV branch which uses (q-2)/4 (:75-77).
isoTV: the reference builds its own spatial gradient with pylops' float32 centred differences
(operators_old.py:35-45); here the gradient is the fp64 statement of the same centred stencil and `L`
must be that operator (`CenteredDerivative2D`), single frame.
GS (group sparsity, :45-52,79-91): as in the reference the caller's L is replaced by kron(I_nt, Ls) with Ls the 2-D
differences of operators_old.py:75-85 (applied through the CSR kernels), and the weights are
(||(Ls X)_{i,:}||^2 + e^2)^(q/2-1) with X the C-order reshape of x - both exactly as written there.
"""
import numpy as np
import torch
from scipy import sparse

from .. import kernels as K
from ..operators import as_operator, to_device_vector
from ._common import ErrorTracker, LazyHistory, dev_scalar, host_column, tikhonov_projected, single_threaded_host_blas
from ._gks_core import GKSBases, adjoint_L_weighted, apply_L_with_weights, choose_lambda, expand, factor_pair


def _old_first_derivative_1d(n):
    """Rows 0..n-2 of I - subdiag(1): (L x)_0 = x_0, (L x)_i = x_i - x_{i-1}   (operators_old.py:66-72)."""
    D = sparse.spdiags(data=np.ones(n - 1), diags=-1, m=n, n=n)
    return (sparse.identity(n, format="csr") - D).tocsr()[0:-1, :]


def _old_first_derivative_2d(nx, ny):
    """operators_old.py:75-85."""
    return sparse.vstack((sparse.kron(sparse.identity(nx), _old_first_derivative_1d(nx)),
                          sparse.kron(_old_first_derivative_1d(ny), sparse.identity(ny)))).tocsr()


def _group_sparsity_weights(Ls, xd, n_space, qnorm):
    """wr_i = (||(Ls X)_{i,:}||^2 + e^2)^(q/2-1), X = x reshaped (n_space, nt) in C order, repeated for every frame
    (MMGKS.py:79-91, as written: C-order reshape of the frame-major x and exp(2) as the smoothing constant)."""
    nt = xd.numel() // n_space
    acc = None
    for t in range(nt):
        col = xd if nt == 1 else xd[t::nt].contiguous()
        d = Ls.apply_dev(col)
        sq = K.vec_mul(d, d)
        acc = sq if acc is None else K.vec_add(acc, sq, out=acc)
    # (v^2 + eps^2)^expo with v = sqrt(acc) (the reference takes the norm, then squares it) and eps = e
    wr = K.irls_weights(torch.sqrt(acc), float(np.exp(1.0)), qnorm / 2 - 1)
    return wr if nt == 1 else wr.repeat(nt)


@single_threaded_host_blas
def MMGKS(A, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=5, regparam="gcv", x_true=None, **kwargs):
    # unlike the other drivers the reference does not validate delta up front (:30-36); the discrepancy-principle
    # routine raises the same Exception when it is missing (discrepancy_principle.py:21-23)
    delta = kwargs["delta"] if ("delta" in kwargs) else None
    if "dp_stop" in kwargs:  # the reference forwards it twice: golub_kahan(A, b, projection_dim, dp_stop, **kwargs) (:37)
        raise TypeError("golub_kahan() got multiple values for argument 'dp_stop'")
    isoTV_option = kwargs["isoTV"] if ("isoTV" in kwargs) else False
    GS_option = kwargs["GS"] if ("GS" in kwargs) else False
    epsilon = kwargs["epsilon"] if ("epsilon" in kwargs) else 0.1
    prob_dims = kwargs["prob_dims"] if ("prob_dims" in kwargs) else False
    iso = isoTV_option in ["isoTV", "ISOTV", "IsoTV"]
    gs = GS_option in ["GS", "gs", "Gs"]
    if gs and prob_dims is False:
        raise TypeError("For Isotropic Group Sparsity you must enter the dimension of the dynamic problem. (x_mmgks, info_mmgks) = MMGKS(A, data_vec, L, pnorm=2, qnorm=1, projection_dim=2, n_iter =3, regparam = 'gcv', x_true = None, GS = 'GS', prob_dims = (nx,ny, nt))")
    if iso and prob_dims is False:
        raise TypeError("For Isotropic TV you must enter the dimension of the dynamic problem! Example: (x_mmgks, info_mmgks) = MMGKS(A, data_vec, L, pnorm=2, qnorm=1, projection_dim=2, n_iter =3, regparam = 'gcv', x_true = None, isoTV = 'isoTV', prob_dims = (nx,ny, nt))")

    A = as_operator(A)
    Ls = None
    if gs:
        # the reference REPLACES the caller's L by kron(I_nt, Ls), Ls = the 2-D differences of operators_old.py:75-85
        # (row 0 of each 1-D block is x_0, row i is x_i - x_{i-1})                                   (MMGKS.py:45-52)
        gnx, gny, gnt = (int(v) for v in prob_dims)
        Ls_host = _old_first_derivative_2d(gnx, gny)
        Ls = as_operator(Ls_host, A.device)
        L = Ls if gnt == 1 else sparse.kron(sparse.identity(gnt), Ls_host).tocsr()
    L = as_operator(L, A.device)
    dev = A.device
    m, n = A.shape
    p = L.shape[0]
    bd = to_device_vector(b, dev)
    if iso:
        from ..operators import CenteredDerivative2D

        if not isinstance(L, CenteredDerivative2D) or p != 2 * n:
            raise TypeError("isoTV needs L = CenteredDerivative2D(nx, ny) (the fp64 statement of the reference's "
                            "first_derivative_operator_2d), single frame")
    comm = kwargs.get("b200_comm")  # dist.FrameComm: frame-sharded dynamic CT (A block diagonal over the ranks' frames)
    bases = GKSBases(A, L, bd, projection_dim, n_iter, comm=comm)
    x_history = LazyHistory()
    lambda_history, residuals = [], []
    err = ErrorTracker(x_true, dev, comm=kwargs.get("b200_comm"))
    keep = kwargs.get("b200_history", "lazy")
    rp_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("b200_")}
    tm = torch.empty(m, dtype=K.F64, device=dev)
    tp = torch.empty(p, dtype=K.F64, device=dev)
    ra = torch.empty(n, dtype=K.F64, device=dev)
    rb = torch.empty(n, dtype=K.F64, device=dev)
    xd = A.adjoint_dev(bd)  # x = A.T @ b                                                      (MMGKS.py:43)
    lambdah = 0
    ii = -1
    for ii in range(n_iter):
        k = bases.k
        if pnorm == 2:
            wf = None  # (v^2+eps^2)^0 == 1 exactly: the A@x of :56 cannot influence anything
        else:
            A.apply_dev(xd, out=tm)
            K.vec_sub(tm, bd, out=tm)  # v = A@x - b                                             (:56)
            wf = K.irls_weights(tm, epsilon, pnorm / 2 - 1)  #                                   (:57)
        if iso:
            wr = L.iso_weights(xd, epsilon, (qnorm - 2) / 4)  #                                  (:64-78)
        elif gs:
            wr = _group_sparsity_weights(Ls, xd, gnx * gny, qnorm)  #                            (:79-91)
        else:
            _, wr = apply_L_with_weights(L, xd, epsilon, qnorm / 2 - 1, comm=comm)  # u = L@x; wr           (:60,93)
        R_A, R_L, c_plain, c_w, resid_w = factor_pair(bases, bd, wf=wf, wr=wr)  #               (:58-59,94-95)
        lambdah = choose_lambda(regparam, R_A, R_L, c_w, resid_w, delta, rp_kwargs, c_plain=c_plain)  # (:96-103)
        lambda_history.append(lambdah)
        y = tikhonov_projected(R_A, R_L, c_plain, lambdah)  #                                    (:106)
        yd = dev_scalar(y, dev)
        xd = K.basis_combine(bases.V, k, yd, out=xd)  # x = V @ y                                (:107)
        if keep != "none":
            x_history.append_lift(bases.V, k, y)
        err.add(xd)
        if ii >= R_L.shape[0]:
            break
        K.basis_combine(bases.AV, k, yd, out=tm)
        if wf is None:
            K.vec_sub(tm, bd, out=tm)
        else:
            K.vec_wsub(wf, tm, bd, out=tm)  # ra = wf*(AV@y - b)                                  (:111)
        A.adjoint_dev(tm, out=ra)  # ra = A.T @ ra                                               (:115)
        K.basis_combine(bases.LV, k, yd, out=tp)
        adjoint_L_weighted(L, tp, wr, out=rb, comm=comm)  # rb = L.T @ (wr*(LV@y))                          (:113-117)
        K.vec_axpy(float(lambdah), rb, ra, out=ra)  # r = ra + lambdah*rb                        (:118)
        expand(bases, ra, 2, residuals)  #                                                       (:119-129)
    info = {"xHistory": x_history, "regParam": lambdah, "regParam_history": lambda_history,
            "Residual": [float(v) for v in torch.stack(residuals)[:, 1].cpu().numpy()] if residuals else [], "its": ii}
    if x_true is not None:
        info["relError"] = err.values()
    return (host_column(xd), info)
