"""Machinery shared by GKS and MMGKS: the generalised Krylov bases V, AV, LV on the device, the Gram-based
replacement of the per-iteration tall-skinny QRs, and the normal-equations residual + reorthogonalisation.

Reference loops: trips/solvers/GKS.py:36-105 and trips/solvers/MMGKS.py:37-137.

Replacing the QRs.  The reference factors AV*wf (m x k) and LV*wr (p x k) with Householder QR on the host at every
iteration and then only uses R_A, R_L, Q_A^T b, Q_A^T(wf*b) and ||wf*b - Q_A Q_A^T wf*b||.  Those are obtained here
from one double-double Gram pass per basis (tb200_weighted_gram) and a k x k double-double Cholesky on the host
(tb200_gram_factor_dd): R = chol(G) is the Householder R up to the signs of its rows, which neither the
regularisation-parameter rules nor y = argmin ||R_A y - Q_A^T b||^2 + lam ||R_L y||^2 can see.
"""
import numpy as np

from .. import kernels as K
from ..decompositions import golub_kahan_device
from ..kernels import Basis
from ..operators import CenteredDerivative2D, SpaceTimeDerivative
from ..reg_param.discrepancy_principle import discrepancy_principle_projected
from ..reg_param.gcv import generalized_crossvalidation
from ..reg_param.l_curve import l_curve


class GKSBases:
    def __init__(self, A, L, bd, projection_dim, n_iter, comm=None, dp_stop=False, gk_kwargs=None):
        self.A, self.L, self.comm = A, L, comm
        dev = bd.device
        kmax = projection_dim + n_iter + 1
        gk_kwargs = gk_kwargs or {}
        # golub_kahan(A, b, projection_dim, dp_stop, **kwargs)   (GKS.py:36, MMGKS.py:37); dp_stop may return fewer columns
        st = golub_kahan_device(A, bd, projection_dim, dp_stop=dp_stop, gk_eta=gk_kwargs.get("gk_eta", 1.001),
                                gk_delta=gk_kwargs.get("gk_delta", 0.001), comm=comm)
        projection_dim = st.k
        self.V = Basis(A.shape[1], kmax, dev)
        self.AV = Basis(A.shape[0], kmax, dev)
        self.LV = Basis(L.shape[0], kmax, dev)
        self.gram_AV, self.gram_LV = K.IncrementalGram(), K.IncrementalGram()  # for the passes without row weights
        for j in range(projection_dim):  # AV = A@V ; LV = L@V                            (GKS.py:37-38)
            self.V.next_col().copy_(st.V.col(j))
            self.V.push()
            self.append_images()
        del st

    @property
    def k(self):
        return self.V.k

    def append_images(self):
        """AV, LV gain the images of the newest column of V  (GKS.py:92-96, MMGKS.py:124-128)."""
        vn = self.V.col(self.V.k - 1)
        self.A.apply_dev(vn, out=self.AV.next_col())
        self.AV.push()
        apply_L(self.L, vn, self.comm, out=self.LV.next_col())
        self.LV.push()


def apply_L(L, x, comm=None, out=None, wout=None, eps=0.0, expo=0.0):
    """u = L x (optionally with fused IRLS weights); frame-sharded difference operators fetch their one-frame halo."""
    if isinstance(L, SpaceTimeDerivative):
        x_next = None
        if comm is not None and comm.frames:
            N = L.nx * L.ny
            x_next = comm.halo_from_next(x[:N])
        return L.apply_dev(x, out=out, wout=wout, eps=eps, expo=expo, x_next=x_next)
    if wout is not None:
        raise TypeError("fused weights need a matrix-free difference operator")
    return L.apply_dev(x, out=out)


def apply_L_with_weights(L, x, eps, expo, comm=None):
    """u = L x and wr = (u^2 + eps^2)^expo, in one pass when L is a matrix-free difference operator."""
    import torch

    if isinstance(L, SpaceTimeDerivative):
        wr = torch.empty(L.shape[0], dtype=K.F64, device=x.device)
        u = apply_L(L, x, comm, wout=wr, eps=eps, expo=expo)
        return u, wr
    u = L.apply_dev(x)
    return u, K.irls_weights(u, eps, expo)


def adjoint_L_weighted(L, r, w, out=None, comm=None):
    """L^T (w . r)   (w None: L^T r).  Frame-sharded: the temporal rows of my last frame also act on the next rank."""
    if isinstance(L, SpaceTimeDerivative):
        rt_prev = None
        if comm is not None and comm.frames:
            N = L.nx * L.ny
            if L.has_next:
                lo = L.shape[0] - N  # temporal rows of my last frame: the final N rows
                blk = r[lo:lo + N] if w is None else K.vec_mul(w[lo:lo + N].contiguous(), r[lo:lo + N].contiguous())
            else:
                blk = r[:N]  # nothing to hand on (last rank); participates for the receive only
            rt_prev = comm.halo_from_prev(blk.contiguous())
        return L.adjoint_dev(r, out=out, w=w, rt_prev=rt_prev, wt_prev=None)
    if w is None:
        return L.adjoint_dev(r, out=out)
    if isinstance(L, CenteredDerivative2D):
        return L.adjoint_dev(r, out=out, w=w)
    return L.adjoint_dev(K.vec_mul(w, r), out=out)


def _gram_comm(comm, space):
    """The communicator to sum a Gram matrix with, or None when the basis' space is replicated."""
    return comm if (comm is not None and comm.is_sharded(space)) else None


def householder_factor(M, Z):
    """The reference's route, on the host: (R, C, resid2) of la.qr(M) with the rows of R signed so that its diagonal is
    non-negative (the convention of the Gram / Cholesky route), C = Q^T Z, resid2[e] = ||Z[:, e] - Q Q^T Z[:, e]||^2.
    M: m x k, Z: m x ne NumPy arrays (already weighted).  MMGKS.py:58-59,94-95; GKS.py:54-58."""
    import scipy.linalg as la

    Q, R = la.qr(M, mode="economic")
    sg = np.where(np.diag(R) < 0, -1.0, 1.0)
    Q, R = Q * sg[None, :], R * sg[:, None]
    C = Q.T @ Z
    resid2 = np.sum((Z - Q @ C) ** 2, axis=0)
    return np.triu(R), C, resid2


def _factor(B, k, w, extras, flags, inc, comm):
    """R, Q^T extras, residuals of [diag(w) B[:, :k] | extras]: double-double Gram pass + Cholesky; if the Gram matrix is
    not numerically positive definite (a basis vector in the null space of L, extreme weights - situations the
    reference's Householder QR walks through), the columns are brought to the host and factored the reference's way."""
    if w is None and inc is not None:
        Ghi, Glo = inc.update(B, k, extras=extras, extra_weighted=flags, comm=comm)
    else:
        Ghi, Glo = K.weighted_gram(B, k, w, extras=extras, extra_weighted=flags, comm=comm)
    try:
        return K.gram_factor(Ghi, Glo, k)
    except np.linalg.LinAlgError:
        if comm is not None:  # row-sharded basis: no rank holds the columns
            raise
        import warnings

        warnings.warn("Gram matrix of the Krylov basis is not positive definite: falling back to a Householder QR on "
                      "the host for this iteration", RuntimeWarning, stacklevel=3)
        wh = None if w is None else w.cpu().numpy()[:, None]
        M = B.to_numpy(k)
        Z = np.stack([e.cpu().numpy() for e in extras], axis=1) if extras else np.zeros((M.shape[0], 0))
        if wh is not None:
            M = M * wh
            for e, f in enumerate(flags):
                if f:
                    Z[:, e] = Z[:, e] * wh[:, 0]
        return householder_factor(M, Z)


def factor_pair(bases, bd, wf=None, wr=None):
    """R_A, R_L, c_plain = Q_A^T b, c_w = Q_A^T (wf*b), resid_w = ||wf*b - Q_A Q_A^T wf*b||."""
    k = bases.k
    comm = bases.comm
    if wf is None:
        R_A, C, res2 = _factor(bases.AV, k, None, (bd,), (0,), bases.gram_AV, _gram_comm(comm, "data"))
        c_plain = c_w = C[:, 0:1]
        resid_w = float(np.sqrt(res2[0]))
    else:
        R_A, C, res2 = _factor(bases.AV, k, wf, (bd, bd), (0, 1), None, _gram_comm(comm, "data"))
        c_plain, c_w = C[:, 0:1], C[:, 1:2]
        resid_w = float(np.sqrt(res2[1]))
    R_L, _, _ = _factor(bases.LV, k, wr, (), (), bases.gram_LV, _gram_comm(comm, "reg"))
    return R_A, R_L, c_plain, c_w, resid_w


def choose_lambda(regparam, R_A, R_L, c_w, resid_w, delta, rp_kwargs, c_plain=None):
    """GKS.py:60-69 / MMGKS.py:96-103 with the long-vector products already projected."""
    if isinstance(regparam, str) and regparam == "l_curve":  # l_curve(R_A, R_L, Q_A.T @ b): the UNWEIGHTED b (MMGKS.py:101)
        return l_curve(R_A, R_L, c_w if c_plain is None else c_plain)
    if isinstance(regparam, str) and regparam == "gcv":
        return generalized_crossvalidation(None, R_A, R_L, c_w, **rp_kwargs)
    if isinstance(regparam, str) and regparam == "dp":
        return discrepancy_principle_projected(R_A, R_L, c_w, resid_w, delta, rp_kwargs.get("eta", 1.01),
                                               rp_kwargs.get("explicitProj", False))
    if isinstance(regparam, str):
        raise NotImplementedError(f"regparam={regparam!r}: 'gcv', 'dp', 'l_curve' or a number")
    return regparam


def expand(bases, r, n_reorth, residual_history):
    """r -= V (V^T r) n_reorth times, record ||r||, append r/||r|| and its images  (GKS.py:86-96, MMGKS.py:119-129)."""
    k = bases.k
    nrm = K.new_pair(r.device)
    comm = bases.comm
    for it in range(n_reorth):
        h = K.basis_dots(bases.V, k, r)
        if comm is not None:
            comm.sum_(h, "model")
        K.basis_combine(bases.V, k, h, w=r, sign=-1.0, out=r, norm_out=nrm if it == n_reorth - 1 else None)
    if n_reorth == 0:
        K.vec_norm2(r, out=nrm)
    if comm is not None:
        comm.sync_norm_(nrm, "model")
    K.vec_div(r, nrm[1:2], out=bases.V.next_col())
    bases.V.push()
    bases.append_images()
    residual_history.append(nrm)  # device pair; converted once at the end
