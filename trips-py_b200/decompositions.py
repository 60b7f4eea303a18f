"""Krylov cores: Golub-Kahan bidiagonalisation and Arnoldi, device resident.

Reference: trips/utilities/decompositions.py
  golub_kahan        :118-205   batch GK used to seed GKS / MMGKS
  arnoldi_update     :207-228   one Arnoldi step, modified Gram-Schmidt, no reorthogonalisation
  golub_kahan_update :230-255   one GK step, no reorthogonalisation
(`arnoldi` :20-116 is defective in the reference - H[i,i] is never set - and is not used by the five solvers
this package covers; it is deliberately not provided.)

Two layers:
  * GKState / ArnoldiState keep U, V (and the scalars alpha, beta, h) on the GPU, pre-allocated, and run one step
    as 2 SpMVs with fused recurrence + norm epilogues and 2 scaling passes; nothing returns to the host unless the
    caller asks for the projected matrix.  This is what the solvers and the headline benchmark use.
  * golub_kahan_update(A, U, S, V), golub_kahan(A, b, n_iter), arnoldi_update(A, V, H) have the reference's
    signatures and return NumPy arrays of the same shapes, with the arithmetic on the GPU (host vectors in, host
    vectors out).  Host bases are grown in place inside pinned, column-contiguous capacity buffers instead of
    being re-copied with np.hstack at every call.
"""
import numpy as np
import torch

from . import kernels as K
from .kernels import F64, Basis
from .operators import as_operator, to_device_vector


def apply_fused(op, x, out, adjoint=False, coef=None, z=None, norm_out=None):
    """out = op x - coef*z with optional fused norm; falls back to separate vector kernels for operators
    without a fused epilogue (stencil operators)."""
    fn = op.adjoint_dev if adjoint else op.apply_dev
    if op.fused:
        return fn(x, out=out, coef=coef, z=z, norm_out=norm_out)
    fn(x, out=out)
    if z is not None:
        K.vec_axpy(coef, z, out, out=out, norm_out=norm_out, sign=-1.0)
    elif norm_out is not None:
        K.vec_norm2(out, out=norm_out)
    return out


class GKState:
    """Golub-Kahan bidiagonalisation state on the device (one `step()` == one golub_kahan_update).

    `comm` selects a sharded mode (dist.FrameComm: block-diagonal operators, every vector local; dist.RowComm: rows of
    A by angle, u-space local and v-space replicated): only the squared norms over split spaces are all-reduced."""

    def __init__(self, A, b_dev, kmax, comm=None):
        self.A, self.comm = A, comm
        m, n = A.shape
        dev = b_dev.device
        self.U = Basis(m, kmax + 1, dev)
        self.V = Basis(n, max(kmax, 1), dev)
        self._alloc_scalars(kmax)
        # U[:,0] = b / ||b||   (Hybrid_LSQR.py:64-65, decompositions.py:159)
        self.beta0 = torch.zeros(2, dtype=F64, device=dev)
        K.vec_norm2(b_dev, out=self.beta0)
        self._sync(self.beta0, "data")
        K.vec_div(b_dev, self.beta0[1:2], out=self.U.next_col())
        self.U.push()

    def _sync(self, pair, space):
        if self.comm is not None:
            self.comm.sync_norm_(pair, space)

    def _alloc_scalars(self, kmax):
        dev = self.U.data.device
        self.alpha = torch.zeros((kmax + 1, 2), dtype=F64, device=dev)  # [:,0] = sum of squares, [:,1] = norm
        self.beta = torch.zeros((kmax + 1, 2), dtype=F64, device=dev)

    @property
    def k(self):
        return self.V.k

    def step(self):
        k = self.V.k
        if k + 1 >= self.alpha.shape[0]:
            old_a, old_b = self.alpha, self.beta
            self._alloc_scalars(2 * (k + 1))
            self.alpha[:k + 1].copy_(old_a[:k + 1])
            self.beta[:k + 1].copy_(old_b[:k + 1])
        u_k = self.U.col(k)
        v = self.V.next_col()
        A = self.A
        if self.comm is None and getattr(A, "A_sell", None) is not None and A.order == "sequential" \
                and A.A_sell.vals.dtype == F64:
            # the whole step behind one C-ABI call (tb200_gk_step_sell_f64)
            self.V.push()
            u = self.U.next_col()
            K.gk_step_sell(A.A_sell, A.AT_sell, u_k, self.V.col(k - 1) if k else None,
                           self.beta[k - 1, 1:2] if k else None, v, u, self.alpha[k], self.beta[k])
            self.U.push()
            return
        if self.comm is None and getattr(A, "projector", None) is not None:
            # matrix-free CT operator: the whole step behind tb200_gk_step_ct_f64
            self.V.push()
            u = self.U.next_col()
            A.projector.gk_step(u_k, self.V.col(k - 1) if k else None, self.beta[k - 1, 1:2] if k else None, v, u,
                                self.alpha[k], self.beta[k])
            self.U.push()
            return
        # v = A^T u_k - beta_{k-1} v_{k-1} ; alpha = ||v|| ; v /= alpha      (decompositions.py:234-239)
        if k == 0:
            apply_fused(self.A, u_k, v, adjoint=True, norm_out=self.alpha[k])
        else:
            apply_fused(self.A, u_k, v, adjoint=True, coef=self.beta[k - 1, 1:2], z=self.V.col(k - 1),
                        norm_out=self.alpha[k])
        self._sync(self.alpha[k], "model")
        K.vec_div(v, self.alpha[k, 1:2], out=v)
        self.V.push()
        # u = A v - alpha u_k ; beta = ||u|| ; u /= beta                      (decompositions.py:240-242)
        u = self.U.next_col()
        apply_fused(self.A, v, u, coef=self.alpha[k, 1:2], z=u_k, norm_out=self.beta[k])
        self._sync(self.beta[k], "data")
        K.vec_div(u, self.beta[k, 1:2], out=u)
        self.U.push()

    def scalars_host(self):
        """(beta0, alphas[k], betas[k]) as NumPy - the one small D2H a hybrid solver needs per iteration."""
        k = self.V.k
        packed = torch.cat((self.beta0[1:2], self.alpha[:k, 1], self.beta[:k, 1])).cpu().numpy()
        return packed[0], packed[1:1 + k], packed[1 + k:]

    def B_host(self):
        """The (k+1) x k lower-bidiagonal projected matrix (decompositions.py:198-200, 248-254)."""
        _, al, be = self.scalars_host()
        k = al.size
        B = np.zeros((k + 1, k))
        B[np.arange(k), np.arange(k)] = al
        B[np.arange(1, k + 1), np.arange(k)] = be
        return B


class ArnoldiState:
    """Arnoldi state on the device.  reorth='mgs' follows the reference column by column
    (decompositions.py:216-218); reorth='cgs2' is the north-star design: two block projections
    h = V^T w, w -= V h, each a single stream of the basis (h accumulates both passes)."""

    def __init__(self, A, b_dev, kmax, reorth="mgs", comm=None):
        # comm: model space split over the ranks (dist.BandComm / FrameComm) - dot products and norms are summed across them
        self.comm = comm if (comm is not None and comm.is_sharded("model")) else None
        if A.shape[0] != A.shape[1]:
            raise Exception("Please check the size of the matrx A: it should be square in order to apply hybrid GMRES")
        if reorth not in ("mgs", "cgs2"):
            raise ValueError("reorth must be 'mgs' or 'cgs2'")
        self.A, self.reorth = A, reorth
        n = A.shape[1]
        dev = b_dev.device
        self.V = Basis(n, kmax + 1, dev)
        self.H = torch.zeros((kmax + 1, kmax + 2), dtype=F64, device=dev)  # row j = column j of the Hessenberg matrix
        self.nrm = torch.zeros((kmax + 1, 2), dtype=F64, device=dev)
        self.beta0 = torch.zeros(2, dtype=F64, device=dev)
        self._w = torch.empty(n, dtype=F64, device=dev)
        self._h2 = torch.zeros(kmax + 2, dtype=F64, device=dev)
        self._dot = torch.zeros(2, dtype=F64, device=dev)
        K.vec_norm2(b_dev, out=self.beta0)
        self._sync_norm(self.beta0)
        K.vec_div(b_dev, self.beta0[1:2], out=self.V.next_col())
        self.V.push()

    def _sync_norm(self, pair):
        if self.comm is not None:
            self.comm.sync_norm_(pair, "model")

    def _sum(self, t):
        if self.comm is not None:
            self.comm.sum_(t, "model")

    @property
    def k(self):
        return self.V.k - 1

    def step(self):
        k = self.V.k  # number of basis vectors so far; new Hessenberg column has k+1 entries
        if k >= self.H.shape[0]:
            raise RuntimeError("ArnoldiState capacity exceeded")
        w = self._w
        hcol = self.H[k - 1]
        self.A.apply_dev(self.V.col(k - 1), out=w)  # w = A v_k                     (decompositions.py:212)
        if self.reorth == "mgs":
            for j in range(k):  # h_j = v_j . w ; w -= h_j v_j                        (decompositions.py:216-218)
                K.vec_dot(self.V.col(j), w, out=self._dot)
                self._sum(self._dot[0:1])
                hcol[j:j + 1].copy_(self._dot[0:1])
                K.vec_axpy(self._dot[0:1], self.V.col(j), w, out=w, sign=-1.0,
                           norm_out=self.nrm[k - 1] if j == k - 1 else None)
        else:
            K.basis_dots(self.V, k, w, out=hcol)
            self._sum(hcol[:k])
            K.basis_combine(self.V, k, hcol, w=w, sign=-1.0, out=w)
            K.basis_dots(self.V, k, w, out=self._h2)
            self._sum(self._h2[:k])
            K.basis_combine(self.V, k, self._h2, w=w, sign=-1.0, out=w, norm_out=self.nrm[k - 1])
            hcol[:k].add_(self._h2[:k])
        self._sync_norm(self.nrm[k - 1])
        hcol[k:k + 1].copy_(self.nrm[k - 1, 1:2])  # h_{k+1,k} = ||w||              (decompositions.py:224-226)
        K.vec_div(w, self.nrm[k - 1, 1:2], out=self.V.next_col())
        self.V.push()

    def H_host(self):
        k = self.V.k - 1
        Hd = self.H[:k, :k + 1].cpu().numpy()
        return np.ascontiguousarray(Hd.T)  # (k+1) x k upper Hessenberg


# ---- reference-signature functions (NumPy in / NumPy out, arithmetic on the GPU) -------------------------------

class _HostBasis:
    """Pinned, column-contiguous host capacity buffer behind the arrays golub_kahan_update / arnoldi_update return.

    The reference returns FRESH arrays from np.hstack at every call (decompositions.py:243-247).  Here the returned
    array is a view of the first k columns of a buffer, and the next call appends column k in place instead of
    re-copying - but ONLY when that is indistinguishable from the reference's behaviour: the array handed back by the
    caller must be the buffer's newest view (k == high-water mark `self.k`).  A call with an older / truncated view
    (restart, branching, two solves sharing a prefix) would otherwise overwrite a column of an array some caller still
    holds; such a call gets a fresh buffer and a copy, exactly like the reference.  Buffers are found again through the
    data pointer of the array passed in; the registry keeps the `MAX_LIVE` most recently used ones (each pins up to GBs
    of host memory) and `release_host_buffers()` drops them all at the end of a run."""

    registry = {}  # data pointer -> _HostBasis, in least-recently-used order
    MAX_LIVE = 4

    def __init__(self, rows, cap):
        self.t = torch.empty((cap, rows), dtype=F64)
        if torch.cuda.is_available():
            self.t = self.t.pin_memory()
        self.cap, self.rows = cap, rows
        self.k = 0  # high-water mark: columns [0, k) have been handed out
        self.arr = self.t.numpy().T  # (rows, cap), Fortran order: column j contiguous
        self.key = self.arr.__array_interface__["data"][0]
        _HostBasis.registry[self.key] = self
        while len(_HostBasis.registry) > _HostBasis.MAX_LIVE:
            _HostBasis.registry.pop(next(iter(_HostBasis.registry)))

    @classmethod
    def adopt(cls, M, extra=1):
        """Return (buffer, k) such that buffer.arr[:, :k] holds M and columns k .. k+extra-1 may be written."""
        M = np.asarray(M)
        if M.ndim == 1:
            M = M.reshape(-1, 1)
        rows, k = M.shape
        key = M.__array_interface__["data"][0] if k else None
        hb = cls.registry.get(key)
        if hb is not None and hb.rows == rows and M.strides == (8, 8 * rows) and k == hb.k and k + extra <= hb.cap:
            cls.registry[key] = cls.registry.pop(key)  # most recently used
            return hb, k
        # not the newest view of a live buffer (or no room left): copy, as the reference's np.hstack does.
        # pinning is expensive (page-locking): start with room for a typical run and double from there
        new = cls(rows, max(2 * (k + extra), 64 if rows * 64 * 8 <= (4 << 30) else 16))
        new.arr[:, :k] = M
        new.k = k
        if hb is not None and hb.rows == rows and k == hb.k:
            cls.registry.pop(key, None)  # outgrown: its contents moved to `new`
        return new, k

    def view(self, k):
        self.k = max(self.k, k)
        return self.arr[:, :k]


def release_host_buffers():
    """Drop the pinned host capacity buffers of golub_kahan_update / arnoldi_update (arrays already returned stay
    valid: they keep their buffer alive; later calls with them copy into a fresh buffer)."""
    _HostBasis.registry.clear()


_COPY_STREAMS = {}


def _copy_stream(dev):
    """Side stream for host<->device copies that overlap with the operator kernels (one per device)."""
    st = _COPY_STREAMS.get(dev)
    if st is None:
        st = _COPY_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return st


def golub_kahan_update(A, U, S, V, b200_comm=None):
    """One Golub-Kahan step; same contract as trips.utilities.decompositions.golub_kahan_update (:230-255).
    (b200_comm: a dist.RowComm when A is a row-sharded operator and U holds this rank's rows - the squared norm of the
    new u is then summed over the ranks; everything else is as on one GPU.)

    U: m x k, S: (k x (k-1)) bidiagonal or np.empty(1) on the first call, V: n x (k-1) (ignored on the first
    call, exactly as the reference replaces the caller's dummy).  Returns (U: m x (k+1), S: (k+1) x k, V: n x k).

    Host traffic per call: u_k and v_{k-1} up, the new u and v down, all through pinned buffers.  Two of the four
    copies hide behind the kernels: v_{k-1} arrives on a side stream while A^T u_k runs (it is only needed by the
    recurrence that follows), and the new v leaves on the side stream while A v runs."""
    A = as_operator(A)
    dev = A.device
    main = torch.cuda.current_stream(dev)
    side = _copy_stream(dev)
    first = (S.shape[0] == 1)
    k = 1 if first else S.shape[0]
    hu, ku = _HostBasis.adopt(U)
    v_prev = None
    if first:
        hv, kv = _HostBasis.adopt(np.empty((A.shape[1], 0)))
    else:
        hv, kv = _HostBasis.adopt(V)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            v_prev = hv.t[kv - 1].to(dev, non_blocking=True)
    u_k = hu.t[ku - 1].to(dev, non_blocking=True)
    v = torch.empty(A.shape[1], dtype=F64, device=dev)
    pair = torch.zeros(4, dtype=F64, device=dev)
    if first:
        apply_fused(A, u_k, v, adjoint=True, norm_out=pair[0:2])
    else:
        A.adjoint_dev(u_k, out=v)  # v_{k-1} is still in flight
        main.wait_stream(side)
        v_prev.record_stream(main)
        K.vec_axpy(float(S[k - 1, k - 2]), v_prev, v, out=v, norm_out=pair[0:2], sign=-1.0)
    K.vec_div(v, pair[1:2], out=v)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        hv.t[kv].copy_(v, non_blocking=True)  # overlaps with A v below
    u = torch.empty(A.shape[0], dtype=F64, device=dev)
    apply_fused(A, v, u, coef=pair[1:2], z=u_k, norm_out=pair[2:4])
    if b200_comm is not None:
        b200_comm.sync_norm_(pair[2:4], "data")
    K.vec_div(u, pair[3:4], out=u)
    hu.t[ku].copy_(u, non_blocking=True)
    main.wait_stream(side)
    v.record_stream(side)
    sc = pair.cpu().numpy()  # synchronises the stream: the host copies above are complete after this
    alpha, beta = sc[1], sc[3]
    if first:
        Snew = np.array([[alpha], [beta]])
    else:
        Snew = np.zeros((k + 1, k))
        Snew[:k, :k - 1] = S
        Snew[k - 1, k - 1] = alpha
        Snew[k, k - 1] = beta
    return hu.view(ku + 1), Snew, hv.view(kv + 1)


def golub_kahan(A, b, n_iter, dp_stop=False, **kwargs):
    """Batch Golub-Kahan, contract of trips.utilities.decompositions.golub_kahan (:118-205): returns NumPy
    (U: m x (k+1), S: (k+1) x k, V: n x k) with k = n_iter, or fewer when `dp_stop=True` halts the factorisation by the
    discrepancy principle (kwargs gk_eta = 1.001, gk_delta = 0.001 as in the reference, :148-149).  The basis stays on
    the device while it is built."""
    st = golub_kahan_device(A, b, n_iter, dp_stop=dp_stop, gk_eta=kwargs.get("gk_eta", 1.001),
                            gk_delta=kwargs.get("gk_delta", 0.001))
    return st.U.to_numpy(), st.B_host(), st.V.to_numpy()


def golub_kahan_device(A, b, n_iter, kmax=None, dp_stop=False, gk_eta=1.001, gk_delta=0.001, comm=None):
    """GKState after n_iter steps (device-resident result of golub_kahan).

    dp_stop=True follows decompositions.py:164-195: after every step solve min ||S y - U^T b|| on the host (k+1 x k),
    lift x = V y on the device, and stop before the next step once ||A x - b|| <= gk_eta * gk_delta.  Per step this
    costs one extra product with A and two basis passes; only the k+1 coefficients and the residual norm cross PCIe."""
    A = as_operator(A)
    bd = to_device_vector(b, A.device)
    st = GKState(A, bd, max(kmax or n_iter, n_iter), comm=comm)
    if dp_stop != True:  # noqa: E712  (the reference tests `dp_stop == True`, :166,185)
        for _ in range(n_iter):
            st.step()
        return st
    res_norm = np.inf
    m, n = A.shape
    xd = torch.empty(n, dtype=F64, device=A.device)
    rd = torch.empty(m, dtype=F64, device=A.device)
    pair = K.new_pair(A.device)
    for _ in range(n_iter):
        if res_norm <= gk_eta * gk_delta:  #                                                       (:166-168)
            break
        st.step()
        k = st.k
        bhat = K.basis_dots(st.U, k + 1, bd)  # U.T @ b                                            (:189)
        if comm is not None:
            comm.sum_(bhat, "data")
        y = np.linalg.lstsq(st.B_host(), bhat[:k + 1].cpu().numpy(), rcond=None)[0]  #             (:191)
        K.basis_combine(st.V, k, torch.from_numpy(np.ascontiguousarray(y)).to(A.device), out=xd)  # x = V @ y   (:193)
        A.apply_dev(xd, out=rd)
        K.vec_diffnorm2(rd, bd, out=pair)  # ||A x - b||                                           (:195)
        if comm is not None:
            comm.sync_norm_(pair, "data")
        res_norm = float(pair[1])
    return st


def arnoldi_update(A, V, H, reorth="mgs"):
    """One Arnoldi step; contract of trips.utilities.decompositions.arnoldi_update (:207-228).
    V: n x k, H: k x (k-1) or np.empty(1) on the first call.  Returns (V: n x (k+1), H: (k+1) x k)."""
    A = as_operator(A)
    dev = A.device
    k = H.shape[0]
    hv, kv = _HostBasis.adopt(V)
    Vd = hv.t[:kv].to(dev, non_blocking=True)  # (k, n): column j of V is row j
    w = torch.empty(A.shape[0], dtype=F64, device=dev)
    A.apply_dev(Vd[kv - 1], out=w)
    h = torch.zeros(kv + 1, dtype=F64, device=dev)
    nrm = torch.zeros(2, dtype=F64, device=dev)
    if reorth == "mgs":
        dot = torch.zeros(2, dtype=F64, device=dev)
        for j in range(kv):
            K.vec_dot(Vd[j], w, out=dot)
            h[j:j + 1].copy_(dot[0:1])
            K.vec_axpy(dot[0:1], Vd[j], w, out=w, sign=-1.0, norm_out=nrm if j == kv - 1 else None)
    else:
        h2 = torch.zeros(kv, dtype=F64, device=dev)
        K.basis_dots(Vd, kv, w, out=h)
        K.basis_combine(Vd, kv, h, w=w, sign=-1.0, out=w)
        K.basis_dots(Vd, kv, w, out=h2)
        K.basis_combine(Vd, kv, h2, w=w, sign=-1.0, out=w, norm_out=nrm)
        h[:kv].add_(h2)
    h[kv:kv + 1].copy_(nrm[1:2])
    K.vec_div(w, nrm[1:2], out=w)
    hv.t[kv].copy_(w, non_blocking=True)
    hcol = h.cpu().numpy()
    if k == 1:
        Hnew = hcol.reshape(-1, 1)
    else:
        Hnew = np.zeros((k + 1, k))
        Hnew[:k, :k - 1] = H
        Hnew[:, k - 1] = hcol
    return hv.view(kv + 1), Hnew
