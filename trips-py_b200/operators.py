"""Operators for the reference's `A` / `L` arguments, backed by the sm_100a kernels.

The reference has no plugin registry: the API is duck typing on the object passed as A or L
(SURVEY.md section 8b).  The operations it uses are `.shape`, `A @ x`, `A.T @ y` with x 1-D, an (n,1) column or
an n x k block, `L * v`, `.todense()`, and `isinstance(., pylops.LinearOperator)`
(trips/solvers/Hybrid_LSQR.py:63,102; GKS.py:37-38,84,95; MMGKS.py:43-60,117,127;
trips/utilities/decompositions.py:177-181,212,235-240).  `LinearOperator` below provides that surface (and
subclasses pylops.LinearOperator when pylops is importable), so these objects can be handed to the reference's
own solvers unchanged: NumPy in -> H2D -> kernel -> D2H -> NumPy out.  The solvers of this package call the
device-level methods (`apply_dev` / `adjoint_dev`) on torch CUDA tensors with no host round trip.
"""
import os

import numpy as np
import torch

from . import kernels as K
from .kernels import F64, CSRDevice

try:  # pragma: no cover - pylops is not installed in the build image
    if os.environ.get("TRIPS_B200_NO_PYLOPS"):
        raise ImportError
    from pylops import LinearOperator as _PylopsBase
except Exception:  # noqa: BLE001
    _PylopsBase = None

_Base = _PylopsBase if _PylopsBase is not None else object


def default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("trips_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_device_vector(x, device):
    """1-D float64 CUDA tensor from a NumPy array / torch tensor of shape (n,) or (n,1)."""
    if isinstance(x, torch.Tensor):
        t = x.reshape(-1)
        if t.device != device or t.dtype != F64:
            t = t.to(device=device, dtype=F64)
        return t.contiguous()
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    return torch.from_numpy(a).to(device)


class LinearOperator(_Base):
    """pylops-style linear operator whose arithmetic runs on the GPU.

    Subclasses implement `apply_dev(x, out=None)` and `adjoint_dev(y, out=None)` on 1-D float64 CUDA tensors."""

    fused = False  # True: apply_dev/adjoint_dev accept coef=, z=, norm_out= (fused recurrence epilogue)

    def __init__(self, shape, device=None, dtype=np.float64):
        self._tb_shape = (int(shape[0]), int(shape[1]))
        self._tb_dtype = np.dtype(dtype)
        self.device = torch.device(device) if device is not None else default_device()
        if _PylopsBase is not None:  # pragma: no cover
            try:
                _PylopsBase.__init__(self, dtype=self._tb_dtype, shape=self._tb_shape)
            except Exception:  # noqa: BLE001
                pass

    # -- metadata ------------------------------------------------------------------------------------------
    @property
    def shape(self):
        return self._tb_shape

    @shape.setter
    def shape(self, value):  # pylops' constructor assigns it
        self._tb_shape = (int(value[0]), int(value[1]))

    @property
    def dtype(self):
        return self._tb_dtype

    @dtype.setter
    def dtype(self, value):
        self._tb_dtype = np.dtype(value)

    # -- device level (subclasses) -----------------------------------------------------------------------------
    def apply_dev(self, x, out=None):
        raise NotImplementedError

    def adjoint_dev(self, y, out=None):
        raise NotImplementedError

    # -- host/torch level ----------------------------------------------------------------------------------------
    def _run(self, fn, x, n_in):
        if isinstance(x, torch.Tensor):
            if x.dim() == 2 and x.shape[1] > 1:
                cols = [fn(to_device_vector(x[:, j], self.device)) for j in range(x.shape[1])]
                return torch.stack(cols, dim=1)
            y = fn(to_device_vector(x, self.device))
            return y.reshape(-1, 1) if x.dim() == 2 else y
        a = np.asarray(x)
        if a.ndim == 2 and a.shape[1] > 1:
            if a.shape[0] != n_in:
                raise ValueError(f"dimension mismatch: operator expects {n_in} rows, got {a.shape}")
            return np.stack([fn(to_device_vector(a[:, j], self.device)).cpu().numpy() for j in range(a.shape[1])], axis=1)
        if a.size != n_in:
            raise ValueError(f"dimension mismatch: operator expects {n_in} elements, got {a.shape}")
        y = fn(to_device_vector(a, self.device)).cpu().numpy()
        return y.reshape(-1, 1) if a.ndim == 2 else y

    def matvec(self, x):
        return self._run(self.apply_dev, x, self.shape[1])

    def rmatvec(self, y):
        return self._run(self.adjoint_dev, y, self.shape[0])

    matmat = matvec
    rmatmat = rmatvec
    # pylops back ends call these
    _matvec = matvec
    _rmatvec = rmatvec

    def dot(self, x):
        if np.isscalar(x):
            return _Scaled(self, float(x))
        if isinstance(x, LinearOperator):
            return _Product(self, x)
        return self.matvec(x)

    def __matmul__(self, x):
        return self.dot(x)

    def __mul__(self, x):
        return self.dot(x)

    def __rmul__(self, x):
        if np.isscalar(x):
            return _Scaled(self, float(x))
        return NotImplemented

    def __call__(self, x):
        return self.dot(x)

    @property
    def T(self):
        return _Adjoint(self)

    @property
    def H(self):
        return _Adjoint(self)

    def adjoint(self):
        return _Adjoint(self)

    def transpose(self):
        return _Adjoint(self)

    def todense(self):
        """Dense matrix by applying the operator to the identity (as pylops does); for small operators only."""
        m, n = self.shape
        out = np.empty((m, n))
        e = torch.zeros(n, dtype=F64, device=self.device)
        for j in range(n):
            e.zero_()
            e[j] = 1.0
            out[:, j] = self.apply_dev(e).cpu().numpy()
        return out

    def __repr__(self):
        return f"<{self.shape[0]}x{self.shape[1]} {type(self).__name__} on {self.device}>"


class _Adjoint(LinearOperator):
    def __init__(self, op):
        super().__init__((op.shape[1], op.shape[0]), op.device, op.dtype)
        self.op = op
        self.fused = op.fused

    def apply_dev(self, x, out=None, **kw):
        return self.op.adjoint_dev(x, out=out, **kw)

    def adjoint_dev(self, y, out=None, **kw):
        return self.op.apply_dev(y, out=out, **kw)

    @property
    def T(self):
        return self.op

    H = T


class _Scaled(LinearOperator):
    def __init__(self, op, a):
        super().__init__(op.shape, op.device, op.dtype)
        self.op, self.a = op, a

    def apply_dev(self, x, out=None):
        y = self.op.apply_dev(x, out=out)
        return y.mul_(self.a)

    def adjoint_dev(self, y, out=None):
        x = self.op.adjoint_dev(y, out=out)
        return x.mul_(self.a)


class _Product(LinearOperator):
    """A @ B as an operator (e.g. the square operator A^T A handed to Hybrid_GMRES)."""

    def __init__(self, a, b):
        if a.shape[1] != b.shape[0]:
            raise ValueError("operator shapes do not chain")
        super().__init__((a.shape[0], b.shape[1]), a.device, a.dtype)
        self.a, self.b = a, b
        self._tmp = None

    def _mid(self):
        if self._tmp is None:
            self._tmp = torch.empty(self.a.shape[1], dtype=F64, device=self.device)
        return self._tmp

    def apply_dev(self, x, out=None):
        return self.a.apply_dev(self.b.apply_dev(x, out=self._mid()), out=out)

    def adjoint_dev(self, y, out=None):
        return self.b.adjoint_dev(self.a.adjoint_dev(y, out=self._mid()), out=out)


class Identity(LinearOperator):
    def __init__(self, n, device=None):
        super().__init__((n, n), device)

    def apply_dev(self, x, out=None):
        if out is None:
            return x.clone()
        out.copy_(x)
        return out

    adjoint_dev = apply_dev


# ---- CSR tomography operator ------------------------------------------------------------------------------------

class CSROperator(LinearOperator):
    """Sparse matrix with an explicitly stored transpose: both A x and A^T u are gather-only, deterministic SpMVs.
    Stands for the scipy.sparse matrices / ASTRA projector objects the reference passes as A (demos: CT matrices
    from .mat files, `astra.OpTomo`).

    Device layouts (either or both, per matrix): CSR (`A`, `AT`: CSRDevice) and SELL-32-4, the row-interleaved form
    of the same rows (`A_sell`, `AT_sell`: SellDevice).  order='sequential' (default) sums every row in index order
    with separately rounded multiply/add - bit-identical to scipy's csr_matvec / csc_matvec - and prefers the SELL
    layout (the fast path); order='tree' is a per-row tree reduction on CSR (differs from scipy by rounding)."""

    fused = True

    def __init__(self, A=None, AT=None, order="sequential", A_sell=None, AT_sell=None):
        ref, ref_t = (A if A is not None else A_sell), (AT if AT is not None else AT_sell)
        if ref is None or ref_t is None:
            raise ValueError("need A and A^T, each in at least one layout")
        if ref.shape != (ref_t.shape[1], ref_t.shape[0]):
            raise ValueError("AT must have the transposed shape of A")
        if order not in K.ORDERS:
            raise ValueError("order must be 'sequential' (bit-identical to scipy's SpMV) or 'tree' (fastest on CSR)")
        super().__init__(ref.shape, ref.device)
        self.A, self.AT, self.A_sell, self.AT_sell, self.order = A, AT, A_sell, AT_sell, order

    def _csr(self, transposed):
        """CSR form of A (or A^T), converting from SELL on first use."""
        if transposed:
            if self.AT is None:
                self.AT = self.AT_sell.to_csr()
            return self.AT
        if self.A is None:
            self.A = self.A_sell.to_csr()
        return self.A

    def _active(self, transposed):
        sell = self.AT_sell if transposed else self.A_sell
        if self.order == "sequential" and sell is not None:
            return sell
        return self._csr(transposed)

    def with_order(self, order):
        """Same matrices, different summation order of the row sums ('sequential' = scipy's, 'tree' = fastest CSR)."""
        return CSROperator(self.A, self.AT, order, self.A_sell, self.AT_sell)

    def with_layout(self, layout):
        """Restrict to one device layout ('csr' or 'sell'); the other is dropped (converted first if missing)."""
        if layout == "csr":
            return CSROperator(self._csr(False), self._csr(True), self.order)
        if layout == "sell":
            a = self.A_sell if self.A_sell is not None else K.SellDevice.from_csr(self.A)
            at = self.AT_sell if self.AT_sell is not None else K.SellDevice.from_csr(self.AT)
            return CSROperator(None, None, self.order, a, at)
        raise ValueError("layout must be 'csr' or 'sell'")

    @classmethod
    def from_scipy(cls, A, device=None):
        """Upload a scipy.sparse matrix (both layouts); the transpose is formed once on the host (A.T.tocsr())."""
        import scipy.sparse as sp

        device = torch.device(device) if device is not None else default_device()
        A = sp.csr_matrix(A, dtype=np.float64)
        A.sort_indices()
        AT = A.T.tocsr()
        AT.sort_indices()
        a, at = _upload_csr(A, device), _upload_csr(AT, device)
        return cls(a, at, "sequential", K.SellDevice.from_csr(a), K.SellDevice.from_csr(at))

    @classmethod
    def from_dense(cls, A, device=None):
        import scipy.sparse as sp

        return cls.from_scipy(sp.csr_matrix(np.asarray(A, dtype=np.float64)), device)

    def to_scipy(self):
        """The same arrays as a scipy.sparse.csr_matrix (used to feed the oracle identical inputs)."""
        return _download_csr(self._csr(False))

    def transpose_to_scipy(self):
        return _download_csr(self._csr(True))

    def with_f32_storage(self):
        """fp32-storage / fp64-accumulate variant (16 instead of 24 B/nnz per Golub-Kahan iteration)."""
        f = lambda M: None if M is None else M.to_f32_storage()  # noqa: E731
        return CSROperator(f(self.A), f(self.AT), self.order, f(self.A_sell), f(self.AT_sell))

    @property
    def nnz(self):
        return (self.A if self.A is not None else self.A_sell).nnz

    def apply_dev(self, x, out=None, coef=None, z=None, norm_out=None):
        return K.spmv(self._active(False), x, out=out, coef=coef, z=z, norm_out=norm_out, order=self.order)

    def adjoint_dev(self, y, out=None, coef=None, z=None, norm_out=None):
        return K.spmv(self._active(True), y, out=out, coef=coef, z=z, norm_out=norm_out, order=self.order)


def _upload_csr(A, device):
    rowptr = torch.from_numpy(np.ascontiguousarray(A.indptr, dtype=np.int64)).to(device)
    colidx = torch.from_numpy(np.ascontiguousarray(A.indices, dtype=np.int32)).to(device)
    vals = torch.from_numpy(np.ascontiguousarray(A.data, dtype=np.float64)).to(device)
    return CSRDevice(A.shape, rowptr, colidx, vals)


def _download_csr(A):
    import scipy.sparse as sp

    nnz = A.nnz
    idx_dtype = np.int32 if nnz < 2 ** 31 else np.int64
    return sp.csr_matrix((A.vals.to(F64).cpu().numpy(), A.colidx.cpu().numpy().astype(idx_dtype),
                          A.rowptr.cpu().numpy().astype(idx_dtype)), shape=A.shape)


def ct_angles(n_angles):
    """theta = linspace(0, pi, views, endpoint=False)  (trips/test_problems/Tomography.py:56)."""
    return np.linspace(0.0, np.pi, int(n_angles), endpoint=False)


def ct_num_detectors(nx):
    """p = int(sqrt(2) * nx)  (trips/test_problems/Tomography.py:53)."""
    return int(np.sqrt(2) * nx)


class ParallelBeamCT(CSROperator):
    """2-D parallel-beam tomography matrix built on the device (line model: chord length per pixel).

    Geometry follows the reference's conventions (Tomography.py:53-56, io.py:392-400): `views` angles in [0, pi),
    n_det = int(sqrt(2)*nx) unit-spaced detector bins, sinogram ordered angle-major (row = angle*n_det + det),
    image vectorised row-major.  `angle_subset` keeps only those angle indices (row sharding by projection angle).

    layout: 'csr' / 'sell' / 'both' store A and A^T (12 B per entry each); 'implicit' stores nothing: every entry is
    re-evaluated in the kernels (ray-driven forward projection, pixel-driven back-projection) - the reference's operator
    is matrix-free too (astra.OpTomo).  All layouts give bit-identical products."""

    def __init__(self, nx, views, ny=None, n_det=None, angles=None, angle_subset=None, device=None, layout="auto",
                 forward=None):
        device = torch.device(device) if device is not None else default_device()
        ny = nx if ny is None else ny
        n_det = ct_num_detectors(nx) if n_det is None else int(n_det)
        theta = ct_angles(views) if angles is None else np.asarray(angles, dtype=np.float64)
        if angle_subset is not None:
            theta = theta[np.asarray(angle_subset)]
        self.nx, self.ny, self.n_det, self.theta = int(nx), int(ny), n_det, theta
        cos_t = torch.from_numpy(np.cos(theta)).to(device)
        sin_t = torch.from_numpy(np.sin(theta)).to(device)
        K._lib.require_device()
        self.projector = None
        if layout == "auto":  # SELL is the fast stored path; keep CSR too while it is cheap (tests, export, 'tree' order)
            layout = "both" if 1.3 * len(theta) * self.nx * self.ny <= 2e8 else "sell"
        if layout == "implicit":
            # nothing stored: ray-driven forward projector, pixel-driven back-projector (forward='index': round 1's
            # forward projector, which streams A's column indices)
            self.projector = K.CTProjector(self.nx, self.ny, n_det, cos_t, sin_t, forward=forward)
            LinearOperator.__init__(self, self.projector.shape, device)
            self.A = self.AT = self.A_sell = self.AT_sell = None
            self.order = "sequential"
            self._explicit = None
            return
        if layout not in ("csr", "sell", "both"):
            raise ValueError("layout must be 'auto', 'csr', 'sell', 'both' or 'implicit'")
        mats = {}
        # the SELL copy of A uses the row-aligned layout of the index-only projector with the values stored next to the
        # indices (K.CtSellDevice): the x-gathers of neighbouring rays then share sectors.  TB200_SELL_ALIGNED=0: plain SELL
        aligned = os.environ.get("TB200_SELL_ALIGNED", "1") != "0"
        for lay in (("csr", "sell") if layout == "both" else (layout,)):
            if lay == "sell" and aligned:
                a = K.ct_build_aligned(self.nx, self.ny, n_det, cos_t, sin_t)
            else:
                a = K.ct_build(self.nx, self.ny, n_det, cos_t, sin_t, transpose=False, layout=lay)
            at = K.ct_build(self.nx, self.ny, n_det, cos_t, sin_t, transpose=True, layout=lay)
            if a.nnz != at.nnz:
                raise RuntimeError(f"CT builder: nnz(A)={a.nnz} differs from nnz(A^T)={at.nnz}")
            mats[lay] = (a, at)
        csr, sell = mats.get("csr", (None, None)), mats.get("sell", (None, None))
        super().__init__(csr[0], csr[1], "sequential", sell[0], sell[1])

    # -- layout='implicit' ------------------------------------------------------------------------------------------
    def explicit(self, layout="csr"):
        """The same operator with stored matrices (built on first use) - export, 'tree' order, cross-checks."""
        if self.projector is None:
            return self
        if self._explicit is None:
            self._explicit = ParallelBeamCT(self.nx, len(self.theta), ny=self.ny, n_det=self.n_det, angles=self.theta,
                                            device=self.device, layout=layout)
        return self._explicit

    @property
    def nnz(self):
        return self.projector.nnz if self.projector is not None else super().nnz

    def _csr(self, transposed):
        return self.explicit()._csr(transposed) if self.projector is not None else super()._csr(transposed)

    def with_order(self, order):
        return self.explicit().with_order(order) if (self.projector is not None and order != "sequential") else \
            (self if self.projector is not None else super().with_order(order))

    def with_layout(self, layout):
        return self.explicit(layout).with_layout(layout) if self.projector is not None else super().with_layout(layout)

    def with_f32_storage(self):
        return self.explicit("sell").with_f32_storage() if self.projector is not None else super().with_f32_storage()

    def apply_dev(self, x, out=None, coef=None, z=None, norm_out=None):
        if self.projector is not None:
            return self.projector.forward(x, out=out, coef=coef, z=z, norm_out=norm_out)
        return super().apply_dev(x, out=out, coef=coef, z=z, norm_out=norm_out)

    def adjoint_dev(self, y, out=None, coef=None, z=None, norm_out=None):
        if self.projector is not None:
            return self.projector.backproject(y, out=out, coef=coef, z=z, norm_out=norm_out)
        return super().adjoint_dev(y, out=out, coef=coef, z=z, norm_out=norm_out)


def fan_geometry(nx):
    """(source_origin, detector_origin, detector_pixel_size) = (3 nx, nx, 4/3): Tomography.define_proj_id
    (trips/test_problems/Tomography.py:57-59)."""
    so, dd = 3.0 * nx, 1.0 * nx
    return so, dd, (so + dd) / so


class FanBeamCT(CSROperator):
    """Flat-detector fan-beam tomography matrix, line (chord-length) model - the geometry the reference's Tomography
    class requests from ASTRA (`astra.create_proj_geom('fanflat', ...)` + `'line_fanflat'`, Tomography.py:57-67):
    `views` angles in [0, pi), int(sqrt(2)*nx) bins of width (so+dd)/so, source at 3*nx, detector at nx.
    A and the exact transpose are built on the device and stored (CSR and/or SELL-32-4).  layout='implicit' keeps only
    the stored transpose: the forward product is the ray-driven matrix-free projector with per-ray geometry
    (tb200_ctfan_forward_rays_f64, bit-identical to the stored product); the pixel-driven back-projector does not carry
    over to a fan (every candidate bin has its own normal), so A^T u stays a stored SpMV - half the memory, and the
    forward product no longer streams 12 bytes per entry.
    ASTRA itself is not part of the reference tree: rotation sense and image-axis orientation follow ASTRA's documented
    conventions but cannot be pinned against it."""

    def __init__(self, nx, views, ny=None, n_det=None, angles=None, angle_subset=None, device=None, layout="auto",
                 source_origin=None, detector_origin=None, detector_pixel_size=None):
        device = torch.device(device) if device is not None else default_device()
        ny = nx if ny is None else ny
        n_det = ct_num_detectors(nx) if n_det is None else int(n_det)
        theta = ct_angles(views) if angles is None else np.asarray(angles, dtype=np.float64)
        if angle_subset is not None:
            theta = theta[np.asarray(angle_subset)]
        so, dd, dps = fan_geometry(nx)
        so = float(source_origin) if source_origin is not None else so
        dd = float(detector_origin) if detector_origin is not None else dd
        dps = float(detector_pixel_size) if detector_pixel_size is not None else (so + dd) / so
        self.nx, self.ny, self.n_det, self.theta, self.fan = int(nx), int(ny), n_det, theta, (so, dd, dps)
        cos_t = torch.from_numpy(np.cos(theta)).to(device)
        sin_t = torch.from_numpy(np.sin(theta)).to(device)
        K._lib.require_device()
        self._fan_forward = None
        if layout == "auto":  # small: both stored layouts (export, tests); large: matrix-free forward + stored transpose
            layout = "both" if 1.5 * len(theta) * self.nx * self.ny <= 2e8 else "implicit"
        if layout == "implicit":
            at = K.ct_build(self.nx, self.ny, n_det, cos_t, sin_t, transpose=True, layout="sell", fan=self.fan)
            LinearOperator.__init__(self, (at.shape[1], at.shape[0]), device)
            self.A = self.AT = self.A_sell = None
            self.AT_sell, self.order = at, "sequential"
            self._fan_forward = (cos_t, sin_t)
            return
        if layout not in ("csr", "sell", "both"):
            raise ValueError("layout must be 'auto', 'csr', 'sell', 'both' or 'implicit'")
        mats = {}
        for lay in (("csr", "sell") if layout == "both" else (layout,)):
            a = K.ct_build(self.nx, self.ny, n_det, cos_t, sin_t, transpose=False, layout=lay, fan=self.fan)
            at = K.ct_build(self.nx, self.ny, n_det, cos_t, sin_t, transpose=True, layout=lay, fan=self.fan)
            if a.nnz != at.nnz:
                raise RuntimeError(f"fan-beam builder: nnz(A)={a.nnz} differs from nnz(A^T)={at.nnz}")
            mats[lay] = (a, at)
        csr, sell = mats.get("csr", (None, None)), mats.get("sell", (None, None))
        super().__init__(csr[0], csr[1], "sequential", sell[0], sell[1])


    # -- layout='implicit': matrix-free forward product, stored transpose ----------------------------------------------------
    @property
    def nnz(self):
        return self.AT_sell.nnz if self._fan_forward is not None else super().nnz

    def _csr(self, transposed):
        if self._fan_forward is not None and not transposed and self.A is None:
            # the forward matrix on demand (export / cross-checks): built, not stored with the operator
            return K.ct_build(self.nx, self.ny, self.n_det, *self._fan_forward, transpose=False, layout="csr", fan=self.fan)
        return super()._csr(transposed)

    def apply_dev(self, x, out=None, coef=None, z=None, norm_out=None):
        if self._fan_forward is None:
            return super().apply_dev(x, out=out, coef=coef, z=z, norm_out=norm_out)
        m, n = self.shape
        K._vec(x, n, "x")
        out = torch.empty(m, dtype=F64, device=self.device) if out is None else K._vec(out, m, "out")
        ch, cd = 0.0, None
        if z is not None:
            K._vec(z, m, "z")
            ch, cd = (0.0, coef) if isinstance(coef, torch.Tensor) else (float(coef), None)
        cos_t, sin_t = self._fan_forward
        L = K.lib()
        ws = None
        if norm_out is not None:
            ws = K.Workspace.get(self.device).buf("ct_fw", int(L.tb200_ct_forward_rays_workspace_len(self.n_det, len(self.theta))))
        K.check(L.tb200_ctfan_forward_rays_f64(*self.fan, self.nx, self.ny, self.n_det, len(self.theta), K._p(cos_t), K._p(sin_t),
                                               K._p(x), K._p(out), ch, K._p(cd), K._p(z), K._p(norm_out), K._p(ws), K._stream()),
                "ctfan_forward_rays")
        K._lib.count(2 if norm_out is not None else 1)
        return out


class BlockDiagCT(CSROperator):
    """Block-diagonal dynamic-CT operator, one parallel-beam block per time frame, stored as ONE CSR matrix
    (and one CSR transpose) so a frame-major x is applied in a single launch.  Mirrors `pylops.BlockDiag` of
    per-frame projectors (trips/utilities/io.py:420) and the per-frame diagonal blocks of the CrossPhantom /
    Emoji operators (io.py:206-225).  `frame_angles[t]` holds the angles (radians) of frame t."""

    def __init__(self, nx, frame_angles, n_det=None, device=None):
        device = torch.device(device) if device is not None else default_device()
        n_det = ct_num_detectors(nx) if n_det is None else int(n_det)
        self.nx, self.n_det, self.nt = int(nx), n_det, len(frame_angles)
        K._lib.require_device()
        parts, parts_t = [], []
        for th in frame_angles:
            th = np.asarray(th, dtype=np.float64)
            c = torch.from_numpy(np.cos(th)).to(device)
            s = torch.from_numpy(np.sin(th)).to(device)
            parts.append(K.ct_build(nx, nx, n_det, c, s, transpose=False))
            parts_t.append(K.ct_build(nx, nx, n_det, c, s, transpose=True))
        a, at = _block_diag(parts), _block_diag(parts_t)
        super().__init__(a, at, "sequential", K.SellDevice.from_csr(a), K.SellDevice.from_csr(at))


def _block_diag(parts):
    m = sum(p.shape[0] for p in parts)
    n = sum(p.shape[1] for p in parts)
    rowptrs, cols, vals = [], [], []
    nnz_off, col_off = 0, 0
    for i, p in enumerate(parts):
        rp = p.rowptr if i == len(parts) - 1 else p.rowptr[:-1]
        rowptrs.append(rp + nnz_off)
        cols.append(p.colidx + col_off)
        vals.append(p.vals)
        nnz_off += p.nnz
        col_off += p.shape[1]
    return CSRDevice((m, n), torch.cat(rowptrs), torch.cat(cols).to(torch.int32), torch.cat(vals))


# ---- PSF blur ----------------------------------------------------------------------------------------------------

def gauss_psf(dim, spread):
    """Normalised Gaussian PSF, same formula as trips/test_problems/Deblurring2D.py:48-64 (Gauss)."""
    m, n = int(dim[0]), int(dim[1])
    s1, s2 = (spread, spread) if np.isscalar(spread) else (spread[0], spread[1])
    gx = np.arange(-np.fix(n / 2), np.ceil(n / 2))
    gy = np.arange(-np.fix(m / 2), np.ceil(m / 2))
    X, Y = np.meshgrid(gx, gy)
    psf = np.exp(-0.5 * ((X ** 2) / (s1 ** 2) + (Y ** 2) / (s2 ** 2)))
    psf /= psf.sum()
    return psf


class PSFBlur2D(LinearOperator):
    """Deblurring operator of Deblurring2D.forward_Op (Deblurring2D.py:66-73):
    A x = ndimage.convolve(X, PSF, mode='reflect'), A^T b = ndimage.convolve(B, flipud(fliplr(PSF)), 'reflect')."""

    def __init__(self, psf, nx, ny, device=None, mode="reflect"):
        super().__init__((nx * ny, nx * ny), device)
        psf = np.ascontiguousarray(psf, dtype=np.float64)
        self.psf, self.nx, self.ny = psf, int(nx), int(ny)
        self.mode = {"reflect": 0, "constant": 1}[mode]
        ph, pw = psf.shape
        # ndimage.convolve(x, w) == correlate(x, w[::-1, ::-1]) with the origin moved by one for even sizes
        self.ch = ph // 2 - (1 if ph % 2 == 0 else 0)
        self.cw = pw // 2 - (1 if pw % 2 == 0 else 0)
        self._w_fwd = torch.from_numpy(np.ascontiguousarray(psf[::-1, ::-1])).to(self.device)
        self._w_adj = torch.from_numpy(psf).to(self.device)

    def apply_dev(self, x, out=None):
        return K.correlate2d(x, self._w_fwd, self.nx, self.ny, self.ch, self.cw, self.mode, out=out)

    def adjoint_dev(self, y, out=None):
        return K.correlate2d(y, self._w_adj, self.nx, self.ny, self.ch, self.cw, self.mode, out=out)


# ---- finite-difference regularisation operators -------------------------------------------------------------------

class FirstDerivative1D(LinearOperator):
    """(n-1) x n forward difference, (L x)_i = x_i - x_{i+1}  (trips/utilities/operators.py:24-28)."""

    def __init__(self, n, device=None):
        super().__init__((n - 1, n), device)

    def apply_dev(self, x, out=None):
        return K.fd1d_apply(x, out=out)

    def adjoint_dev(self, r, out=None):
        return K.fd1d_adjoint(r, out=out)


class SpaceTimeDerivative(LinearOperator):
    """L = [ I_t (x) L_2D ; D_t (x) I ] on frame-major x (trips/utilities/operators.py:39-45); nt = 1 gives the 2-D
    operator [ I (x) D ; D (x) I ] of operators.py:30-36.  Matrix-free, with the IRLS weights fused:
      apply_dev(x, wout=, eps=, expo=)  also writes (u^2+eps^2)^expo       (MMGKS.py:60,93)
      adjoint_dev(r, w=)                computes L^T (w . r)                (MMGKS.py:113-117)
    For frame-sharded dynamic CT, `x_next` / `rt_prev` / `wt_prev` carry the one-frame halos."""

    def __init__(self, nx, ny, nt=1, device=None, has_next=False):
        self.nx, self.ny, self.nt, self.has_next = int(nx), int(ny), int(nt), bool(has_next)
        p2 = self.nx * (self.ny - 1) + (self.nx - 1) * self.ny
        rows = self.nt * p2 + (self.nt if self.has_next else self.nt - 1) * self.nx * self.ny
        super().__init__((rows, self.nt * self.nx * self.ny), device)

    def apply_dev(self, x, out=None, wout=None, eps=0.0, expo=0.0, x_next=None):
        if self.has_next and x_next is None:
            raise ValueError("this shard needs the next rank's first frame (x_next)")
        return K.fd_apply(x, self.nt, self.nx, self.ny, x_next=x_next, out=out, wout=wout, eps=eps, expo=expo)

    def adjoint_dev(self, r, out=None, w=None, rt_prev=None, wt_prev=None):
        return K.fd_adjoint(r, self.nt, self.nx, self.ny, has_next=self.has_next, w=w, rt_prev=rt_prev, wt_prev=wt_prev,
                            out=out)


class FirstDerivative2D(SpaceTimeDerivative):
    """2-D forward-difference operator of gen_first_derivative_operator_2D (operators.py:30-36)."""

    def __init__(self, nx, ny, device=None):
        super().__init__(nx, ny, 1, device)


class CenteredDerivative2D(LinearOperator):
    """Centred-difference gradient [I (x) D ; D (x) I] (2*nx*ny rows): the fp64 statement of the reference's
    `first_derivative_operator_2d` (trips/utilities/operators_old.py:35-45; pylops FirstDerivative(kind='centered') in
    float32 there - SURVEY.md F12).  It is the operator MMGKS' isoTV branch (MMGKS.py:61-78) is written for: the
    regulariser passed as L must have 2*nx*ny rows.  `iso_weights` fuses L x with (u1^2+u2^2+eps^2)^expo."""

    def __init__(self, nx, ny, device=None):
        self.nx, self.ny = int(nx), int(ny)
        super().__init__((2 * self.nx * self.ny, self.nx * self.ny), device)

    def apply_dev(self, x, out=None):
        return K.cd2d_apply(x, self.nx, self.ny, out=out)

    def adjoint_dev(self, r, out=None, w=None):
        return K.cd2d_adjoint(r, self.nx, self.ny, w=w, out=out)

    def iso_weights(self, x, eps, expo, out=None):
        if out is None:
            out = torch.empty(self.shape[0], dtype=F64, device=x.device)
        K.cd2d_apply(x, self.nx, self.ny, out=None, wout=out, eps=eps, expo=expo, want_u=False)
        return out


# ---- framelet analysis operator -------------------------------------------------------------------------------------

def framelet_filters(level, n):
    """The three n x n filter matrices (low pass, two high passes) of the piecewise-linear B-spline framelet at one level,
    reflective boundaries - trips/utilities/operators.py:50-83 (construct_H): stencils [1 2 1]/4, sqrt(2)/4 [-1 0 1],
    [-1 2 -1]/4 with the two outer taps `level` samples away from the centre."""
    import scipy.sparse as sp

    ones = np.ones(n)
    rows = np.arange(level)
    lo_r, lo_c = rows, level - rows - 1                # taps falling off the left edge fold back onto these columns
    hi_r, hi_c = n - 1 - rows, n - level + rows        # ... and off the right edge
    out = []
    for centre, left, right, fold_l, fold_r, scale in ((2.0, 1.0, 1.0, 1.0, 1.0, 0.25),
                                                       (0.0, -1.0, 1.0, -1.0, 1.0, np.sqrt(2) / 4),
                                                       (2.0, -1.0, -1.0, -1.0, -1.0, 0.25)):
        H = sp.spdiags(left * ones, -level, n, n) + sp.spdiags(right * ones, level, n, n)
        if centre:
            H = H + sp.spdiags(centre * ones, 0, n, n)
        H = H.tolil()
        for r, c in zip(lo_r, lo_c):
            H[r, c] += fold_l
        for r, c in zip(hi_r, hi_c):
            H[r, c] += fold_r
        out.append((H.tocsr() * scale).tocsr())
    return tuple(out)


def framelet_analysis(n, levels):
    """(2*levels+1) n x n analysis matrix, assembled exactly as the reference does (operators.py:86-101,
    create_analysis_operator): the deepest level contributes its three filters as they are; every shallower level stacks
    its two high passes under what came from below and multiplies the stack by the low pass handed down from above."""
    import scipy.sparse as sp

    def build(level, carry):
        H0, H1, H2 = framelet_filters(level, n)
        if level == levels:
            return sp.vstack((H0, H1, H2)).tocsr()
        below = build(level + 1, H0)
        stack = sp.vstack((below, H1, H2)).tocsr()
        return stack if carry is None else (stack @ carry).tocsr()

    return build(1, None)


class FrameletOperator(CSROperator):
    """Two-dimensional framelet analysis operator W x = vec_F(W_n X W_m^T), X = x reshaped (n, m) in Fortran order -
    `create_framelet_operator(n, m, l)` of the reference (trips/utilities/operators.py:104-113), used as the
    regularisation operator L of GKS / MMGKS.  Applied as ONE sparse matrix, kron(W_m, W_n), through the CSR kernels
    (the reference's two sparse-dense products sum in a different order: agreement to rounding, not bitwise)."""

    def __init__(self, n, m, levels, device=None):
        import scipy.sparse as sp

        self.n, self.m, self.levels = int(n), int(m), int(levels)
        Wn, Wm = framelet_analysis(self.n, self.levels), framelet_analysis(self.m, self.levels)
        op = CSROperator.from_scipy(sp.kron(Wm, Wn, format="csr"), device)
        super().__init__(op.A, op.AT, "sequential", op.A_sell, op.AT_sell)


def as_operator(A, device=None):
    """Accept what the reference accepts for A / L and return a GPU operator: our operators pass through;
    scipy.sparse matrices and (small) dense arrays are uploaded as CSR with an explicit transpose."""
    if isinstance(A, LinearOperator):
        return A
    try:
        import scipy.sparse as sp

        if sp.issparse(A):
            return CSROperator.from_scipy(A, device)
    except ImportError:  # pragma: no cover
        pass
    if isinstance(A, np.ndarray) and A.ndim == 2:
        return CSROperator.from_dense(A, device)
    raise TypeError(
        f"cannot run {type(A).__name__} on the GPU: pass a trips_b200 operator, a scipy.sparse matrix or a dense "
        "ndarray (host callbacks such as pylops.FunctionOperator would be a CPU fallback, which this package refuses)")
