"""Thin torch-tensor front end over the C ABI: shape/dtype checks, workspaces, stream plumbing.

PyTorch is used only for device memory and streams.  Every function enqueues on torch's current stream and
returns without synchronising.  Vectors are 1-D float64 CUDA tensors; a Basis is a (kmax, n) row-major tensor,
i.e. column j of the mathematical basis is the contiguous row j.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import c_ptr, check, lib

F64 = torch.float64


def _p(t):
    return None if t is None else c_ptr(t.data_ptr())


# torch's raw accessors: torch.cuda.current_stream() builds a Stream object through three Python layers (15 us per kernel
# launch, measured 0.3 ms per MMGKS iteration); these return the same handle / index in well under a microsecond
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _current_device():
    return _raw_device() if _raw_device is not None else torch.cuda.current_device()


def _stream():
    if _raw_stream is not None:
        return c_ptr(_raw_stream(_current_device()))
    return c_ptr(torch.cuda.current_stream().cuda_stream)


def _on_current_device(t, name):
    """Kernels are enqueued on the CURRENT device's current stream: a tensor of another GPU would be dereferenced by
    the wrong device (illegal address or silent peer reads).  Refuse instead of guessing."""
    cur = _current_device()
    if t.device.index != cur:
        raise RuntimeError(f"{name} lives on {t.device} but the current CUDA device is cuda:{cur}; run the call under "
                           f"`with torch.cuda.device({t.device.index}):` (one process per GPU sets it once)")


def _vec(t, n=None, name="vector"):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == F64 and t.dim() == 1 and t.is_contiguous()):
        raise TypeError(f"{name} must be a contiguous 1-D float64 CUDA tensor")
    _on_current_device(t, name)
    if n is not None and t.numel() != n:
        raise ValueError(f"{name} has {t.numel()} elements, expected {n}")
    return t


class Workspace:
    """Caller-owned scratch the library asks for (tb200_*_workspace_len), cached per device."""

    _cache = {}

    def __init__(self, device):
        self.device = device
        self._bufs = {}
        self.scalars = torch.zeros(64, dtype=F64, device=device)  # small pool of device scalars / norm pairs
        self._next = 0

    @classmethod
    def get(cls, device):
        ws = cls._cache.get(device)  # fast path: the very device object seen before, still current
        if ws is not None and ws.device.index == _current_device():
            return ws
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device.index != torch.cuda.current_device():
            raise RuntimeError(f"operator / workspace device {device} is not the current CUDA device "
                               f"cuda:{torch.cuda.current_device()}; run the call under `with torch.cuda.device(...)`")
        ws = cls._cache.get(device)
        if ws is None:
            _lib.require_device()
            ws = cls._cache[device] = Workspace(device)
        return ws

    def buf(self, key, length):
        length = max(int(length), 1)
        t = self._bufs.get(key)
        if t is None or t.numel() < length:
            t = self._bufs[key] = torch.empty(length, dtype=F64, device=self.device)
        return t

    def reduce(self):
        return self.buf("reduce", lib().tb200_reduce_workspace_len())

    def spmv(self, m):
        return self.buf("spmv", lib().tb200_spmv_workspace_len(int(m)))

    def basis(self, k):
        return self.buf("basis", lib().tb200_basis_workspace_len(int(k)))

    def gram(self, K):
        return self.buf("gram", lib().tb200_gram_workspace_len(int(K)))

    def pair(self):
        """A fresh 2-double slot (sum of squares, norm) from a small ring."""
        i = self._next
        self._next = (i + 2) % 64
        return self.scalars[i:i + 2]


def new_pair(device):
    return torch.zeros(2, dtype=F64, device=device)


# ---- SpMV ---------------------------------------------------------------------------------------------------

class CSRDevice:
    """One CSR matrix resident in HBM: int64 rowptr, int32 colidx, fp64 (or fp32) vals."""

    def __init__(self, shape, rowptr, colidx, vals):
        self.shape = (int(shape[0]), int(shape[1]))
        if rowptr.dtype != torch.int64 or colidx.dtype != torch.int32 or vals.dtype not in (F64, torch.float32):
            raise TypeError("CSRDevice needs int64 rowptr, int32 colidx and float64/float32 vals")
        if rowptr.numel() != self.shape[0] + 1:
            raise ValueError("rowptr length must be m+1")
        if self.shape[1] >= 2 ** 31:
            raise ValueError("column count must fit int32")
        self.rowptr, self.colidx, self.vals = rowptr.contiguous(), colidx.contiguous(), vals.contiguous()
        self.nnz = int(colidx.numel())
        if self.vals.numel() != self.nnz:
            raise ValueError("colidx and vals differ in length")
        if self.nnz and (self.vals.data_ptr() % 32 or self.colidx.data_ptr() % 16):
            raise ValueError("vals must be 32-byte aligned and colidx 16-byte aligned")
        self.device = self.vals.device

    @property
    def nbytes(self):
        return self.nnz * (4 + self.vals.element_size()) + 8 * (self.shape[0] + 1)

    def to_f32_storage(self):
        return CSRDevice(self.shape, self.rowptr, self.colidx, self.vals.to(torch.float32))


ORDERS = {"sequential": 0, "tree": 1}


def _sell_addresses(sliceptr, rowlen, skip=None):
    """Device tensors (row, addr) for every stored entry of a SELL-32-4 matrix, in CSR (row-major) order; skip: leading
    padding of each row inside its lane (row-aligned CT layout)."""
    m = rowlen.numel()
    rows = torch.repeat_interleave(torch.arange(m, device=rowlen.device, dtype=torch.int64), rowlen.to(torch.int64))
    start = torch.zeros(m + 1, dtype=torch.int64, device=rowlen.device)
    torch.cumsum(rowlen.to(torch.int64), 0, out=start[1:])
    j = torch.arange(rows.numel(), device=rowlen.device, dtype=torch.int64) - start[rows]
    if skip is not None:
        j = j + skip.to(torch.int64)[rows]
    addr = sliceptr[rows >> 5] + (j >> 2) * 128 + (rows & 31) * 4 + (j & 3)
    return rows, addr, start


def sell_slice_pointers(rowlen):
    """Slice widths (longest row of each 32-row slice, rounded up to 4) -> int64 slice pointers."""
    m = rowlen.numel()
    nsl = (m + 31) // 32
    padded = torch.zeros(nsl * 32, dtype=torch.int64, device=rowlen.device)
    padded[:m] = rowlen
    w = (padded.view(nsl, 32).max(dim=1).values + 3) // 4 * 4
    sliceptr = torch.zeros(nsl + 1, dtype=torch.int64, device=rowlen.device)
    torch.cumsum(w * 32, 0, out=sliceptr[1:])
    return sliceptr


class SellDevice:
    """The same rows as a CSR matrix, interleaved in groups of 32 ("SELL-32-4", see csrc/spmv.cu): entry j of row r at
    sliceptr[r//32] + (j//4)*128 + (r%32)*4 + j%4.  Per-row order is CSR order, so products are bit-identical."""

    def __init__(self, shape, sliceptr, rowlen, colidx, vals):
        self.shape = (int(shape[0]), int(shape[1]))
        if sliceptr.dtype != torch.int64 or rowlen.dtype != torch.int32 or colidx.dtype != torch.int32 \
                or vals.dtype not in (F64, torch.float32):
            raise TypeError("SellDevice needs int64 sliceptr, int32 rowlen/colidx and float64/float32 vals")
        self.sliceptr, self.rowlen, self.colidx, self.vals = sliceptr, rowlen, colidx, vals
        self.stored = int(colidx.numel())
        self.nnz = int(rowlen.sum().item())
        self.device = vals.device

    @classmethod
    def from_csr(cls, A):
        rowlen = (A.rowptr[1:] - A.rowptr[:-1]).to(torch.int32)
        sliceptr = sell_slice_pointers(rowlen)
        total = int(sliceptr[-1].item())
        colidx = torch.zeros(max(total, 1), dtype=torch.int32, device=A.device)[:total]
        vals = torch.zeros(max(total, 1), dtype=A.vals.dtype, device=A.device)[:total]
        _, addr, _ = _sell_addresses(sliceptr, rowlen)
        colidx[addr] = A.colidx
        vals[addr] = A.vals
        return cls(A.shape, sliceptr, rowlen, colidx, vals)

    def to_csr(self):
        _, addr, start = _sell_addresses(self.sliceptr, self.rowlen)
        return CSRDevice(self.shape, start, self.colidx[addr].contiguous(), self.vals[addr].contiguous())

    @property
    def nbytes(self):
        return self.stored * (4 + self.vals.element_size()) + 8 * self.sliceptr.numel() + 4 * self.rowlen.numel()

    def to_f32_storage(self):
        return SellDevice(self.shape, self.sliceptr, self.rowlen, self.colidx, self.vals.to(torch.float32))


class CtSellDevice(SellDevice):
    """The stored parallel-beam CT matrix A (rows = rays) in the ROW-ALIGNED SELL-32-4 layout of the index-only projector
    (CTProjector(forward='index')), values next to the indices: entry j of ray r sits at position rowskip[r] + j of its
    lane, so that the 32 rays of a slice cross the same image rows at the same positions and their x-gathers share
    32-byte sectors; the indices of shallow rays (|sin| > |cos|) address the transposed image, formed per product in
    `xT`.  The order in which a row's entries are added is CSR order: same bits as the plain layout and as scipy."""

    def __init__(self, proj, vals):
        SellDevice.__init__(self, proj.shape, proj.sliceptr, proj.rowlen, proj.colidx, vals)
        self.proj = proj  # geometry table, rowskip, xT scratch, cta_order, nx / ny / n_det / n_ang

    def to_csr(self):
        p = self.proj
        skip = p.rowskip if p.rowskip is not None else torch.zeros_like(self.rowlen)
        rows, addr, start = _sell_addresses(self.sliceptr, self.rowlen, skip)
        col = self.colidx[addr].to(torch.int64)
        if p.tshallow:  # ix*ny + iy -> iy*nx + ix for the rows whose indices address the transposed image
            ang = rows // p.n_det
            shallow = (p.sin_t.abs() > p.cos_t.abs())[ang]
            col = torch.where(shallow, (col % p.ny) * p.nx + col // p.ny, col)
        return CSRDevice(self.shape, start, col.to(torch.int32).contiguous(), self.vals[addr].contiguous())

    def to_f32_storage(self):
        return CtSellDevice(self.proj, self.vals.to(torch.float32))

    @property
    def nbytes(self):
        p = self.proj
        return (SellDevice.nbytes.fget(self) + (4 * p.rowskip.numel() if p.rowskip is not None else 0)
                + (8 * p.xT.numel() if p.xT is not None else 0) + 8 * p.geom.numel())


def _ct_sell_args(A):
    p = A.proj
    return (p.nx, p.ny, p.n_det, p.n_ang, _p(p.geom), _p(A.sliceptr), _p(A.rowlen), _p(p.rowskip), _p(A.colidx), _p(A.vals))


def _spmv_sell(A, x, out, coef, z, norm_out):
    m, n = A.shape
    coef_host, coef_dev = 0.0, None
    if z is not None:
        if isinstance(coef, torch.Tensor):
            coef_dev = coef
        else:
            coef_host = float(coef)
    ws = Workspace.get(A.device).spmv(m) if norm_out is not None else None
    if isinstance(A, CtSellDevice):
        p = A.proj
        check(lib().tb200_ct_spmv_sell_f64(*_ct_sell_args(A), int(A.vals.dtype != F64), _p(p.cta_order), _p(p.xT), _p(x),
                                           _p(out), coef_host, _p(coef_dev), _p(z), _p(norm_out), _p(ws), _stream()),
              "ct_spmv_sell")
        _lib.count((2 if norm_out is not None else 1) + (1 if p.xT is not None else 0))
        return out
    fn = lib().tb200_spmv_sell_f64 if A.vals.dtype == F64 else lib().tb200_spmv_sell_f32s
    check(fn(m, n, _p(A.sliceptr), _p(A.rowlen), _p(A.colidx), _p(A.vals), _p(x), _p(out), coef_host, _p(coef_dev), _p(z),
             _p(norm_out), _p(ws), _stream()), "spmv_sell")
    _lib.count(2 if norm_out is not None else 1)
    return out


GK_STEP_EVENTS = None  # bench.py: callable returning four torch.cuda.Event to be recorded around the two SpMV launches


def gk_step_sell(A, AT, u_k, v_prev, beta_prev, v_out, u_out, alpha_pair, beta_pair):
    """One Golub-Kahan step in a single C-ABI call (6 kernels) on SELL matrices; scalars stay on the device."""
    m, n = A.shape
    ws = Workspace.get(A.device).spmv(max(m, n))
    ev = None
    if GK_STEP_EVENTS is not None:
        ev = (ctypes.c_void_p * 4)(*[e.cuda_event for e in GK_STEP_EVENTS()])
    if isinstance(A, CtSellDevice) and A.vals.dtype == F64:
        p = A.proj
        check(lib().tb200_gk_step_sell_ct_f64(*_ct_sell_args(A), _p(p.cta_order), _p(p.xT), _p(AT.sliceptr), _p(AT.rowlen),
                                              _p(AT.colidx), _p(AT.vals), _p(u_k), _p(v_prev), _p(beta_prev), _p(v_out),
                                              _p(u_out), _p(alpha_pair), _p(beta_pair), _p(ws), ev, _stream()), "gk_step")
        _lib.count(6 + (1 if p.xT is not None else 0))
        return
    check(lib().tb200_gk_step_sell_f64(m, n, _p(A.sliceptr), _p(A.rowlen), _p(A.colidx), _p(A.vals), _p(AT.sliceptr),
                                       _p(AT.rowlen), _p(AT.colidx), _p(AT.vals), _p(u_k), _p(v_prev), _p(beta_prev),
                                       _p(v_out), _p(u_out), _p(alpha_pair), _p(beta_pair), _p(ws), ev, _stream()), "gk_step")
    _lib.count(6)


def spmv(A, x, out=None, coef=None, z=None, norm_out=None, order="sequential"):
    """out = A x - coef*z (z/coef optional), optionally norm_out[:] = (||out||^2, ||out||).

    coef may be a Python float or a 1-element device tensor (kept on the device, no sync).
    order 'sequential' reproduces scipy's csr_matvec bit for bit; 'tree' is the fastest reduction order on CSR.
    A may be a CSRDevice or a SellDevice (sequential order only)."""
    m, n = A.shape
    if isinstance(A, SellDevice):
        if order != "sequential":
            raise ValueError("the SELL layout implements the sequential (scipy) summation order only")
        _vec(x, n, "x")
        if out is None:
            out = torch.empty(m, dtype=F64, device=A.device)
        _vec(out, m, "out")
        if z is not None:
            _vec(z, m, "z")
        return _spmv_sell(A, x, out, coef, z, norm_out)
    _vec(x, n, "x")
    if out is None:
        out = torch.empty(m, dtype=F64, device=A.device)
    _vec(out, m, "out")
    coef_host, coef_dev = 0.0, None
    if z is not None:
        _vec(z, m, "z")
        if isinstance(coef, torch.Tensor):
            coef_dev = coef
        else:
            coef_host = float(coef)
    ws = Workspace.get(A.device).spmv(m) if norm_out is not None else None
    fn = lib().tb200_spmv_csr_f64 if A.vals.dtype == F64 else lib().tb200_spmv_csr_f32s
    check(fn(ORDERS[order], m, n, A.nnz, _p(A.rowptr), _p(A.colidx), _p(A.vals), _p(x), _p(out), coef_host, _p(coef_dev),
             _p(z), _p(norm_out), _p(ws), _stream()), "spmv")
    _lib.count(2 if norm_out is not None else 1)
    return out


# ---- BLAS-1 -------------------------------------------------------------------------------------------------

def _scalar(a):
    if isinstance(a, torch.Tensor):
        return 0.0, a
    return float(a), None


def vec_div(x, d, out=None):
    n = x.numel()
    out = torch.empty_like(x) if out is None else _vec(out, n, "out")
    dh, dd = _scalar(d)
    check(lib().tb200_vec_div(n, _p(_vec(x)), dh, _p(dd), _p(out), _stream()), "vec_div")
    _lib.count()
    return out


def vec_axpy(a, x, y, out=None, norm_out=None, sign=1.0):
    """out = y + sign*(a*x) with sign = +-1 (two roundings, as NumPy)."""
    n = x.numel()
    _vec(x), _vec(y, n, "y")
    out = torch.empty_like(x) if out is None else _vec(out, n, "out")
    ah, ad = _scalar(a)
    ws = Workspace.get(x.device).reduce() if norm_out is not None else None
    check(lib().tb200_vec_axpy(n, ah, _p(ad), float(sign), _p(x), _p(y), _p(out), _p(norm_out), _p(ws), _stream()), "vec_axpy")
    _lib.count(2 if norm_out is not None else 1)
    return out


def _reduce(fn, name, x, y, out):
    n = x.numel()
    _vec(x)
    if y is not None:
        _vec(y, n, "y")
    if out is None:
        out = new_pair(x.device)
    ws = Workspace.get(x.device).reduce()
    if y is None:
        check(fn(n, _p(x), _p(out), _p(ws), _stream()), name)
    else:
        check(fn(n, _p(x), _p(y), _p(out), _p(ws), _stream()), name)
    _lib.count(2)
    return out


def vec_norm2(x, out=None):
    """out = (sum x^2, ||x||) on the device."""
    return _reduce(lib().tb200_vec_norm2, "vec_norm2", x, None, out)


def vec_dot(x, y, out=None):
    return _reduce(lib().tb200_vec_dot, "vec_dot", x, y, out)


def vec_diffnorm2(x, y, out=None):
    return _reduce(lib().tb200_vec_diffnorm2, "vec_diffnorm2", x, y, out)


def vec_binary(mode, x, y, w=None, out=None):
    n = x.numel()
    _vec(x), _vec(y, n, "y")
    if w is not None:
        _vec(w, n, "w")
    out = torch.empty_like(x) if out is None else _vec(out, n, "out")
    check(lib().tb200_vec_binary(int(mode), n, _p(x), _p(y), _p(w), _p(out), _stream()), "vec_binary")
    _lib.count()
    return out


def vec_mul(x, y, out=None):
    return vec_binary(0, x, y, out=out)


def vec_sub(x, y, out=None):
    return vec_binary(1, x, y, out=out)


def vec_wsub(w, x, y, out=None):
    """out = w * (x - y)"""
    return vec_binary(2, x, y, w=w, out=out)


def vec_add(x, y, out=None):
    return vec_binary(3, x, y, out=out)


def irls_weights(v, eps, expo, out=None):
    n = v.numel()
    out = torch.empty_like(v) if out is None else _vec(out, n, "out")
    check(lib().tb200_irls_weights(n, _p(_vec(v)), float(eps), float(expo), _p(out), _stream()), "irls_weights")
    _lib.count()
    return out


# ---- basis ----------------------------------------------------------------------------------------------------

class Basis:
    """Pre-allocated column store: kmax columns of length n, column j = row j of a (kmax, n) tensor.

    Appending is a pointer bump; the reference re-copies the whole basis with np.hstack / np.column_stack at
    every iteration (decompositions.py:243-247, GKS.py:91-96)."""

    def __init__(self, n, kmax, device):
        self.n, self.kmax = int(n), int(kmax)
        self.data = torch.empty((self.kmax, self.n), dtype=F64, device=device)
        self.k = 0

    def col(self, j):
        if j < 0:
            j += self.k
        return self.data[j]

    def next_col(self):
        """The (not yet counted) slot a kernel should write the next column into."""
        if self.k >= self.kmax:
            self._grow()
        return self.data[self.k]

    def push(self):
        self.k += 1

    def _grow(self):
        new = torch.empty((max(2 * self.kmax, 4), self.n), dtype=F64, device=self.data.device)
        new[:self.k].copy_(self.data[:self.k])
        self.data, self.kmax = new, new.shape[0]

    def to_numpy(self, k=None):
        k = self.k if k is None else k
        return self.data[:k].T.cpu().numpy()


def basis_dots(V, k, w, out=None):
    """h[0:k] = V[:, :k]^T w"""
    data = V.data if isinstance(V, Basis) else V
    n = data.shape[1]
    _vec(w, n, "w")
    if out is None:
        out = torch.empty(max(k, 1), dtype=F64, device=w.device)
    if k == 0:
        return out
    ws = Workspace.get(w.device).basis(k)
    check(lib().tb200_basis_dots(n, int(k), _p(data), n, _p(w), _p(out), _p(ws), _stream()), "basis_dots")
    _lib.count(2)
    return out


def basis_combine(V, k, h, w=None, sign=1.0, out=None, norm_out=None):
    """out = w + sign * V[:, :k] h  (w None: out = sign * V h)."""
    data = V.data if isinstance(V, Basis) else V
    n = data.shape[1]
    if out is None:
        out = torch.empty(n, dtype=F64, device=data.device)
    _vec(out, n, "out")
    if w is not None:
        _vec(w, n, "w")
    ws = Workspace.get(data.device).basis(1) if norm_out is not None else None
    check(lib().tb200_basis_combine(n, int(k), _p(data), n, _p(h), _p(w), float(sign), _p(out), _p(norm_out), _p(ws),
                                    _stream()), "basis_combine")
    _lib.count(2 if norm_out is not None else 1)
    return out


def weighted_gram(B, k, w=None, extras=(), extra_weighted=(), comm=None, first_col=0):
    """Double-double Gram matrix of [diag(w) B[:, :k] | extras]; returns host arrays (Ghi, Glo) of shape (K, K).
    With a communicator (row-sharded bases) the per-rank double-double partials are all-gathered and summed in
    double-double on the host, so the result keeps its accuracy and is identical on every rank.
    first_col > 0: only the columns >= first_col (and their mirror rows) are computed; the rest of the result is zero."""
    data = B.data if isinstance(B, Basis) else B
    m = data.shape[1]
    ne = len(extras)
    K = int(k) + ne
    dev = data.device
    both = (torch.zeros if first_col > 0 else torch.empty)((2, K, K), dtype=F64, device=dev)
    Ghi, Glo = both[0], both[1]
    ws = Workspace.get(dev).gram(K)
    ext = (ctypes.c_void_p * max(ne, 1))(*[e.data_ptr() for e in extras]) if ne else None
    ewt = (ctypes.c_int * max(ne, 1))(*[int(bool(f)) for f in extra_weighted]) if ne else None
    for e in extras:
        _vec(e, m, "extra column")
    if w is not None:
        _vec(w, m, "w")
    if first_col > 0:
        check(lib().tb200_weighted_gram_panel(m, int(k), _p(data), m, _p(w), ne, ext, ewt, int(first_col), _p(Ghi), _p(Glo),
                                              _p(ws), _stream()), "weighted_gram_panel")
    else:
        check(lib().tb200_weighted_gram(m, int(k), _p(data), m, _p(w), ne, ext, ewt, _p(Ghi), _p(Glo), _p(ws), _stream()),
              "weighted_gram")
    _lib.count(2)
    if comm is None:
        both = both.cpu().numpy()  # one D2H, synchronises
        return both[0], both[1]
    parts = comm.allgather(both).cpu().numpy()  # (world, 2, K, K)
    hi, lo = np.zeros((K, K)), np.zeros((K, K))
    for p in parts:  # double-double accumulation (TwoSum), fixed rank order
        s = hi + p[0]
        bb = s - hi
        e = (hi - (s - bb)) + (p[0] - bb)
        lo = lo + (e + p[1])
        hi = s
    s = hi + lo
    return s, lo - (s - hi)


class IncrementalGram:
    """Host copy of the double-double Gram matrix of an UNWEIGHTED, append-only basis [B[:, :k] | extras].  When the
    basis has only gained columns since the last call, the device computes the new columns of G alone
    (tb200_weighted_gram_panel: O(k) block products, HBM-bound) and the old block is reused - the reference re-factors
    the whole of AV and LV at every iteration (GKS.py:54-58; MMGKS.py:57-59 with pnorm = 2, where wf == 1)."""

    def __init__(self):
        self.k, self.key, self.hi, self.lo = 0, None, None, None

    def update(self, B, k, extras=(), extra_weighted=(), comm=None):
        data = B.data if isinstance(B, Basis) else B
        key = (data.data_ptr(), tuple(e.data_ptr() for e in extras))
        c0 = self.k if (self.hi is not None and key == self.key and 0 < self.k < k) else 0
        hi, lo = weighted_gram(B, k, None, extras=extras, extra_weighted=extra_weighted, comm=comm, first_col=c0)
        if c0 > 0:  # columns < c0 of the basis block are the old ones (the extras sit after column k and are recomputed)
            hi[:c0, :c0] = self.hi[:c0, :c0]
            lo[:c0, :c0] = self.lo[:c0, :c0]
        self.k, self.key, self.hi, self.lo = int(k), key, hi, lo
        return hi, lo


def gram_factor(Ghi, Glo, k):
    """Host double-double Cholesky of the leading k x k block: returns (R, C, resid2) as in tb200_gram_factor_dd."""
    K = Ghi.shape[0]
    ne = K - k
    Ghi = np.ascontiguousarray(Ghi, dtype=np.float64)
    Glo = np.ascontiguousarray(Glo, dtype=np.float64)
    R = np.zeros((k, k))
    C = np.zeros((k, max(ne, 1)))
    res = np.zeros(max(ne, 1))
    rc = lib().tb200_gram_factor_dd(k, ne, Ghi.ctypes.data, Glo.ctypes.data, R.ctypes.data, C.ctypes.data, res.ctypes.data)
    if rc != 0:
        raise np.linalg.LinAlgError(f"Gram matrix is not positive definite at leading minor {rc}")
    return R, C[:, :ne], res[:ne]


# ---- CT builder ---------------------------------------------------------------------------------------------

def ct_build(nx, ny, n_det, cos_t, sin_t, transpose=False, layout="csr", fan=None):
    """Build A (rows = rays) or A^T (rows = pixels) for the angles in the device tables cos_t/sin_t, directly in the
    requested device layout ('csr' -> CSRDevice, 'sell' -> SellDevice).  fan = (source_origin, detector_origin,
    detector_pixel_size) selects the flat-detector fan beam instead of the parallel beam."""
    dev = cos_t.device
    n_ang = cos_t.numel()
    rows = nx * ny if transpose else n_ang * n_det
    counts = torch.zeros(rows, dtype=torch.int32, device=dev)
    which = "cols" if transpose else "rows"
    if fan is None:
        pre, cnt_fn, fill_fn = (), getattr(lib(), f"tb200_ct_count_{which}"), getattr(lib(), f"tb200_ct_fill_{which}")
    else:
        pre = tuple(float(v) for v in fan)
        cnt_fn, fill_fn = getattr(lib(), f"tb200_ctfan_count_{which}"), getattr(lib(), f"tb200_ctfan_fill_{which}")
    check(cnt_fn(*pre, nx, ny, n_det, n_ang, _p(cos_t), _p(sin_t), _p(counts), _stream()), "ct_count")
    shape = (nx * ny, n_ang * n_det) if transpose else (n_ang * n_det, nx * ny)
    _lib.count(2)
    if layout == "sell":
        sliceptr = sell_slice_pointers(counts)
        total = int(sliceptr[-1].item())
        colidx = torch.zeros(max(total, 1), dtype=torch.int32, device=dev)[:total]
        vals = torch.zeros(max(total, 1), dtype=F64, device=dev)[:total]
        check(fill_fn(*pre, nx, ny, n_det, n_ang, _p(cos_t), _p(sin_t), _p(sliceptr), 1, _p(colidx), _p(vals), _stream()),
              "ct_fill")
        return SellDevice(shape, sliceptr, counts, colidx, vals)
    if layout != "csr":
        raise ValueError("layout must be 'csr' or 'sell'")
    rowptr = torch.zeros(rows + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=rowptr[1:])
    del counts
    nnz = int(rowptr[-1].item())
    colidx = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
    vals = torch.empty(max(nnz, 1), dtype=F64, device=dev)[:nnz]
    check(fill_fn(*pre, nx, ny, n_det, n_ang, _p(cos_t), _p(sin_t), _p(rowptr), 0, _p(colidx), _p(vals), _stream()), "ct_fill")
    return CSRDevice(shape, rowptr, colidx, vals)


def ct_build_aligned(nx, ny, n_det, cos_t, sin_t):
    """The stored parallel-beam A in the row-aligned SELL-32-4 layout (CtSellDevice)."""
    proj = CTProjector(nx, ny, n_det, cos_t, sin_t, forward="stored")
    vals, proj.vals = proj.vals, None
    return CtSellDevice(proj, vals)


def forward_cta_order(nx, ny, n_det, cos_t, sin_t, device, force=False):
    """Heaviest-first CTA list of the ray-driven forward projector (tb200_ct_forward_rays_f64, cta_order): entry b =
    angle * blocks_per_angle + block.  The work of a CTA is estimated from the geometry - per ray the image rows it can
    meet times the candidates per row (lockstep form), or the rows of the run form times their aligned groups - and the
    CTAs are sorted by it (longest processing time first), so the SMs drain evenly at the end of the launch: on one
    rank's angles of an 8-GPU run the kernel's SMs were busy between 57 % and 98 % of the time in centre-out order
    (profiles/r2_forward_one_of_8_ranks_ncu_full.txt); heaviest first: 0.88 -> 0.77 ms.  Scheduling only: the product
    never depends on it.  The list is tied to the launch plan (tb200_ct_forward_rays_plan) at the time of the call."""
    n_ang = int(cos_t.numel())
    rays, nblk = ctypes.c_int(0), ctypes.c_int(0)
    check(lib().tb200_ct_forward_rays_plan(int(n_det), n_ang, ctypes.byref(rays), ctypes.byref(nblk)), "ct_forward_rays_plan")
    rays, nblk = rays.value, nblk.value
    # large launches (128-ray CTAs: many waves) drain evenly in the built-in centre-out order, which also keeps the
    # co-resident CTAs in similar image rows (measured at 2048^2 x 720: 4.43 ms against 4.55 ms heaviest-first)
    if n_ang == 0 or (rays == 128 and not force):
        return None
    c = cos_t.detach().cpu().numpy().astype(np.float64)[:, None]
    s = sin_t.detach().cpu().numpy().astype(np.float64)[:, None]
    ac, as_ = np.abs(c), np.abs(s)
    sd = (np.arange(nblk * rays) - 0.5 * (n_det - 1))[None, :]
    live = (np.arange(nblk * rays) < n_det)[None, :]
    x0, y0 = 0.5 * (nx - 1), 0.5 * (ny - 1)
    d2 = 0.5 * (ac + as_)
    with np.errstate(divide="ignore", invalid="ignore"):
        # rows met: |cy - (sd - cx c)/s| < d2/|s| for some |cx| <= x0 + 1/2  (every row when the ray is parallel to the y axis)
        reach = (ac * (x0 + 0.5) + d2) / np.maximum(as_, 1e-300)
        lo = np.clip(sd / np.where(as_ > 0, s, 1.0) - reach, -y0 - 0.5, y0 + 0.5)
        hi = np.clip(sd / np.where(as_ > 0, s, 1.0) + reach, -y0 - 0.5, y0 + 0.5)
        rows = np.where(as_ * (y0 + 0.5) + d2 > np.abs(sd) - ac * (x0 + 0.5), np.maximum(hi - lo, 0.0), 0.0)
        rows = np.where(as_ < 1e-12, np.where(np.abs(sd) < ac * (x0 + 0.5) + d2, float(ny), 0.0), rows)
        width = (ac + as_) / np.maximum(ac, 1e-300)
    run_form = as_ > 7.9 * ac
    per_row = np.where(run_form, 4.0 * (np.minimum(np.ceil(width), nx) + 2) // 4 * 1.0 + 5.0, np.ceil(np.minimum(width, 10.0)) + 0.3)
    work = np.where(live, rows * per_row, 0.0).reshape(n_ang, nblk, rays).sum(axis=2).reshape(-1)
    order = np.argsort(-work, kind="stable").astype(np.int32)
    return torch.from_numpy(order).to(device)


class CTProjector:
    """Matrix-free parallel-beam CT operator (csrc/ct_forward.cu, ct_project.cu): neither values nor indices are stored.
    forward='rays' (default): the ray-driven forward projector enumerates each ray's pixels on the fly; back-projection
    is pixel-driven.  forward='index' is round 1's projector, which streams A's SELL-32-4 column indices (4 B per entry)
    and re-evaluates only the values - kept for A/B measurements (TB200_CT_FORWARD=index).  forward='stored' builds the
    same row-aligned index layout WITH the values (the A side of ParallelBeamCT(layout='sell'), see CtSellDevice).  All
    bit-identical to the stored-matrix SpMVs."""

    def __init__(self, nx, ny, n_det, cos_t, sin_t, align=True, forward=None):
        dev = cos_t.device
        self.nx, self.ny, self.n_det, self.n_ang = int(nx), int(ny), int(n_det), int(cos_t.numel())
        self.shape = (self.n_ang * self.n_det, self.nx * self.ny)
        self.device = dev
        m = self.shape[0]
        self.cos_t, self.sin_t = cos_t, sin_t
        self.geom = torch.zeros(max(6 * self.n_ang, 2), dtype=F64, device=dev)
        check(lib().tb200_ct_geometry(self.n_ang, _p(cos_t), _p(sin_t), _p(self.geom), _stream()), "ct_geometry")
        _lib.count(1)
        self.forward_mode = forward or os.environ.get("TB200_CT_FORWARD", "rays")
        if self.forward_mode not in ("rays", "index", "stored"):
            raise ValueError("forward must be 'rays', 'index' or 'stored'")
        self._nnz = None
        self.sliceptr = self.rowlen = self.rowskip = self.colidx = self.cta_order = self.xT = self.vals = None
        self.stored = 0
        self.tshallow = False
        if self.forward_mode in ("index", "stored"):
            self._build_index(align, with_vals=self.forward_mode == "stored")
        elif os.environ.get("TB200_CT_FORWARD_ORDER", "lpt") == "lpt":
            self.cta_order = forward_cta_order(self.nx, self.ny, self.n_det, cos_t, sin_t, dev)

    @property
    def nnz(self):
        """Entries of the (never stored) matrix: one counting pass of the builder, on first use."""
        if self._nnz is None:
            counts = torch.zeros(max(self.shape[0], 1), dtype=torch.int32, device=self.device)
            check(lib().tb200_ct_count_rows(self.nx, self.ny, self.n_det, self.n_ang, _p(self.cos_t), _p(self.sin_t),
                                            _p(counts), _stream()), "ct_count")
            _lib.count(1)
            self._nnz = int(counts.to(torch.int64).sum().item())
        return self._nnz

    def _build_index(self, align, with_vals=False):
        dev, m = self.device, self.shape[0]
        cos_t, sin_t = self.cos_t, self.sin_t
        self.rowlen = torch.zeros(m, dtype=torch.int32, device=dev)
        first = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
        run0 = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
        check(lib().tb200_ct_count_rows_first(self.nx, self.ny, self.n_det, self.n_ang, _p(cos_t), _p(sin_t),
                                              _p(self.rowlen), _p(first), _p(run0), _stream()), "ct_count")
        # Shallow rays (|sin| > |cos|) run along the image rows: neighbouring rays sit in different rows, so in the image
        # their gathers share nothing.  Their indices therefore address the TRANSPOSED image (formed per product in a
        # scratch vector), where they behave exactly like steep rays do in the image itself.
        self.tshallow = os.environ.get("TB200_CT_TRANSPOSE", "1") != "0"
        self.xT = torch.empty(self.shape[1], dtype=F64, device=dev) if self.tshallow else None
        # Alignment.  Lane = ray; the 32 rays of a slice are neighbours on the detector.  A steep ray crosses every image
        # row in 1 + |tan| pixels on average, but rays that enter through the side of the image start at different rows,
        # so at the same position j the lanes would sit in different rows and every x-gather would touch its own
        # sector.  Give each ray a leading padding of (its first row - the slice's first row) * (1 + |tan|) positions:
        # the lanes then walk through the same rows together (simulated: 0.49 -> 0.35 sectors per entry).  Shallow rays
        # in the transposed image: the same with rows and columns exchanged (key: where the ray enters its first row).
        span = self.rowlen
        if os.environ.get("TB200_CT_ALIGN", "1") == "0":  # measurement switch (tools/mf_probe.py)
            align = False
        if align and m > 0:
            nsl = (m + 31) // 32
            big = torch.iinfo(torch.int32).max
            ang = torch.arange(m, device=dev, dtype=torch.int64) // self.n_det
            ct, st = cos_t.abs()[ang], sin_t.abs()[ang]
            steep = ct >= st
            live = self.rowlen > 0

            def lead_of(key, mask, rate):
                """rate * (key - min of key over the slice's rows selected by mask), 0 elsewhere"""
                f = torch.full((nsl * 32,), big, dtype=torch.int32, device=dev)
                f[:m] = torch.where(mask, key[:m], torch.full_like(key[:m], big))
                fmin = f.view(nsl, 32).min(dim=1).values.repeat_interleave(32)[:m]
                return torch.where(mask, (f[:m] - fmin).to(F64) * rate, torch.zeros_like(rate))

            # measured (tools/mf_probe.py, bench.py): aligning the rays up to 45 degrees is best when the launch has many
            # waves of CTAs (whole cfg4 problem: 6.93 vs 7.10 ms); on an angle shard (1/8 of the rows, ~3 waves) the
            # padding of the longest, near-diagonal slices lengthens the critical path, and stopping at tan = 0.7 wins
            # (0.99 vs 1.13 ms)
            waves = (m / 128.0) / (4 * torch.cuda.get_device_properties(dev).multi_processor_count)
            tan_max = float(os.environ.get("TB200_CT_ALIGN_TAN", "1.0" if waves >= 8 else "0.7"))
            lead = lead_of(first, live & steep & (st <= tan_max * ct), 1.0 + st / ct.clamp_min(1e-300))
            if self.tshallow:
                lead = lead + lead_of(run0, live & ~steep & (ct <= tan_max * st), 1.0 + ct / st.clamp_min(1e-300))
            self.rowskip = torch.floor(lead).to(torch.int32).contiguous()
            span = self.rowlen + self.rowskip
        self.sliceptr = sell_slice_pointers(span)
        total = int(self.sliceptr[-1].item())
        self.colidx = torch.zeros(max(total, 1), dtype=torch.int32, device=dev)[:total]
        if with_vals:
            self.vals = torch.zeros(max(total, 1), dtype=F64, device=dev)[:total]
            check(lib().tb200_ct_fill_rows_aligned_vals(self.nx, self.ny, self.n_det, self.n_ang, _p(cos_t), _p(sin_t),
                                                        _p(self.sliceptr), _p(self.rowskip), int(self.tshallow),
                                                        _p(self.colidx), _p(self.vals), _stream()), "ct_fill")
        else:
            check(lib().tb200_ct_fill_rows_aligned(self.nx, self.ny, self.n_det, self.n_ang, _p(cos_t), _p(sin_t),
                                                   _p(self.sliceptr), _p(self.rowskip), int(self.tshallow), _p(self.colidx),
                                                   _stream()), "ct_fill")
        _lib.count(3)
        self.stored = total
        self._nnz = int(self.rowlen.sum().item())
        # CTA schedule of the forward projector: groups of four slices, heaviest first (central rays are ~2000 entries
        # long, peripheral ones a handful: in launch order the last wave would be stragglers)
        width = self.sliceptr[1:] - self.sliceptr[:-1]
        ngroups = (width.numel() + 3) // 4
        padded = torch.zeros(ngroups * 4, dtype=torch.int64, device=dev)
        padded[:width.numel()] = width
        self.cta_order = torch.argsort(padded.view(ngroups, 4).sum(dim=1), descending=True, stable=True).to(torch.int32)

    @property
    def nbytes(self):
        if self.forward_mode == "rays":
            return 8 * self.geom.numel() + 16 * self.n_ang
        return (4 * self.stored + 8 * self.sliceptr.numel() + 4 * self.rowlen.numel() + 8 * self.geom.numel()
                + (8 * self.xT.numel() if self.xT is not None else 0))

    def _coef(self, coef, z):
        if z is None:
            return 0.0, None
        return (0.0, coef) if isinstance(coef, torch.Tensor) else (float(coef), None)

    def fp64_instruction_counts(self):
        """Thread-level fp64 instructions one launch of each projector executes (bench.py's fp64-issue roofline).
        Per-unit figures are read off the SASS of the main loops and cross-checked against ncu's
        smsp__inst_executed_pipe_fp64 (profiles/):
          forward (index stream): 12 per stored position (3 index->coordinate, 3 projection, 1 t, 1 margin, 1 slope,
                  1 compare, 1 product, 1 add), padding included because it is executed;
          back-projection: 19 per pixel and angle (see ct_project.cu) + 12 per pixel of epilogue (recurrence, dd norm)."""
        npix = self.nx * self.ny
        if self.forward_mode == "rays":
            fwd, fwd_name = self._rays_fp64_count(), "ct_forward_rays_kernel (ray-driven forward projection A v, matrix-free)"
        else:  # ncu: 11.09 per stored position (profiles/r2_fp64_instr_counts_r1_kernels.csv)
            fwd = int(11.09 * self.stored) + 14 * self.shape[0]
            fwd_name = "spmv_sell_kernel<double,4,GEOM,8> (forward projection A v: index stream, values re-evaluated)"
        return {"forward": fwd, "forward_kernel": fwd_name,
                "backproject": 19 * npix * self.n_ang + 12 * npix,
                "backproject_kernel": "ct_backproject_kernel<4> (matrix-free back-projection A^T u)"}

    def _rays_fp64_count(self):
        """fp64 instructions of one ct_forward_rays launch: 9 per candidate + 2 per (ray, row) step of the lockstep form
        (9 per candidate of the run form), counted from the geometry: an angle with |s/c| <= 3 walks n*|c| (ray, row)
        pairs with LMAX candidates each; the run form evaluates aligned groups of four pixels over each run."""
        c, s = self.cos_t.abs().cpu().numpy(), self.sin_t.abs().cpu().numpy()
        npix = float(self.nx * self.ny)
        total = 0.0
        for ci, si in zip(c, s):
            if si > 3.0 * ci:  # run form: runs of 1 + si/ci pixels rounded out to groups of four, npix*ci runs
                run = 1.0 + (si / ci if ci > 0 else self.nx)
                total += 9.0 * npix * max(ci, 1.0 / self.nx) * (min(run, self.nx) + 3.0)
            else:
                lmax = int(np.ceil((ci + si) / ci + 2.1e-6))
                total += npix * ci * (9.0 * lmax + 2.0)
        return int(total) + 14 * self.shape[0]

    def forward(self, x, out=None, coef=None, z=None, norm_out=None):
        m, n = self.shape
        _vec(x, n, "x")
        out = torch.empty(m, dtype=F64, device=self.device) if out is None else _vec(out, m, "out")
        if z is not None:
            _vec(z, m, "z")
        ch, cd = self._coef(coef, z)
        if self.forward_mode == "rays":
            ws = None
            if norm_out is not None:
                ws = Workspace.get(self.device).buf("ct_fw", int(lib().tb200_ct_forward_rays_workspace_len(self.n_det, self.n_ang)))
            check(lib().tb200_ct_forward_rays_f64(self.nx, self.ny, self.n_det, self.n_ang, _p(self.geom), _p(x), _p(out), ch,
                                                  _p(cd), _p(z), _p(norm_out), _p(ws), _p(self.cta_order), _stream()),
                  "ct_forward_rays")
            _lib.count(2 if norm_out is not None else 1)
            return out
        ws = Workspace.get(self.device).spmv(m) if norm_out is not None else None
        check(lib().tb200_ct_forward_f64(self.nx, self.ny, self.n_det, self.n_ang, _p(self.geom), _p(self.sliceptr),
                                         _p(self.rowlen), _p(self.rowskip), _p(self.colidx), _p(self.cta_order), _p(self.xT), _p(x), _p(out), ch,
                                         _p(cd), _p(z),
                                         _p(norm_out),
                                         _p(ws), _stream()), "ct_forward")
        _lib.count((2 if norm_out is not None else 1) + (1 if self.xT is not None else 0))  # + the image transpose
        return out

    def backproject_rows(self, u, out, iy_begin, iy_end):
        """Image rows [iy_begin, iy_end) of A^T u into out (full-length vector): bands for comm/compute overlap."""
        m, n = self.shape
        _vec(u, m, "u")
        _vec(out, n, "out")
        check(lib().tb200_ct_backproject_rows_f64(self.nx, self.ny, int(iy_begin), int(iy_end), self.n_det, self.n_ang,
                                                  _p(self.geom), _p(u), _p(out), 0.0, None, None, None, None, _stream()),
              "ct_backproject_rows")
        _lib.count(1)
        return out

    def backproject(self, u, out=None, coef=None, z=None, norm_out=None):
        m, n = self.shape
        _vec(u, m, "u")
        out = torch.empty(n, dtype=F64, device=self.device) if out is None else _vec(out, n, "out")
        if z is not None:
            _vec(z, n, "z")
        ch, cd = self._coef(coef, z)
        ws = None
        if norm_out is not None:
            ws = Workspace.get(self.device).buf("ct_bp", int(lib().tb200_ct_backproject_workspace_len(self.nx, self.ny)))
        check(lib().tb200_ct_backproject_f64(self.nx, self.ny, self.n_det, self.n_ang, _p(self.geom), _p(u), _p(out), ch,
                                             _p(cd), _p(z), _p(norm_out), _p(ws), _stream()), "ct_backproject")
        _lib.count(2 if norm_out is not None else 1)
        return out


    def gk_step(self, u_k, v_prev, beta_prev, v_out, u_out, alpha_pair, beta_pair):
        """One Golub-Kahan step in a single C-ABI call (6 kernels + the image transpose); scalars stay on the device."""
        m, _ = self.shape
        need = max(int(lib().tb200_spmv_workspace_len(m)), int(lib().tb200_ct_backproject_workspace_len(self.nx, self.ny)),
                   int(lib().tb200_ct_forward_rays_workspace_len(self.n_det, self.n_ang)))
        ws = Workspace.get(self.device).buf("ct_gk", need)
        ev = None
        if GK_STEP_EVENTS is not None:
            ev = (ctypes.c_void_p * 4)(*[e.cuda_event for e in GK_STEP_EVENTS()])
        check(lib().tb200_gk_step_ct_f64(self.nx, self.ny, self.n_det, self.n_ang, _p(self.geom), _p(self.sliceptr),
                                         _p(self.rowlen), _p(self.rowskip), _p(self.colidx), _p(self.cta_order), _p(self.xT), _p(u_k), _p(v_prev),
                                         _p(beta_prev), _p(v_out),
                                         _p(u_out), _p(alpha_pair), _p(beta_pair), _p(ws), ev, _stream()), "gk_step_ct")
        _lib.count(7 if self.xT is not None else 6)  # 6 kernels (+ the image transpose of the index form)


# ---- stencils -------------------------------------------------------------------------------------------------

def correlate2d(x, W, nrow, ncol, ch, cw, mode=0, out=None):
    ph, pw = W.shape
    _vec(x, nrow * ncol, "image")
    out = torch.empty_like(x) if out is None else _vec(out, nrow * ncol, "out")
    check(lib().tb200_correlate2d_f64(nrow, ncol, _p(x), _p(W), ph, pw, ch, cw, mode, _p(out), _stream()), "correlate2d")
    _lib.count()
    return out


def fd_rows(nt, nrow, ncol, has_next=False):
    return int(lib().tb200_fd_rows(nt, nrow, ncol, int(has_next)))


def fd_apply(x, nt, nrow, ncol, x_next=None, out=None, wout=None, eps=0.0, expo=0.0):
    rows = fd_rows(nt, nrow, ncol, x_next is not None)
    _vec(x, nt * nrow * ncol, "x")
    if out is None:
        out = torch.empty(rows, dtype=F64, device=x.device)
    _vec(out, rows, "out")
    if wout is not None:
        _vec(wout, rows, "wout")
    check(lib().tb200_fd_apply(nt, nrow, ncol, _p(x), _p(x_next), _p(out), _p(wout), float(eps), float(expo), _stream()),
          "fd_apply")
    _lib.count()
    return out


def fd_adjoint(r, nt, nrow, ncol, has_next=False, w=None, rt_prev=None, wt_prev=None, out=None):
    rows = fd_rows(nt, nrow, ncol, has_next)
    _vec(r, rows, "r")
    if w is not None:
        _vec(w, rows, "w")
    if out is None:
        out = torch.empty(nt * nrow * ncol, dtype=F64, device=r.device)
    _vec(out, nt * nrow * ncol, "out")
    for name, h in (("rt_prev", rt_prev), ("wt_prev", wt_prev)):
        if h is not None:
            _vec(h, nrow * ncol, name)
    check(lib().tb200_fd_adjoint(nt, nrow, ncol, int(has_next), _p(r), _p(w), _p(rt_prev), _p(wt_prev), _p(out), _stream()),
          "fd_adjoint")
    _lib.count()
    return out


def cd2d_apply(x, nrow, ncol, out=None, wout=None, eps=0.0, expo=0.0, want_u=True):
    """u = [I (x) D ; D (x) I] x with centred differences; optional fused isoTV weights into wout (both halves)."""
    n = nrow * ncol
    _vec(x, n, "x")
    if want_u and out is None:
        out = torch.empty(2 * n, dtype=F64, device=x.device)
    if out is not None:
        _vec(out, 2 * n, "out")
    if wout is not None:
        _vec(wout, 2 * n, "wout")
    check(lib().tb200_cd2d_apply(nrow, ncol, _p(x), _p(out), _p(wout), float(eps), float(expo), _stream()), "cd2d_apply")
    _lib.count()
    return out


def cd2d_adjoint(r, nrow, ncol, w=None, out=None):
    n = nrow * ncol
    _vec(r, 2 * n, "r")
    if w is not None:
        _vec(w, 2 * n, "w")
    if out is None:
        out = torch.empty(n, dtype=F64, device=r.device)
    check(lib().tb200_cd2d_adjoint(nrow, ncol, _p(r), _p(w), _p(out), _stream()), "cd2d_adjoint")
    _lib.count()
    return out


def fd1d_apply(x, out=None):
    n = _vec(x, name="x").numel()
    if out is None:
        out = torch.empty(max(n - 1, 0), dtype=F64, device=x.device)
    _vec(out, max(n - 1, 0), "out")
    check(lib().tb200_fd1d_apply(n, _p(x), _p(out), _stream()), "fd1d_apply")
    _lib.count()
    return out


def fd1d_adjoint(r, out=None):
    n = _vec(r, name="r").numel() + 1
    if out is None:
        out = torch.empty(n, dtype=F64, device=r.device)
    _vec(out, n, "out")
    check(lib().tb200_fd1d_adjoint(n, _p(r), _p(out), _stream()), "fd1d_adjoint")
    _lib.count()
    return out
