"""Multi-GPU sharding for static CT: rows of A by projection angle, one process per GPU (torchrun / torch.distributed).

The reference has no parallelism of any kind; this layer is new.

Partition: angle a goes to rank a mod G (round robin: nnz per angle varies with
|cos|+|sin|, contiguous blocks would be imbalanced).  Rank g holds A_g (its angles' rows), the explicit transpose
A_g^T, b_g and its slice U_g of the left basis.  One Golub-Kahan step is
    z_g = A_g^T u_g                      local SpMV (gather form, deterministic)
    z   = sum_g z_g                      ONE all-reduce of an n-vector over NVLink (33.5 MB at 2048^2)
    v   = z - beta v_prev ; alpha = ||v||; v /= alpha      replicated, bitwise identical on every rank
    u_g = A_g v - alpha u_g              local SpMV with fused recurrence and local ||u_g||^2
    beta^2 = sum_g ||u_g||^2             one scalar all-reduce
There is no other data-path communication: u-space vectors never leave their rank.

The arithmetic is delegated to a small backend object so the distributed algebra can be exercised on CPU with the
gloo backend in tests (tests/test_dist_gloo.py supplies a NumPy stand-in); the product backend below is the CUDA
kernels.  The package itself has no CPU path.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import kernels as K
from .kernels import F64


def shard_angles(views, world_size, rank):
    """Angle indices owned by `rank`: round robin over ranks."""
    return np.arange(rank, views, world_size)


def gather_sinogram_order(views, world_size, n_det):
    """Permutation that maps the concatenation of the ranks' local sinograms back to angle-major order."""
    order = np.concatenate([shard_angles(views, world_size, r) for r in range(world_size)])
    inv = np.empty(views, dtype=np.int64)
    inv[order] = np.arange(views)
    return (inv[:, None] * n_det + np.arange(n_det)[None, :]).reshape(-1)


def shard_frames(nt, world_size, rank):
    """Contiguous block of time frames owned by `rank` (contiguous so the time derivative needs neighbours only)."""
    lo = rank * nt // world_size
    hi = (rank + 1) * nt // world_size
    return lo, hi


class _Comm:
    """What the solvers ask of a communicator.  Vectors live in one of three spaces - 'data' (rows of A), 'model'
    (columns of A) and 'reg' (rows of L) - and a communicator says which of them are split over the ranks:
    reductions (norms, dots, Gram matrices) over a split space are summed across ranks, the others are already
    replicated and must NOT be summed."""

    SHARDED = ()
    frames = False  # True: difference operators need one-frame halos (FrameComm)

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._bufs = {}

    def is_sharded(self, space):
        return space in self.SHARDED

    def allreduce_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def sum_(self, t, space):
        """Sum a vector of partial reductions taken over `space`."""
        return self.allreduce_(t) if self.is_sharded(space) else t

    def sync_norm_(self, pair, space="data"):
        """pair[0] = local sum of squares over `space` -> global; pair[1] = its square root."""
        if self.is_sharded(space):
            dist.all_reduce(pair[0:1], op=dist.ReduceOp.SUM, group=self.group)
            pair[1:2] = torch.sqrt(pair[0:1])
        return pair

    def total(self, count, space):
        """Global length of a `space` vector whose local length is `count`."""
        if not self.is_sharded(space):
            return int(count)
        t = torch.tensor([int(count)], dtype=torch.int64, device=self._device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    def _device(self):
        return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")

    def allgather(self, t):
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t.contiguous(), group=self.group)
        return torch.stack(parts)


class RowComm(_Comm):
    """Static CT, rows of A sharded by projection angle: data-space vectors (b, u,
    residuals, AV) are local slices; model-space vectors (x, v, V) and regulariser-space vectors (L x, LV) are
    replicated and bitwise identical on every rank.  The only large exchange is the sum of the partial back-projections
    A_g^T u_g (ShardedRowsOperator.adjoint_dev); everything else is a scalar, a k-vector or a k x k Gram matrix."""

    SHARDED = ("data",)


class FrameComm(_Comm):
    """Communication of the frame-sharded (dynamic CT) solvers: the operator is block diagonal
    over time frames, so A, A^T, the bases and every vector are local to the rank that owns the frames; what crosses
    ranks is (i) scalars, k-vectors and k x k Gram matrices (all-reduce / all-gather of a few hundred doubles) and
    (ii) ONE FRAME of halo per application of the temporal difference operator or its adjoint (0.5 MB at 256^2)."""

    SHARDED = ("data", "model", "reg")
    frames = True

    def __init__(self, group=None, peer=None):
        """peer=True: halos travel through NVLink peer memory (tb200_halo_exchange) instead of NCCL send / recv; default:
        peer memory on the NCCL backend (env TB200_FRAME_HALO=nccl switches back), torch.distributed elsewhere."""
        super().__init__(group)
        if peer is None:
            peer = dist.get_backend(group) == "nccl" and os.environ.get("TB200_FRAME_HALO", "peer") != "nccl"
        self.peer, self._pc, self._pc_n = bool(peer), None, 0

    def _peer_exchange(self, send_prev, send_next, n):
        if self._pc is None or self._pc_n < n:  # (collective: every rank gets here in the same exchange)
            if self._pc is not None:
                self._pc.destroy()
            self._pc, self._pc_n = PeerComm({"halo": 4 * n}, group=self.group), n
        return self._pc.halo_exchange("halo", send_prev, send_next, n=n)

    @property
    def first(self):
        return self.rank == 0

    @property
    def last(self):
        return self.rank == self.world - 1

    def _buf(self, key, like, n):
        b = self._bufs.get(key)
        if b is None or b.numel() != n or b.device != like.device:
            b = self._bufs[key] = torch.empty(n, dtype=like.dtype, device=like.device)
        return b

    def _exchange(self, send, send_to, recv_from, key):
        ops, buf = [], None
        if send_to is not None:
            ops.append(dist.P2POp(dist.isend, send.contiguous(), send_to, self.group))
        if recv_from is not None:
            buf = self._buf(key, send, send.numel())
            ops.append(dist.P2POp(dist.irecv, buf, recv_from, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return buf

    def halo_from_next(self, my_first_frame):
        """Every rank hands its first frame to the previous rank; returns the next rank's first frame (None on the last)."""
        if self.peer:
            return self._peer_exchange(None if self.first else my_first_frame.contiguous(), None, my_first_frame.numel())[1]
        return self._exchange(my_first_frame, None if self.first else self.rank - 1,
                              None if self.last else self.rank + 1, "next")

    def halo_from_prev(self, my_last_block):
        """Every rank hands a block to the next rank; returns the previous rank's block (None on the first)."""
        if self.peer:
            return self._peer_exchange(None, None if self.last else my_last_block.contiguous(), my_last_block.numel())[0]
        return self._exchange(my_last_block, None if self.last else self.rank + 1,
                              None if self.first else self.rank - 1, "prev")


# Image bands per sharded back-projection: with > 1 the all-reduce of band c runs under the kernel of band c+1.
# Measured at 2 GPUs: no gain (6.31 vs 6.25 ms per step - four smaller launches and NCCL's CTAs competing for SMs cost
# what the hidden reduction saves), so the default is one band; TB200_ADJOINT_BANDS=4 enables the overlap.
ADJOINT_BANDS = int(os.environ.get("TB200_ADJOINT_BANDS", "1"))


def adjoint_allreduce(op, u, out, group=None, bands=None):
    """out = sum over ranks of A_g^T u_g.  With the matrix-free CT operator the image is back-projected in horizontal
    bands and every band is all-reduced (NCCL, asynchronously on its own stream) while the next one is computed, so
    only the last band's all-reduce is exposed; other operators do one product and one all-reduce.  The sums are the
    same numbers either way (an all-reduce is element-wise)."""
    proj = getattr(op, "projector", None)
    bands = ADJOINT_BANDS if bands is None else bands
    if proj is None or bands <= 1 or proj.ny < 4 * bands:
        op.adjoint_dev(u, out=out)
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out
    nx, ny = proj.nx, proj.ny
    edges = [(ny * c // bands) // 4 * 4 for c in range(bands)] + [ny]
    pending = []
    for c in range(bands):
        r0, r1 = edges[c], edges[c + 1]
        proj.backproject_rows(u, out, r0, r1)
        pending.append(dist.all_reduce(out[r0 * nx:r1 * nx], op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in pending:
        w.wait()
    return out


class CudaBackend:
    """Vector operations of the distributed step on the CUDA kernels (1-D float64 CUDA tensors)."""

    def adjoint_allreduce(self, op, u, out, group):
        return adjoint_allreduce(op, u, out, group)

    def empty(self, n, like):
        return torch.empty(n, dtype=F64, device=like.device)

    def zeros(self, n, like):
        return torch.zeros(n, dtype=F64, device=like.device)

    def apply(self, op, x, out, coef=None, z=None, norm_out=None):
        return op.apply_dev(x, out=out, coef=coef, z=z, norm_out=norm_out)

    def adjoint(self, op, u, out):
        return op.adjoint_dev(u, out=out)

    def axpy_norm(self, a, x, y, out, norm_out, sign):
        return K.vec_axpy(a, x, y, out=out, norm_out=norm_out, sign=sign)

    def norm2(self, x, out):
        return K.vec_norm2(x, out=out)

    def div(self, x, d, out):
        return K.vec_div(x, d, out=out)

    def sqrt_(self, pair):
        pair[1:2] = torch.sqrt(pair[0:1])


class DistGKState:
    """Golub-Kahan bidiagonalisation with rows sharded by angle.  `A_local` is this rank's CSROperator, `b_local`
    its part of the right-hand side.  With world_size == 1 (or no process group) it degenerates to GKState."""

    exchange_name = "nccl all-reduce of the partial back-projections"

    def __init__(self, A_local, b_local, kmax, backend=None, group=None, exchange=None):
        self.A, self.be = A_local, backend or CudaBackend()
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        m, n = A_local.shape
        be = self.be
        self.kmax = kmax
        self.U = [be.empty(m, b_local) for _ in range(kmax + 1)]
        self.V = [be.empty(n, b_local) for _ in range(kmax)]
        self.alpha = [be.zeros(2, b_local) for _ in range(kmax)]
        self.beta = [be.zeros(2, b_local) for _ in range(kmax)]
        self.beta0 = be.zeros(2, b_local)
        self.k = 0
        be.norm2(b_local, self.beta0)
        self._allreduce_sq(self.beta0)
        be.div(b_local, self.beta0[1:2], self.U[0])

    def _allreduce_sq(self, pair):
        """pair[0] holds a local sum of squares: make it global and refresh pair[1] = sqrt."""
        if self.distributed:
            dist.all_reduce(pair[0:1], op=dist.ReduceOp.SUM, group=self.group)
            self.be.sqrt_(pair)

    def step(self):
        k, be = self.k, self.be
        if k >= self.kmax:
            raise RuntimeError("DistGKState capacity exceeded")
        u_k, v = self.U[k], self.V[k]
        if self.distributed and hasattr(be, "adjoint_allreduce"):
            be.adjoint_allreduce(self.A, u_k, v, self.group)  # z = sum_g A_g^T u_g, reduction overlapped with the kernel
        else:
            be.adjoint(self.A, u_k, v)  # z_g = A_g^T u_g
            if self.distributed:
                dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.group)  # z = sum_g z_g
        if k == 0:
            be.norm2(v, self.alpha[k])
        else:
            be.axpy_norm(self.beta[k - 1][1:2], self.V[k - 1], v, v, self.alpha[k], -1.0)  # v = z - beta v_prev
        be.div(v, self.alpha[k][1:2], v)
        u = self.U[k + 1]
        be.apply(self.A, v, u, coef=self.alpha[k][1:2], z=u_k, norm_out=self.beta[k])  # u_g = A_g v - alpha u_g
        self._allreduce_sq(self.beta[k])
        be.div(u, self.beta[k][1:2], u)
        self.k += 1

    def bench_hooks(self, timed):
        """bench.py: wrap (timed = a decorator) or restore (None) the two operator launches of a step."""
        proj = getattr(self.A, "projector", None)
        if timed is None:
            for obj, name in getattr(self, "_hooked", []):
                delattr(obj, name)  # instance attribute shadows the class method
            self._hooked = []
            return
        self._hooked = []
        if proj is not None:  # A^T first (adjoint), then A: the order of the events list
            proj.backproject = timed(proj.backproject)
            proj.forward = timed(proj.forward)
            self._hooked = [(proj, "backproject"), (proj, "forward")]

    def scalars_host(self):
        k = self.k
        to = lambda t: float(t[1])  # noqa: E731
        return to(self.beta0), np.array([to(a) for a in self.alpha[:k]]), np.array([to(b) for b in self.beta[:k]])

    def B_host(self):
        _, al, be = self.scalars_host()
        k = al.size
        B = np.zeros((k + 1, k))
        B[np.arange(k), np.arange(k)] = al
        B[np.arange(1, k + 1), np.arange(k)] = be
        return B


# ---- static CT behind the solver interface -------------------------------------------------------------------------

from .operators import LinearOperator, ParallelBeamCT  # noqa: E402  (operators does not import this module)


class ShardedRowsOperator(LinearOperator):
    """This rank's row block A_g of a row-sharded operator, presented to the solvers as the whole operator:
    shape (m_g, n); `apply_dev` is local (x replicated -> local rows), `adjoint_dev` sums the ranks' partial
    back-projections with ONE all-reduce of an n-vector before the fused recurrence/norm epilogue, so its result
    is replicated and bitwise identical on every rank.  Pass the matching RowComm as `b200_comm=` to the solver;
    `A.T @ A` of this operator is a replicated square operator and needs no communicator (Hybrid_GMRES)."""

    fused = True

    def __init__(self, local_op, comm):
        super().__init__(local_op.shape, local_op.device)
        self.local, self.comm = local_op, comm

    def apply_dev(self, x, out=None, coef=None, z=None, norm_out=None):
        return self.local.apply_dev(x, out=out, coef=coef, z=z, norm_out=norm_out)  # norm_out: LOCAL sum of squares

    def adjoint_dev(self, u, out=None, coef=None, z=None, norm_out=None):
        if out is None:
            out = torch.empty(self.shape[1], dtype=F64, device=self.device)
        adjoint_allreduce(self.local, u, out, self.comm.group)
        if z is not None:
            K.vec_axpy(coef, z, out, out=out, norm_out=norm_out, sign=-1.0)
        elif norm_out is not None:
            K.vec_norm2(out, out=norm_out)
        return out


def sharded_ct(nx, views, comm, **kwargs):
    """(operator, row index) of this rank's share of ParallelBeamCT(nx, views): angles round robin over the ranks.
    `rows` selects the rank's entries of a full angle-major sinogram: b_local = b_full[rows]."""
    mine = shard_angles(views, comm.world, comm.rank)
    op = ParallelBeamCT(nx, views, angle_subset=mine, **kwargs)
    rows = (mine[:, None] * op.n_det + np.arange(op.n_det)[None, :]).reshape(-1)
    return ShardedRowsOperator(op, comm), rows


# ---- NVLink peer-memory exchange (csrc/comm.cu): no NCCL on the data path --------------------------------------------

import ctypes  # noqa: E402

from . import _lib  # noqa: E402
from ._lib import c_ptr, check, lib  # noqa: E402


class _DeviceMemory:
    """Zero-copy torch view of raw device memory (the arena is cudaMalloc'ed by libtripsb200, not by torch)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class PeerComm:
    """One CUDA-IPC arena per rank, mapped into every other rank (tb200_comm_*).  `regions` maps a name to a length in
    doubles; every rank must pass the same dict: the layout is the same everywhere, so `peer_ptr(name, r)` is where rank
    r keeps that region.  torch.distributed (any backend) is used ONCE, to exchange the 64-byte IPC handles."""

    BOX_ALPHA, BOX_BETA, BOX_MISC, BOX_HALO = 0, 1, 2, 3

    def __init__(self, regions, group=None, device=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        L = lib()
        if self.world > int(L.tb200_comm_max_ranks()):
            raise ValueError(f"PeerComm supports up to {L.tb200_comm_max_ranks()} ranks")
        off = int(L.tb200_comm_mailbox_bytes())
        self.offsets, self.counts = {}, {}
        for name, count in regions.items():
            self.offsets[name], self.counts[name] = off, int(count)
            off += (8 * int(count) + 255) // 256 * 256
        self.bytes = off
        hb = int(L.tb200_comm_handle_bytes())
        handle = (ctypes.c_ubyte * hb)()
        comm = ctypes.c_void_p()
        check(L.tb200_comm_init(self.rank, self.world, self.bytes, ctypes.byref(comm), handle), "comm_init")
        self._comm = comm
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        blob = (ctypes.c_ubyte * (hb * self.world)).from_buffer_copy(b"".join(handles))
        check(L.tb200_comm_connect(self._comm, blob), "comm_connect")
        self.base = [int(L.tb200_comm_arena(self._comm, r)) for r in range(self.world)]
        self._epoch = {}
        self._scratch = torch.zeros(2, dtype=F64, device=self.device)
        self.barrier()  # every arena is mapped everywhere before anyone stores into a peer

    def local(self, name):
        return torch.as_tensor(_DeviceMemory(self.base[self.rank] + self.offsets[name], self.counts[name]), device=self.device)

    def peer_ptr(self, name, r, offset_doubles=0):
        return self.base[r] + self.offsets[name] + 8 * int(offset_doubles)

    def peers_of(self, name, offset_doubles=0):
        """ctypes array of the device pointers of `name` (+ offset) in every OTHER rank's arena."""
        ptrs = [self.peer_ptr(name, r, offset_doubles) for r in range(self.world) if r != self.rank]
        return (ctypes.c_void_p * max(len(ptrs), 1))(*ptrs), len(ptrs)

    def _next_epoch(self, box):
        e = self._epoch.get(box, 0) + 1
        self._epoch[box] = e
        return e

    def allreduce_dd(self, box, partials, npart, out, nval=1):
        """out[2j] = sum over ranks and partials of value j (double-double, rank order), out[2j+1] = its square root."""
        check(lib().tb200_comm_allreduce_dd(self._comm, int(box), self._next_epoch(box), K._p(partials), int(npart), int(nval),
                                            K._p(out), K._stream()), "comm_allreduce_dd")
        _lib.count(1)
        return out

    def allreduce_scale(self, box, partials, npart, x, out, keep_begin, keep, pair_out):
        """pair_out = (sum over ranks and partials, its square root); out = x / sqrt; keep = out[keep_begin : +len(keep)].
        tb200_comm_allreduce_dd + tb200_comm_scale in one launch."""
        check(lib().tb200_comm_allreduce_scale(self._comm, int(box), self._next_epoch(box), K._p(partials), int(npart),
                                               x.numel(), K._p(x), K._p(out), int(keep_begin), keep.numel(), K._p(keep),
                                               K._p(pair_out), K._stream()), "comm_allreduce_scale")
        _lib.count(1)

    def barrier(self):
        self.allreduce_dd(self.BOX_MISC, None, 0, self._scratch)

    def push(self, name, src, offset_doubles=0, ranks=None):
        mask = sum(1 << r for r in (range(self.world) if ranks is None else ranks))
        check(lib().tb200_comm_push(self._comm, mask, self.offsets[name] + 8 * int(offset_doubles), K._p(src), src.numel(),
                                    K._stream()), "comm_push")
        _lib.count(1)

    def halo_exchange(self, name, send_prev, send_next, box=None, n=None):
        """One-block halo exchange with the neighbouring ranks through the arena region `name` (>= 4 * n doubles):
        send_prev goes to rank - 1, send_next to rank + 1 (either may be None at the ends); returns (recv_prev, recv_next) =
        what rank - 1 sent as its send_next and what rank + 1 sent as its send_prev (None at the ends).
        tb200_halo_exchange: peer stores + one mailbox epoch, no NCCL."""
        if n is None:
            n = (send_prev if send_prev is not None else send_next).numel()
        n = int(n)
        if 4 * n > self.counts[name]:
            raise ValueError(f"arena region {name!r} holds {self.counts[name]} doubles, the halo needs {4 * n}")
        box = self.BOX_HALO if box is None else box
        recv_prev = torch.empty(n, dtype=F64, device=self.device) if self.rank > 0 else None
        recv_next = torch.empty(n, dtype=F64, device=self.device) if self.rank + 1 < self.world else None
        check(lib().tb200_halo_exchange(self._comm, int(box), self._next_epoch(box), self.offsets[name], K._p(send_prev),
                                        K._p(send_next), n, K._p(recv_prev), K._p(recv_next), K._p(self._scratch), K._stream()),
              "halo_exchange")
        _lib.count(2)
        return recv_prev, recv_next

    def destroy(self):
        if self._comm is not None:
            torch.cuda.synchronize(self.device)
            lib().tb200_comm_destroy(self._comm)
            self._comm = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:  # noqa: BLE001
            pass


FUSE_ALLREDUCE_SCALE = os.environ.get("TB200_FUSE_ALLREDUCE_SCALE", "1") != "0"


def band_rows(ny, world, rank):
    """Image rows [lo, hi) owned by `rank`: equal bands, boundaries on multiples of 4 (the back-projector's CTA tile)."""
    edge = lambda r: ny if r >= world else (ny * r // world) // 4 * 4  # noqa: E731
    return edge(rank), edge(rank + 1)


class BandComm(_Comm):
    """Communicator of the solvers on a BandShardedCT operator: data space split by angle, model space by image band;
    scalar reductions (norms, dots, k-vectors) over either are summed through torch.distributed."""

    SHARDED = ("data", "model")


class BandShardedCT(LinearOperator):
    """The matrix-free parallel-beam operator split over the G GPUs of a node: THIS RANK'S ANGLES x THIS
    RANK'S IMAGE BAND.  shape = (m_loc, n_band): apply_dev maps the rank's band of x to the rank's rows of A x,
    adjoint_dev the rank's rows of u to the rank's band of A^T u; the other ranks' parts travel over NVLink peer memory
    (PeerComm), never through NCCL.  Because the projectors are matrix-free, each output element is computed on ONE rank
    from the whole input in the single-GPU order: products are bit-identical to ParallelBeamCT's.
    `rows` / `band` say which entries of a full sinogram / image are this rank's.  `gk_state(b_local, kmax)` is the fused
    Golub-Kahan recurrence (ShardedGKState); the solvers take `b200_comm=BandComm()` next to this operator."""

    fused = True

    def __init__(self, nx, views, ny=None, n_det=None, angles=None, group=None, device=None):
        from .operators import ct_angles, ct_num_detectors, default_device

        dev = torch.device(device) if device is not None else default_device()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        G, r = self.world, self.rank
        self.nx, self.ny = int(nx), int(nx if ny is None else ny)
        self.n_det = ct_num_detectors(nx) if n_det is None else int(n_det)
        theta = ct_angles(views) if angles is None else np.asarray(angles, dtype=np.float64)
        self.theta = theta
        self.views = len(theta)
        self.mine = shard_angles(self.views, G, r)
        self.L = -(-self.views // G)  # angles per rank, padded
        self.n_loc_ang = len(self.mine)
        self.m_loc = self.n_loc_ang * self.n_det
        self.n = self.nx * self.ny
        self.m_pad = G * self.L * self.n_det
        self.row_lo, self.row_hi = band_rows(self.ny, G, r)
        self.n_band = (self.row_hi - self.row_lo) * self.nx
        super().__init__((self.m_loc, self.n_band), dev)
        K._lib.require_device()
        # geometry: all angles (back-projection; slot 5 = first row of the angle in the gathered sinogram) and mine (forward)
        cos_t, sin_t = torch.from_numpy(np.cos(theta)).to(dev), torch.from_numpy(np.sin(theta)).to(dev)
        self.geom_all = torch.zeros(max(6 * self.views, 2), dtype=F64, device=dev)
        check(lib().tb200_ct_geometry(self.views, K._p(cos_t), K._p(sin_t), K._p(self.geom_all), K._stream()), "ct_geometry")
        a = np.arange(self.views)
        offs = torch.from_numpy((((a % G) * self.L + a // G) * self.n_det).astype(np.int64)).to(dev)
        self.geom_all.view(torch.int64)[:6 * self.views].view(self.views, 6)[:, 5] = offs
        self.geom_loc = torch.zeros(max(6 * self.n_loc_ang, 2), dtype=F64, device=dev)
        idx = torch.from_numpy(self.mine).to(dev)
        cm, sm = cos_t[idx].contiguous(), sin_t[idx].contiguous()
        check(lib().tb200_ct_geometry(self.n_loc_ang, K._p(cm), K._p(sm), K._p(self.geom_loc), K._stream()), "ct_geometry")
        _lib.count(2)
        self.cta_order = None  # heaviest-first CTA list of the forward projector for this rank's angles
        if os.environ.get("TB200_CT_FORWARD_ORDER", "lpt") == "lpt":
            self.cta_order = K.forward_cta_order(self.nx, self.ny, self.n_det, cm, sm, dev)
        # every rank's copies of a gathered image ("vt") and a gathered sinogram ("ut") live in the peer-mapped arena
        self.comm = PeerComm({"vt": self.n, "ut": self.m_pad}, group=group, device=dev)
        self.vt, self.ut = self.comm.local("vt"), self.comm.local("ut")
        self.ut.zero_()  # padding rows of ranks with fewer angles stay zero
        self.comm.barrier()
        nws = max(int(lib().tb200_ct_backproject_workspace_len(self.nx, self.ny)),
                  int(lib().tb200_ct_forward_rays_workspace_len(self.n_det, max(self.n_loc_ang, 1))),
                  int(lib().tb200_reduce_workspace_len()))
        self.ws = torch.zeros(nws, dtype=F64, device=dev)
        self.my_chunk = r * self.L * self.n_det  # where my angles sit in a gathered sinogram
        self.vt_peers = self.comm.peers_of("vt")
        self.ut_peers = self.comm.peers_of("ut", self.my_chunk)
        self.npart = ctypes.c_int64(0)

    @property
    def rows(self):
        """Indices of this rank's entries in a full angle-major sinogram: b_local = b_full[rows]."""
        return (self.mine[:, None] * self.n_det + np.arange(self.n_det)[None, :]).reshape(-1)

    @property
    def band(self):
        """(lo, hi): this rank's entries of a full row-major image are x_full[lo:hi]."""
        return self.row_lo * self.nx, self.row_hi * self.nx

    @property
    def nnz(self):
        return None

    def gk_state(self, b_local, kmax):
        return ShardedGKState(self, b_local, kmax)

    # -- the two sharded projector launches --------------------------------------------------------------------------
    def launch_backproject(self, u_full, out_full, peers, coef_dev, z_band, coef_host=0.0):
        ptrs, n = peers if peers is not None else (None, 0)
        check(lib().tb200_ct_backproject_sharded_f64(
            self.nx, self.ny, self.row_lo, self.row_hi, self.n_det, self.views, K._p(self.geom_all), K._p(u_full), K._p(out_full),
            ptrs, n, float(coef_host), K._p(coef_dev), K._p(z_band), K._p(self.ws), ctypes.byref(self.npart), K._stream()),
            "ct_backproject_sharded")
        _lib.count(1)

    def launch_forward(self, x_full, out_rows, peers, coef_dev, z_rows, coef_host=0.0):
        ptrs, n = peers if peers is not None else (None, 0)
        check(lib().tb200_ct_forward_rays_sharded_f64(
            self.nx, self.ny, self.n_det, self.n_loc_ang, K._p(self.geom_loc), K._p(x_full), K._p(out_rows), ptrs, n,
            float(coef_host), K._p(coef_dev), K._p(z_rows), K._p(self.ws), ctypes.byref(self.npart), K._p(self.cta_order),
            K._stream()), "ct_forward_rays_sharded")
        _lib.count(1)

    def _finish(self, norm_out):
        if norm_out is not None:  # LOCAL sum of squares; the solver's communicator sums it over the ranks
            check(lib().tb200_reduce_finalize(K._p(self.ws), self.npart.value, K._p(norm_out), K._stream()), "reduce_finalize")
            _lib.count(1)

    def _coef(self, coef, z):
        if z is None:
            return 0.0, None
        return (0.0, coef) if isinstance(coef, torch.Tensor) else (float(coef), None)

    # -- operator interface (solver level): gather by peer pushes, then the local projector -----------------------------
    def apply_dev(self, x, out=None, coef=None, z=None, norm_out=None):
        """rows_g(A x) from the band x: every rank pushes its band into all ranks' copy of the image first."""
        K._vec(x, self.n_band, "x (this rank's image band)")
        out = torch.empty(self.m_loc, dtype=F64, device=self.device) if out is None else K._vec(out, self.m_loc, "out")
        self.comm.barrier()  # every rank is done reading the previous gathered image
        self.comm.push("vt", x, self.row_lo * self.nx)
        self.comm.barrier()
        ch, cd = self._coef(coef, z)
        self.launch_forward(self.vt, out, None, cd, z, ch)
        self._finish(norm_out)
        return out

    def adjoint_dev(self, u, out=None, coef=None, z=None, norm_out=None):
        """band_g(A^T u) from the rank's rows of u: every rank pushes its rows into all ranks' copy of the sinogram first."""
        K._vec(u, self.m_loc, "u (this rank's sinogram rows)")
        self.comm.barrier()
        self.comm.push("ut", u, self.my_chunk)
        self.comm.barrier()
        ch, cd = self._coef(coef, z)
        full = torch.empty(self.n, dtype=F64, device=self.device)  # the kernel addresses by global pixel number
        self.launch_backproject(self.ut, full, None, cd, z, ch)
        self._finish(norm_out)
        band = full[self.row_lo * self.nx:self.row_hi * self.nx]
        if out is None:
            return band.clone()
        K._vec(out, self.n_band, "out").copy_(band)
        return out

    def close(self):
        self.comm.destroy()


class ShardedGKState:
    """Golub-Kahan bidiagonalisation of the matrix-free parallel-beam operator over G GPUs.

    u-space (sinogram) is split by projection angle (round robin), v-space (image) by row band, and so are the retained
    bases U and V.  The two projectors are matrix-free, so each can be restricted to ANY subset of its outputs:
      back-projection:  rank g computes ITS BAND of A^T u from the whole u   (every pixel sums over all angles in order)
      forward:          rank g computes ITS ANGLES of A v from the whole v
    No partial sums ever cross GPUs - there is no floating-point reduction of vectors, so the factors are the same bits as
    on one GPU.  What a step exchanges is the two vectors themselves: each projector's epilogue stores its part of the
    result straight into every rank's copy over NVLink peer memory while the rest of its grid is still computing
    (tb200_ct_*_sharded_f64), and the squared norms travel as double-double partials through the mailboxes of
    tb200_comm_allreduce_dd, which is also the barrier that publishes the stores.  Per step and rank: 6 kernels, no NCCL.
        K1 back-project band (+ peer stores)   K2 alpha = all-reduce(dd)   K3 v = vt / alpha (own band -> V)
        K4 forward own angles (+ peer stores)  K5 beta  = all-reduce(dd)   K6 u = ut / beta  (own rows -> U)"""

    exchange_name = "peer-to-peer stores from the projector epilogues + mailbox all-reduce (NVLink, no NCCL on the data path)"

    def __init__(self, op, b_local, kmax):
        if not isinstance(b_local, torch.Tensor):
            from .operators import to_device_vector

            b_local = to_device_vector(b_local, op.device)
        self.op, self.comm = op, op.comm
        dev = op.device
        for name in ("nx", "ny", "n_det", "views", "n", "m_loc", "m_pad", "n_band", "row_lo", "row_hi", "my_chunk", "vt", "ut", "ws"):
            setattr(self, name, getattr(op, name))
        if b_local.numel() != self.m_loc:
            raise ValueError(f"rank {op.rank} owns {op.n_loc_ang} angles: b_local must have {self.m_loc} entries")
        self.v_full = torch.empty(self.n, dtype=F64, device=dev)
        self.u_full = torch.zeros(self.m_pad, dtype=F64, device=dev)
        self.U = K.Basis(self.m_loc, kmax + 1, dev)
        self.V = K.Basis(self.n_band, max(kmax, 1), dev)
        self.alpha = torch.zeros((kmax + 1, 2), dtype=F64, device=dev)
        self.beta = torch.zeros((kmax + 1, 2), dtype=F64, device=dev)
        self.beta0 = torch.zeros(2, dtype=F64, device=dev)
        self._npart = op.npart
        # u_1 = b / ||b||: the norm through the mailboxes, my rows pushed to every rank's copy
        check(lib().tb200_vec_dot_partials(self.m_loc, K._p(b_local), None, K._p(self.ws), ctypes.byref(self._npart), K._stream()),
              "vec_dot_partials")
        self.comm.allreduce_dd(PeerComm.BOX_BETA, self.ws, self._npart.value, self.beta0)
        self.comm.barrier()  # every rank is done with whatever used the gathered sinogram before
        self.comm.push("ut", b_local.contiguous(), self.my_chunk)
        self.comm.barrier()
        self._scale_u(self.beta0)
        _lib.count(1)

    @property
    def k(self):
        return self.V.k

    def _scale_u(self, pair):
        check(lib().tb200_comm_scale(self.m_pad, K._p(self.ut), K._p(pair[1:2]), K._p(self.u_full), self.my_chunk, self.m_loc,
                                     K._p(self.U.next_col()), K._stream()), "comm_scale")
        self.U.push()
        _lib.count(1)

    def backproject(self, k):
        """K1: my band of vt = A^T u_full - beta_{k-1} v_{k-1}, stored into every rank's vt; partial norms -> ws."""
        self.op.launch_backproject(self.u_full, self.vt, self.op.vt_peers, self.beta[k - 1, 1:2] if k else None,
                                   self.V.col(k - 1) if k else None)

    def forward(self, k):
        """K4: my angles of ut = A v_full - alpha_k u_k, stored into every rank's ut; partial norms -> ws."""
        self.op.launch_forward(self.v_full, self.ut[self.my_chunk:], self.op.ut_peers, self.alpha[k, 1:2], self.U.col(k))

    def step(self):
        k = self.V.k
        if k + 1 >= self.alpha.shape[0]:
            raise RuntimeError("ShardedGKState capacity exceeded")
        self.backproject(k)
        if FUSE_ALLREDUCE_SCALE:  # K2 + K3 and K5 + K6 as one launch each
            self.comm.allreduce_scale(PeerComm.BOX_ALPHA, self.ws, self._npart.value, self.vt, self.v_full, self.row_lo * self.nx,
                                      self.V.next_col(), self.alpha[k])
            self.V.push()
            self.forward(k)
            self.comm.allreduce_scale(PeerComm.BOX_BETA, self.ws, self._npart.value, self.ut, self.u_full, self.my_chunk,
                                      self.U.next_col(), self.beta[k])
            self.U.push()
            return
        self.comm.allreduce_dd(PeerComm.BOX_ALPHA, self.ws, self._npart.value, self.alpha[k])
        check(lib().tb200_comm_scale(self.n, K._p(self.vt), K._p(self.alpha[k, 1:2]), K._p(self.v_full), self.row_lo * self.nx,
                                     self.n_band, K._p(self.V.next_col()), K._stream()), "comm_scale")
        self.V.push()
        _lib.count(1)
        self.forward(k)
        self.comm.allreduce_dd(PeerComm.BOX_BETA, self.ws, self._npart.value, self.beta[k])
        self._scale_u(self.beta[k])

    def bench_hooks(self, timed):
        """bench.py: wrap (timed = a decorator) or restore (None) the two projector launches of a step."""
        if timed is None:
            for name in ("backproject", "forward"):
                self.__dict__.pop(name, None)
            return
        self.backproject = timed(self.backproject)  # A^T first, then A: the order of bench.py's event list
        self.forward = timed(self.forward)

    def host_step(self, hu_k, hv_prev, beta_prev, hu_out, hv_out):
        """One golub_kahan_update (decompositions.py:230-255) with HOST bases, this rank's part of it: hu_k = my rows of
        u_k, hv_prev = my band of v_{k-1} (None at the first step), beta_prev = S[k-1, k-2]; the new v band / u rows are
        written to the pinned host tensors hv_out / hu_out.  Returns (alpha, beta).  Per call: H2D of u_k and v_{k-1}, the
        rows of u_k pushed to every rank over NVLink, K1..K6 as in step(), D2H of v, u and the two scalars."""
        u_k = self.U.data[0]
        u_k.copy_(hu_k, non_blocking=True)
        self.comm.barrier()  # every rank is done reading the gathered sinogram of the previous call
        self.comm.push("ut", u_k, self.my_chunk)
        self.comm.barrier()
        zb = None
        if hv_prev is not None:
            zb = self.V.data[0]
            zb.copy_(hv_prev, non_blocking=True)
            self.beta[0, 1:2].fill_(float(beta_prev))
        self.op.launch_backproject(self.ut, self.vt, self.op.vt_peers, self.beta[0, 1:2] if zb is not None else None, zb)
        self.comm.allreduce_dd(PeerComm.BOX_ALPHA, self.ws, self._npart.value, self.alpha[0])
        vb = self.V.data[1] if self.V.kmax > 1 else self.V.data[0]
        check(lib().tb200_comm_scale(self.n, K._p(self.vt), K._p(self.alpha[0, 1:2]), K._p(self.v_full), self.row_lo * self.nx,
                                     self.n_band, K._p(vb), K._stream()), "comm_scale")
        hv_out.copy_(vb, non_blocking=True)
        self.op.launch_forward(self.v_full, self.ut[self.my_chunk:], self.op.ut_peers, self.alpha[0, 1:2], u_k)
        self.comm.allreduce_dd(PeerComm.BOX_BETA, self.ws, self._npart.value, self.beta[1])
        ub = self.U.data[1]
        check(lib().tb200_comm_scale(self.m_pad, K._p(self.ut), K._p(self.beta[1, 1:2]), K._p(self.u_full), self.my_chunk,
                                     self.m_loc, K._p(ub), K._stream()), "comm_scale")
        hu_out.copy_(ub, non_blocking=True)
        _lib.count(4)
        sc = torch.stack((self.alpha[0, 1], self.beta[1, 1])).cpu().numpy()  # synchronises: the host copies are complete
        return float(sc[0]), float(sc[1])

    def scalars_host(self):
        k = self.V.k
        packed = torch.cat((self.beta0[1:2], self.alpha[:k, 1], self.beta[:k, 1])).cpu().numpy()
        return packed[0], packed[1:1 + k], packed[1 + k:]

    def B_host(self):
        _, al, be = self.scalars_host()
        k = al.size
        B = np.zeros((k + 1, k))
        B[np.arange(k), np.arange(k)] = al
        B[np.arange(1, k + 1), np.arange(k)] = be
        return B

    def close(self):
        self.op.close()


def sharded_parity_check(st, nx, views, layout, b_local, steps=10):
    """bench.py, N > 1, outside the timed region: the first `steps` (alpha, beta) of the sharded state `st` against the
    SAME problem on ONE GPU (rank 0 rebuilds the whole operator, gathers the right-hand side in angle-major order and
    repeats the steps).  Returns the maximal relative deviation (rank 0; None elsewhere)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = b_local.device
    _, al, be = st.scalars_host()
    n_det = b_local.numel() // len(shard_angles(views, world, rank))
    parts = [torch.empty(len(shard_angles(views, world, r)) * n_det, dtype=F64, device=dev) for r in range(world)]
    dist.all_gather(parts, b_local.contiguous())
    out = None
    if rank == 0:
        order = torch.from_numpy(gather_sinogram_order(views, world, n_det)).to(dev)
        b_full = torch.cat(parts)[order]
        del parts
        from .decompositions import GKState

        A1 = ParallelBeamCT(nx, views, device=dev, layout=layout)
        s1 = GKState(A1, b_full, steps)
        for _ in range(steps):
            s1.step()
        _, al1, be1 = s1.scalars_host()
        k = min(steps, len(al))
        d = float(max(np.max(np.abs(al[:k] - al1[:k]) / al1[:k]), np.max(np.abs(be[:k] - be1[:k]) / be1[:k])))
        out = {"what": f"first {k} (alpha, beta) of the {world}-GPU run vs the same problem on one GPU",
               "alpha_beta_max_rel_dev_vs_1gpu": d, "bitwise": bool(d == 0.0), "ok": bool(d < 1e-9)}
        del A1, s1
        torch.cuda.empty_cache()
    dist.barrier()
    return out
