"""N > 1 path on CPU: world_size-2 `gloo` run of the row-sharded Golub-Kahan step (trips_b200.dist.DistGKState).

The product's arithmetic backend is CUDA only; here the distributed ALGEBRA and the torch.distributed plumbing are
exercised with a NumPy stand-in backend that lives in this test (it implements the same six vector operations on
CPU tensors).  Checked: round-robin angle sharding, the n-vector all-reduce + scalar all-reduce per step, and
agreement of the sharded factors with the oracle's single-process Golub-Kahan on the full matrix."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import trips_oracle as O


class NumpyBackend:
    """Same interface as trips_b200.dist.CudaBackend, on CPU float64 tensors."""

    def empty(self, n, like):
        return torch.empty(n, dtype=torch.float64)

    def zeros(self, n, like):
        return torch.zeros(n, dtype=torch.float64)

    def apply(self, op, x, out, coef=None, z=None, norm_out=None):
        y = op @ x.numpy()
        if z is not None:
            y = y - float(coef) * z.numpy()
        out.copy_(torch.from_numpy(y))
        if norm_out is not None:
            norm_out[0] = float(y @ y)
            norm_out[1] = float(np.sqrt(y @ y))
        return out

    def adjoint(self, op, u, out):
        out.copy_(torch.from_numpy(op.T @ u.numpy()))
        return out

    def axpy_norm(self, a, x, y, out, norm_out, sign):
        r = y.numpy() + sign * (float(a) * x.numpy())
        out.copy_(torch.from_numpy(r))
        norm_out[0] = float(r @ r)
        norm_out[1] = float(np.sqrt(r @ r))
        return out

    def norm2(self, x, out):
        v = x.numpy()
        out[0] = float(v @ v)
        out[1] = float(np.sqrt(v @ v))
        return out

    def div(self, x, d, out):
        out.copy_(x / float(d))
        return out

    def sqrt_(self, pair):
        pair[1] = float(np.sqrt(float(pair[0])))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nx, views, steps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from trips_b200.dist import DistGKState, shard_angles

        theta = O.ct_angles(views)
        mine = shard_angles(views, world, rank)
        A_loc = O.ct_matrix(nx, theta[mine])
        x_true = O.shepp_logan(nx).reshape(-1)
        b_loc = A_loc @ x_true + 0.01 * np.random.default_rng(100 + rank).standard_normal(A_loc.shape[0])
        st = DistGKState(A_loc, torch.from_numpy(b_loc), steps, backend=NumpyBackend())
        assert st.distributed
        for _ in range(steps):
            st.step()
        V = torch.stack(st.V[:steps]).numpy()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), B=st.B_host(), V=V, b=b_loc, angles=mine,
                 U=torch.stack(st.U[:steps + 1]).numpy())
    finally:
        dist.destroy_process_group()


def test_shard_angles_round_robin_and_inverse_permutation():
    from trips_b200.dist import gather_sinogram_order, shard_angles

    parts = [shard_angles(10, 4, r) for r in range(4)]
    assert [p.tolist() for p in parts] == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    n_det = 3
    order = gather_sinogram_order(10, 4, n_det)
    local = np.concatenate([(p[:, None] * n_det + np.arange(n_det)).reshape(-1) for p in parts])  # global ray ids, rank-major
    assert np.array_equal(local[order], np.arange(10 * n_det))


def test_dist_gk_world2_gloo_matches_single_process_oracle(tmp_path):
    nx, views, steps, world = 24, 16, 6, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nx, views, steps, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    # replicated quantities are bitwise identical on both ranks
    assert np.array_equal(parts[0]["B"], parts[1]["B"]) and np.array_equal(parts[0]["V"], parts[1]["V"])
    # assemble the global problem in angle-major order and run the oracle on one process
    n_det = O.ct_num_detectors(nx)
    A = O.ct_matrix(nx, O.ct_angles(views))
    b = np.empty(views * n_det)
    U = np.empty((steps + 1, views * n_det))
    for p in parts:
        rows = (p["angles"][:, None] * n_det + np.arange(n_det)).reshape(-1)
        b[rows] = p["b"]
        U[:, rows] = p["U"]
    Uo, Bo, Vo = O.golub_kahan(A, b, steps)
    assert np.allclose(parts[0]["B"], Bo, rtol=1e-10, atol=1e-12)
    assert np.allclose(parts[0]["V"].T, Vo, atol=1e-9) and np.allclose(U.T, Uo, atol=1e-9)


def _comm_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from trips_b200.dist import FrameComm, shard_frames

        comm = FrameComm()
        assert (comm.first, comm.last) == (rank == 0, rank == world - 1)
        lo, hi = shard_frames(7, world, rank)
        frames = torch.arange(lo, hi, dtype=torch.float64).repeat_interleave(4)  # 4 "pixels" per frame, value = frame id
        nxt = comm.halo_from_next(frames[:4])
        prv = comm.halo_from_prev(frames[-4:])
        pair = torch.tensor([float(rank + 1), 0.0], dtype=torch.float64)
        comm.sync_norm_(pair)
        gathered = comm.allgather(torch.full((2, 2), float(rank), dtype=torch.float64))
        np.savez(os.path.join(out_dir, f"comm{rank}.npz"), lo=lo, hi=hi, nxt=-1 if nxt is None else nxt.numpy(),
                 prv=-1 if prv is None else prv.numpy(), pair=pair.numpy(), gathered=gathered.numpy())
    finally:
        dist.destroy_process_group()


def test_frame_comm_halos_and_reductions_world3_gloo(tmp_path):
    """One-frame halos in both directions, scalar all-reduce and all-gather of the frame-sharded (dynamic CT) path."""
    world = 3
    mp.spawn(_comm_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"comm{r}.npz") for r in range(world)]
    bounds = [(int(p["lo"]), int(p["hi"])) for p in parts]
    assert bounds[0][0] == 0 and bounds[-1][1] == 7 and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
    for r, p in enumerate(parts):
        if r < world - 1:
            assert np.array_equal(p["nxt"], np.full(4, float(bounds[r + 1][0])))  # next rank's FIRST frame
        else:
            assert p["nxt"] == -1
        if r > 0:
            assert np.array_equal(p["prv"], np.full(4, float(bounds[r - 1][1] - 1)))  # previous rank's LAST frame
        else:
            assert p["prv"] == -1
        assert p["pair"][0] == 6.0 and p["pair"][1] == np.sqrt(6.0)
        assert np.array_equal(p["gathered"][:, 0, 0], np.arange(world, dtype=float))


def _row_comm_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from trips_b200.dist import FrameComm, RowComm

        rc, fc = RowComm(), FrameComm()
        out = {}
        for name, comm in (("row", rc), ("frame", fc)):
            for space in ("data", "model", "reg"):
                pair = torch.tensor([float(rank + 1), -1.0], dtype=torch.float64)
                comm.sync_norm_(pair, space)
                h = comm.sum_(torch.tensor([1.0, 2.0], dtype=torch.float64), space)
                out[f"{name}_{space}"] = np.concatenate((pair.numpy(), h.numpy(), [comm.total(10 + rank, space)]))
        np.savez(os.path.join(out_dir, f"rc{rank}.npz"), **out)
    finally:
        dist.destroy_process_group()


def test_row_comm_sums_only_the_sharded_space_world2_gloo(tmp_path):
    """RowComm (static CT, rows by angle): only data-space reductions cross ranks; model / regulariser space are
    replicated and must be left alone.  FrameComm: every space is split."""
    world = 2
    mp.spawn(_row_comm_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        p = np.load(tmp_path / f"rc{r}.npz")
        summed = np.array([3.0, np.sqrt(3.0), 2.0, 4.0, 21.0])
        alone = np.array([r + 1.0, -1.0, 1.0, 2.0, 10.0 + r])
        assert np.array_equal(p["row_data"], summed)
        assert np.array_equal(p["row_model"], alone) and np.array_equal(p["row_reg"], alone)
        for space in ("data", "model", "reg"):
            assert np.array_equal(p[f"frame_{space}"], summed)


class _FakeProjector:
    """CPU stand-in with the interface dist.adjoint_allreduce uses: row bands of a 'back-projection' (here: a dense
    matrix-vector product restricted to the pixels of the band)."""

    def __init__(self, At, nx, ny):
        self.At, self.nx, self.ny = At, nx, ny
        self.calls = []

    def backproject_rows(self, u, out, r0, r1):
        self.calls.append((r0, r1))
        out[r0 * self.nx:r1 * self.nx] = self.At[r0 * self.nx:r1 * self.nx] @ u
        return out


class _FakeOp:
    def __init__(self, proj):
        self.projector = proj

    def adjoint_dev(self, u, out=None):
        out[:] = self.projector.At @ u
        return out


def _bands_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from trips_b200.dist import adjoint_allreduce

        nx, ny, m = 5, 22, 7
        g = torch.Generator().manual_seed(100 + rank)
        At = torch.randn(nx * ny, m, dtype=torch.float64, generator=g)
        u = torch.randn(m, dtype=torch.float64, generator=g)
        res = {}
        for bands in (1, 4, 6):
            proj = _FakeProjector(At, nx, ny)
            out = torch.zeros(nx * ny, dtype=torch.float64)
            adjoint_allreduce(_FakeOp(proj), u, out, None, bands=bands)
            res[f"out{bands}"] = out.numpy()
            res[f"calls{bands}"] = np.array(proj.calls, dtype=np.int64).reshape(-1, 2)
        np.savez(os.path.join(out_dir, f"bands{rank}.npz"), part=(At @ u).numpy(), **res)
    finally:
        dist.destroy_process_group()


def test_banded_adjoint_allreduce_world2_gloo(tmp_path):
    """dist.adjoint_allreduce: back-projection in row bands with one asynchronous all-reduce per band gives the same
    sum as one product + one all-reduce; the bands tile the image rows exactly, on multiples of four rows."""
    world = 2
    mp.spawn(_bands_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"bands{r}.npz") for r in range(world)]
    want = parts[0]["part"] + parts[1]["part"]
    for p in parts:
        for bands in (1, 4, 6):
            assert np.array_equal(p[f"out{bands}"], want), bands
        assert p["calls1"].size == 0  # one band = the plain path (operator's own adjoint)
        calls = p["calls4"]
        assert calls[0, 0] == 0 and calls[-1, 1] == 22 and np.array_equal(calls[1:, 0], calls[:-1, 1])
        assert all(c % 4 == 0 for c in calls[:, 0])
        assert p["calls6"].size == 0  # fewer than four rows per band: falls back to the plain path


# ---- band-sharded layout (dist.BandShardedCT / ShardedGKState): the partition algebra on CPU ---------------------------

def _band_worker(rank, world, port, nx, ny, views, steps, out_dir):
    """NumPy statement of ShardedGKState's data movement: u split by angle (round robin), v by image band, the gathered
    sinogram grouped by owner rank with padded chunks and addressed through per-angle row offsets (geom slot 5), the
    gathered image addressed by global pixel; every output element computed by ONE rank from the whole input."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from trips_b200.dist import band_rows, shard_angles

        n_det = O.ct_num_detectors(nx)
        theta = O.ct_angles(views)
        A = O.ct_matrix(nx, theta, ny=ny)  # every rank can evaluate any row / column: the operator is matrix-free
        AT = A.T.tocsr()
        mine = shard_angles(views, world, rank)
        L = -(-views // world)
        rows = (mine[:, None] * n_det + np.arange(n_det)[None, :]).reshape(-1)
        lo, hi = band_rows(ny, world, rank)
        b = np.random.default_rng(5).standard_normal(views * n_det)
        offs = ((np.arange(views) % world) * L + np.arange(views) // world) * n_det  # first row of angle a, gathered layout
        gathered_index = (offs[:, None] + np.arange(n_det)[None, :]).reshape(-1)    # angle-major order -> gathered layout

        def gather_u(u_loc):  # what the peer stores of the forward projector's epilogue assemble on every rank
            chunk = torch.zeros(L * n_det, dtype=torch.float64)
            chunk[:u_loc.size] = torch.from_numpy(u_loc)
            parts = [torch.empty_like(chunk) for _ in range(world)]
            dist.all_gather(parts, chunk)
            return torch.cat(parts).numpy()[gathered_index]  # the back-projector reads angle a at offs[a]

        def gather_v(v_band):  # ... and of the back-projector's epilogue (bands are unequal: pad to the largest)
            sizes = [(band_rows(ny, world, r)[1] - band_rows(ny, world, r)[0]) * nx for r in range(world)]
            buf = torch.zeros(max(sizes), dtype=torch.float64)
            buf[:v_band.size] = torch.from_numpy(v_band)
            parts = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(parts, buf)
            return np.concatenate([p.numpy()[:s] for p, s in zip(parts, sizes)])

        with O.reductions("exact"):
            u_full = gather_u(b[rows])
            beta0 = O._norm(u_full)
            u_full = u_full / beta0
            U, V, al, be = [u_full[rows]], [], [], []
            v_prev, beta_prev = None, 0.0
            for _ in range(steps):
                vt = (AT[lo * nx:hi * nx] @ u_full)  # my band of A^T u: each pixel sums over ALL angles in order
                if v_prev is not None:
                    vt = vt - beta_prev * v_prev
                vt_full = gather_v(vt)
                alpha = O._norm(vt_full)  # the mailbox all-reduce returns the correctly rounded exact total
                v_full = vt_full / alpha
                v_prev = v_full[lo * nx:hi * nx]
                ut = A[rows] @ v_full - alpha * U[-1]  # my angles of A v
                ut_full = gather_u(ut)
                beta_prev = O._norm(ut_full)
                u_full = ut_full / beta_prev
                U.append(u_full[rows])
                V.append(v_prev)
                al.append(alpha)
                be.append(beta_prev)
        np.savez(os.path.join(out_dir, f"b{rank}.npz"), U=np.array(U).T, V=np.array(V).T, al=np.array(al), be=np.array(be),
                 rows=rows, band=np.array([lo * nx, hi * nx]))
    finally:
        dist.destroy_process_group()


def test_band_sharded_partition_reproduces_single_process_golub_kahan_bitwise(tmp_path):
    nx, ny, views, steps, world = 20, 14, 9, 6, 2  # odd view count (padded chunk), band edge off a multiple of 4
    mp.spawn(_band_worker, args=(world, _free_port(), nx, ny, views, steps, str(tmp_path)), nprocs=world, join=True)
    n_det = O.ct_num_detectors(nx)
    A = O.ct_matrix(nx, O.ct_angles(views), ny=ny)
    b = np.random.default_rng(5).standard_normal(views * n_det)
    with O.reductions("exact"):
        U1, S1, V1 = O.golub_kahan(A, b, steps)
    for r in range(world):
        p = np.load(tmp_path / f"b{r}.npz")
        lo, hi = p["band"]
        assert np.array_equal(p["al"], np.diag(S1)) and np.array_equal(p["be"], np.diag(S1, -1))
        assert np.array_equal(p["U"], U1[p["rows"]]) and np.array_equal(p["V"], V1[lo:hi])
