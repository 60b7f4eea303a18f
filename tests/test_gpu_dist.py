"""Multi-GPU (NCCL) parity: the frame-sharded MMGKS / GKS for dynamic CT against the single-GPU solver on the same
problem.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a 1-GPU box."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

import trips_oracle as O

pytestmark = pytest.mark.gpu

NX, NT, PER = 32, 8, 5


def _problem():
    th = O.ct_angles(NT * PER)
    frames = [th[t::NT] for t in range(NT)]
    A = sp.block_diag([O.ct_matrix(NX, f) for f in frames], format="csr")
    base = O.shepp_logan(NX)
    xt = np.concatenate([(base * (1 + 0.1 * t)).ravel() for t in range(NT)]).reshape(-1, 1)
    b, delta = O.add_noise(A @ xt, 0.01, np.random.default_rng(1))
    return frames, xt, b, float(delta)


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import trips_b200 as tb
        from trips_b200.dist import FrameComm, shard_frames

        frames, xt, b, delta = _problem()
        lo, hi = shard_frames(NT, world, rank)
        n_det = O.ct_num_detectors(NX)
        A_loc = tb.BlockDiagCT(NX, frames[lo:hi])
        L_loc = tb.SpaceTimeDerivative(NX, NX, hi - lo, has_next=rank < world - 1)
        b_loc = b[lo * PER * n_det:hi * PER * n_det]
        xt_loc = xt[lo * NX * NX:hi * NX * NX]
        comm = FrameComm()
        x, info = tb.MMGKS(A_loc, b_loc, L_loc, pnorm=2, qnorm=1, projection_dim=1, n_iter=20, regparam="dp", delta=delta,
                           epsilon=0.1, x_true=xt_loc, b200_comm=comm)
        xg, ig = tb.GKS(A_loc, b_loc, L_loc, projection_dim=2, n_iter=12, regparam=0.5, b200_comm=comm)
        np.savez(os.path.join(out_dir, f"x{rank}.npz"), x=x, lam=np.array(info["regParam_history"], dtype=float),
                 rre=np.array(info["relError"]), res=np.array(info["Residual"]), xg=xg)
    finally:
        dist.destroy_process_group()


def test_frame_sharded_mmgks_matches_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import trips_b200 as tb

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"x{r}.npz") for r in range(world)]
    frames, xt, b, delta = _problem()
    A = tb.BlockDiagCT(NX, frames)
    L = tb.SpaceTimeDerivative(NX, NX, NT)
    x1, i1 = tb.MMGKS(A, b, L, pnorm=2, qnorm=1, projection_dim=1, n_iter=20, regparam="dp", delta=delta, epsilon=0.1, x_true=xt)
    xg1, _ = tb.GKS(A, b, L, projection_dim=2, n_iter=12, regparam=0.5)
    x2 = np.concatenate([p["x"] for p in parts])
    xg2 = np.concatenate([p["xg"] for p in parts])
    rel = lambda a, c: np.linalg.norm(a - c) / np.linalg.norm(c)  # noqa: E731
    print("frame-sharded vs single GPU: MMGKS", rel(x2, x1), "GKS", rel(xg2, xg1))
    assert rel(x2, x1) < 1e-9 and rel(xg2, xg1) < 1e-9
    assert np.array_equal(parts[0]["lam"], parts[1]["lam"])  # replicated host problem: identical on both ranks
    assert np.allclose(parts[0]["lam"], np.array(i1["regParam_history"], dtype=float), rtol=1e-8)
    assert np.allclose(parts[0]["rre"], i1["relError"], rtol=1e-8)
    assert np.allclose(parts[0]["res"], i1["Residual"], rtol=1e-6)
