#!/usr/bin/env python
"""configs[3] of BASELINE.json at full size over the GPUs of one node (torchrun, one rank per GPU): whole solvers on the
band-sharded matrix-free operator (dist.BandShardedCT: u by angle, v by image band, NVLink peer-memory exchange) -
Hybrid_LSQR (fused sharded Golub-Kahan recurrence), CGLS and Hybrid_GMRES on A^T A (CGS2 / MGS Arnoldi), CT 2048^2 x
720 views, 50 iterations, wall clock per iteration including the host-side projected problem.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_solver_timings.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import trips_b200 as tb  # noqa: E402
import trips_oracle as O  # noqa: E402
from trips_b200.dist import BandComm, BandShardedCT  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nx, views, iters = int(os.environ.get("NX", 2048)), int(os.environ.get("VIEWS", 720)), 50
    comm = BandComm()
    A = BandShardedCT(nx, views, device=dev)
    lo, hi = A.band
    xt_full = O.shepp_logan(nx).reshape(-1, 1)
    xt = xt_full[lo:hi]
    b_loc = A.apply_dev(torch.from_numpy(xt.ravel().copy()).to(dev))
    g = torch.Generator(device=dev)
    g.manual_seed(7 + rank)
    noise = torch.randn(b_loc.numel(), dtype=torch.float64, device=dev, generator=g)
    sq = torch.stack((b_loc.square().sum(), noise.square().sum()))
    dist.all_reduce(sq)
    e = 0.01 * float(sq[0].sqrt() / sq[1].sqrt()) * noise
    b_loc = b_loc + e
    d2 = e.square().sum().reshape(1)
    dist.all_reduce(d2)
    delta = float(d2.sqrt())

    def run(name, fn):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, info = fn()
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        if rank == 0:
            rre = f" RRE {info['relError'][-1]:.4f}" if "relError" in info else ""
            print(f"[{world} GPU] {name}: {dt / iters * 1e3:.2f} ms/it{rre}", flush=True)

    for _ in range(2):  # (the first pass also pays allocator / NCCL warm-up)
        run("cfg4 Hybrid_LSQR dp 50 it, band-sharded",
            lambda: tb.Hybrid_LSQR(A, b_loc, n_iter=iters, regparam="dp", delta=delta, x_true=xt, b200_comm=comm))
    run("cfg4 CGLS 50 it, band-sharded",
        lambda: tb.CGLS(A, b_loc, np.zeros((A.shape[1], 1)), iters, 0.0, x_true=xt, b200_comm=comm))
    M, rhs = A.T @ A, A.adjoint_dev(b_loc)
    for reorth in ("cgs2", "mgs"):
        run(f"cfg4 Hybrid_GMRES on A^T A ({reorth}) 50 it, band-sharded",
            lambda: tb.Hybrid_GMRES(M, rhs, iters, regparam=1e-2, x_true=xt, b200_reorth=reorth, b200_comm=comm))
    A.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
