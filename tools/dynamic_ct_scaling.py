#!/usr/bin/env python
"""configs[4]: MMGKS with space-time TV on dynamic CT (256^2 x 64 frames, 12 interleaved angles per frame),
block-diagonal operator sharded by time frame with one-frame halos.  Run under torchrun (or plain python for N = 1):
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dynamic_ct_scaling.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import trips_b200 as tb  # noqa: E402
import trips_oracle as O  # noqa: E402
from trips_b200.dist import FrameComm, shard_frames  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    comm = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        comm = FrameComm()
    nx, nt, per, iters = 256, 64, 12, 50
    th = O.ct_angles(nt * per)
    lo, hi = shard_frames(nt, world, rank)
    frames = [th[t::nt] for t in range(lo, hi)]
    A = tb.BlockDiagCT(nx, frames)
    L = tb.SpaceTimeDerivative(nx, nx, hi - lo, has_next=(world > 1 and rank < world - 1))
    base = torch.from_numpy(O.shepp_logan(nx).ravel()).cuda()
    xt = torch.cat([base * (1 + 0.3 * np.sin(2 * np.pi * t / nt)) for t in range(lo, hi)])
    bt = A.apply_dev(xt)
    g = torch.Generator(device="cuda")
    g.manual_seed(7 + rank)
    noise = torch.randn(bt.numel(), dtype=torch.float64, device="cuda", generator=g)
    nb, nn = torch.stack([bt.dot(bt), noise.dot(noise)])
    if comm:
        s = torch.stack([nb, nn])
        comm.allreduce_(s)
        nb, nn = s
    e = 0.01 * float(nb.sqrt() / nn.sqrt()) * noise
    delta = float(nn.sqrt()) * 0.01 * float(nb.sqrt() / nn.sqrt())
    b = bt + e
    kw = dict(pnorm=2, qnorm=1, projection_dim=1, regparam="dp", delta=delta, epsilon=0.1, x_true=xt, b200_history="none")
    if comm:
        kw["b200_comm"] = comm
    tb.MMGKS(A, b, L, n_iter=3, **kw)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, info = tb.MMGKS(A, b, L, n_iter=iters, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"dynamic CT {nx}^2 x {nt} frames, MMGKS space-time TV, {world} GPU(s): {dt / iters * 1e3:.2f} ms/it "
              f"({iters / dt:.1f} it/s), local nnz {A.nnz:.3e}, RRE {info['relError'][-1]:.4f}, lambda {float(info['regParam']):.3e}")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
