#!/usr/bin/env python
"""Time the sequential SpMV on angle subsets (near-vertical / diagonal / near-horizontal rays) under each gather
mode and tile variant: the experiment behind the L1TEX tag-cost model in DESIGN.md."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trips_b200 as tb  # noqa: E402
from trips_b200 import _lib  # noqa: E402


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    nx, views = 2048, 720
    th = np.arange(views)
    subsets = {"vertical(0-22deg)": th[:90], "diag(34-56deg)": th[135:225], "horizontal(79-101deg)": th[315:405],
               "all/4": th[::4]}
    tile = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    for name, sub in subsets.items():
        A = tb.ParallelBeamCT(nx, views, angle_subset=sub)
        m, n = A.shape
        x = torch.randn(n, dtype=torch.float64, device="cuda")
        u = torch.randn(m, dtype=torch.float64, device="cuda")
        y = torch.empty(m, dtype=torch.float64, device="cuda")
        z = torch.empty(n, dtype=torch.float64, device="cuda")
        gb = 12 * A.nnz / 1e9
        row = [f"{name:22s} nnz {A.nnz:.2e}"]
        for mode in (0, 1, 2, 3):
            _lib.check(_lib.lib().tb200_spmv_set_variant(tile + 8 * mode))
            tA = timeit(lambda: A.apply_dev(x, out=y))
            tT = timeit(lambda: A.adjoint_dev(u, out=z))
            row.append(f"mode{mode}: A {gb / tA * 1e3:6.0f} GB/s  AT {gb / tT * 1e3:6.0f} GB/s")
        tr = A.with_order("tree")
        tA = timeit(lambda: tr.apply_dev(x, out=y))
        tT = timeit(lambda: tr.adjoint_dev(u, out=z))
        row.append(f"tree: A {gb / tA * 1e3:6.0f}  AT {gb / tT * 1e3:6.0f}")
        print(" | ".join(row), flush=True)
        del A, tr
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
