#!/usr/bin/env python
"""Time the SELL-32-4 SpMV (parity build) on angle subsets - near-vertical / diagonal / near-horizontal rays - next to
the CSR tree kernel: the experiment behind the gather-cost discussion in DESIGN.md."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trips_b200 as tb  # noqa: E402


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
    for name, sub in subsets.items():
        A = tb.ParallelBeamCT(nx, views, angle_subset=sub, layout="both")
        m, n = A.shape
        x = torch.randn(n, dtype=torch.float64, device="cuda")
        u = torch.randn(m, dtype=torch.float64, device="cuda")
        y = torch.empty(m, dtype=torch.float64, device="cuda")
        z = torch.empty(n, dtype=torch.float64, device="cuda")
        gb = 12 * A.nnz / 1e9
        from trips_b200 import _lib
        res = []
        for mode in (0, 1, 2):
            _lib.check(_lib.lib().tb200_spmv_set_variant(8 * mode))
            res.append((timeit(lambda: A.apply_dev(x, out=y)), timeit(lambda: A.adjoint_dev(u, out=z))))
        _lib.check(_lib.lib().tb200_spmv_set_variant(0))
        (tA, tT) = res[0]
        modes = " ".join(f"[m{i}: A {gb / a * 1e3:5.0f} AT {gb / t * 1e3:5.0f}]" for i, (a, t) in enumerate(res))
        tr = A.with_order("tree")
        tA2 = timeit(lambda: tr.apply_dev(x, out=y))
        tT2 = timeit(lambda: tr.adjoint_dev(u, out=z))
        print(f"{name:22s} nnz {A.nnz:.2e} pad A {A.A_sell.stored / A.nnz - 1:.3f} | SELL seq: A {gb / tA * 1e3:6.0f} GB/s  AT "
              f"{gb / tT * 1e3:6.0f} GB/s {modes} | CSR tree: A {gb / tA2 * 1e3:6.0f}  AT {gb / tT2 * 1e3:6.0f}", flush=True)
        del A, tr
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
