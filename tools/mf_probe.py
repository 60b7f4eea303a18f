#!/usr/bin/env python
"""Time the matrix-free CT projectors (layout='implicit') at the headline size next to the stored SELL-32-4 SpMVs,
per angle class and for the full problem, and check that the two give the same bits."""
import argparse
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--views", type=int, default=720)
    ap.add_argument("--full", action="store_true", help="also the full problem (needs ~110 GB for the stored matrices)")
    ap.add_argument("--no-stored", action="store_true")
    ap.add_argument("--no-index", action="store_true", help="skip round 1's index-streaming forward projector")
    a = ap.parse_args()
    nx, views = a.nx, a.views
    th = np.arange(views)
    q = views // 8
    subsets = {"vertical(0-22deg)": th[:q], "steep(22-44deg)": th[q:2 * q - 2], "shallow(46-68deg)": th[2 * q + 2:3 * q],
               "runs(72-108deg)": th[16 * q // 5:24 * q // 5], "shallow-desc(112-134deg)": th[5 * q - 2:6 * q - 4],
               "all/4": th[::4]}
    subsets["all/8 (one of 8 ranks)"] = th[::8]
    if a.full:
        subsets["all"] = th
    for name, sub in subsets.items():
        if not a.no_index:
            ix = tb.ParallelBeamCT(nx, views, angle_subset=sub, layout="implicit", forward="index")
            xi = torch.randn(ix.shape[1], dtype=torch.float64, device="cuda")
            yi = torch.empty(ix.shape[0], dtype=torch.float64, device="cuda")
            pi = torch.zeros(2, dtype=torch.float64, device="cuda")
            tI = timeit(lambda: ix.apply_dev(xi, out=yi, norm_out=pi))
            mfr = tb.ParallelBeamCT(nx, views, angle_subset=sub, layout="implicit")
            yr = mfr.apply_dev(xi)
            print(f"{name:22s} forward index-stream {tI:7.3f} ms | rays same bits: {bool(torch.equal(yr, yi))}", flush=True)
            del ix, mfr, xi, yi, yr
            torch.cuda.empty_cache()
        mf = tb.ParallelBeamCT(nx, views, angle_subset=sub, layout="implicit")
        m, n = mf.shape
        x = torch.randn(n, dtype=torch.float64, device="cuda")
        u = torch.randn(m, dtype=torch.float64, device="cuda")
        y = torch.empty(m, dtype=torch.float64, device="cuda")
        z = torch.empty(n, dtype=torch.float64, device="cuda")
        pair = torch.zeros(2, dtype=torch.float64, device="cuda")
        tF = timeit(lambda: mf.apply_dev(x, out=y, norm_out=pair))
        tB = timeit(lambda: mf.adjoint_dev(u, out=z, norm_out=pair))
        nnz = mf.nnz
        line = (f"{name:22s} nnz {nnz:.3e} | matrix-free: A {tF:7.3f} ms ({nnz / tF / 1e6:6.1f} Gnnz/s)"
                f"  AT {tB:7.3f} ms ({nnz / tB / 1e6:6.1f} Gnnz/s)")
        if not a.no_stored and 24 * nnz < 150e9:
            y1, z1 = y.clone(), z.clone()
            st = tb.ParallelBeamCT(nx, views, angle_subset=sub, layout="sell")
            sF = timeit(lambda: st.apply_dev(x, out=y, norm_out=pair))
            sB = timeit(lambda: st.adjoint_dev(u, out=z, norm_out=pair))
            same = bool(torch.equal(y, y1)) and bool(torch.equal(z, z1))
            line += f" | stored SELL: A {sF:7.3f} ms  AT {sB:7.3f} ms | speed-up A {sF / tF:4.2f}x AT {sB / tB:4.2f}x | same bits: {same}"
            del st
        print(line, flush=True)
        del mf
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
