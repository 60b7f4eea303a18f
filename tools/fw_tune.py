#!/usr/bin/env python
"""Sweep the tuning knobs of the ray-driven forward projector (tb200_ct_forward_set_tuning) on the headline geometry and
check that the product does not depend on them (same bits)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trips_b200 as tb  # noqa: E402
from trips_b200 import _lib  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(3):
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
    ap.add_argument("--every", type=int, default=1, help="angle subsampling (8 = one rank of eight)")
    a = ap.parse_args()
    sub = np.arange(0, a.views, a.every)
    op = tb.ParallelBeamCT(a.nx, a.views, angle_subset=sub, layout="implicit")
    m, n = op.shape
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    y = torch.empty(m, dtype=torch.float64, device="cuda")
    pair = torch.zeros(2, dtype=torch.float64, device="cuda")
    ref = None
    for minb in (8, 82, 84, 4):  # 82 / 84: 64 registers with 2 / 4 warps per CTA forced
        for run_tan in ((5.0, 7.9) if minb != 8 else (3.0, 5.0, 7.9)):
            _lib.check(_lib.lib().tb200_ct_forward_set_tuning(run_tan, minb))
            pr = op.projector  # the CTA list is tied to the launch plan: rebuild it after changing the knobs
            if pr.cta_order is not None or os.environ.get("TB200_CT_FORWARD_ORDER", "lpt") == "lpt":
                from trips_b200.kernels import forward_cta_order
                pr.cta_order = forward_cta_order(pr.nx, pr.ny, pr.n_det, pr.cos_t, pr.sin_t, pr.device, force=minb > 9)
            t = timeit(lambda: op.apply_dev(x, out=y, norm_out=pair))
            if ref is None:
                ref = y.clone()
            print(f"min_ctas {minb} run_tan {run_tan:4.1f}: forward {t:7.3f} ms  same bits: {bool(torch.equal(y, ref))}", flush=True)
    _lib.check(_lib.lib().tb200_ct_forward_set_tuning(7.9, 8))
    u = torch.randn(m, dtype=torch.float64, device="cuda")
    z = torch.empty(n, dtype=torch.float64, device="cuda")
    print(f"back-projection {timeit(lambda: op.adjoint_dev(u, out=z, norm_out=pair)):7.3f} ms")
    if a.every == 1:  # one rank's image band of an 8 / 4-GPU run: rows [0, ny / G) from all angles
        pr = op.projector
        for G in (8, 4, 2):
            lo, hi = pr.ny // 2 - pr.ny // (2 * G), pr.ny // 2 + pr.ny // (2 * G)
            print(f"back-projection of a 1/{G} band (all angles) {timeit(lambda: pr.backproject_rows(u, z, lo, hi)):7.3f} ms")


if __name__ == "__main__":
    main()
