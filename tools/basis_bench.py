#!/usr/bin/env python
"""basis_dots (h = V^T w) and basis_combine (out = w - V h, with the fused norm) at the sizes of configs[2]: launch time
and achieved HBM bandwidth (8 n (k + 1 or 2) bytes), one row per access vs two."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trips_b200 as tb  # noqa: E402
from trips_b200 import _lib  # noqa: E402

K = tb.kernels


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3  # us


def main():
    lib = _lib.lib()
    hbm = float(os.environ.get("HBM_PEAK", 6550e9))
    print("n | k | kernel | 1 row/access us | 2 rows/access us | GB/s (2 rows) | frac of HBM peak")
    for n in (1 << 20, 1 << 21):
        g = torch.Generator(device="cuda").manual_seed(1)
        kmax = 56
        basis = K.Basis(n, kmax, "cuda")
        for _ in range(kmax):
            basis.next_col().copy_(torch.randn(n, dtype=torch.float64, device="cuda", generator=g))
            basis.push()
        w = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        out = torch.empty_like(w)
        nrm = K.new_pair(w.device)
        for k in (4, 8, 16, 28, 40, 54):
            h = torch.randn(k, dtype=torch.float64, device="cuda", generator=g)
            for name, fn, nbytes in (("dots", lambda: K.basis_dots(basis, k, w), 8 * n * (k + 1)),
                                     ("combine+norm", lambda: K.basis_combine(basis, k, h, w=w, sign=-1.0, out=out, norm_out=nrm),
                                      8 * n * (k + 2)),
                                     ("lift", lambda: K.basis_combine(basis, k, h, out=out), 8 * n * (k + 1))):
                res = []
                for v in (0, 1):
                    lib.tb200_basis_set_vec2(v)
                    res.append(timed(fn))
                lib.tb200_basis_set_vec2(1)
                bw = nbytes / res[1] * 1e6
                print(f"{n} | {k} | {name} | {res[0]:.1f} | {res[1]:.1f} | {bw / 1e9:.0f} | {bw / hbm:.2f}", flush=True)
        del basis


if __name__ == "__main__":
    main()
