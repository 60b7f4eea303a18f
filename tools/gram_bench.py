#!/usr/bin/env python
"""Launch time of the double-double weighted Gram pass (csrc/gram.cu) at the sizes of configs[2] (m = 1024^2 rows for AV,
2 * 1024^2 for LV) against its two bounds: fp64 issue (10 fp64 instructions per multiply-add of the double-double
accumulation, 16 multiply-adds per 4 x 4 block and row) and HBM (8 m K bytes); bulk-copy pipeline vs plain staging."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trips_b200 as tb  # noqa: E402
from trips_b200 import _lib  # noqa: E402

K = tb.kernels


def timed(fn, reps=10):
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
    peak = float(os.environ.get("FP64_PEAK", 18.3e12))
    hbm = float(os.environ.get("HBM_PEAK", 6550e9))
    print("m | k | weighted | 4x4 plain us | 4x4 bulk us | 2x2 bulk us | default us | fp64-bound us (useful entries) | "
          "hbm-bound us | frac of max(bound), default")
    for m in (1 << 20, 1 << 21):
        g = torch.Generator(device="cuda").manual_seed(1)
        kmax = 56
        basis = K.Basis(m, kmax, "cuda")
        for _ in range(kmax):
            basis.next_col().copy_(torch.randn(m, dtype=torch.float64, device="cuda", generator=g))
            basis.push()
        b = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
        w = torch.rand(m, dtype=torch.float64, device="cuda", generator=g) + 0.5
        for k in (4, 8, 16, 20, 28, 32, 36, 40, 44, 48, 54):
            for wt in (None, w):
                ext, ew = ((b,), (0,)) if wt is None else ((b, b), (0, 1))

                def launch():
                    Kt = k + len(ext)
                    ws = K.Workspace.get(b.device).gram(Kt)
                    Ghi = torch.empty((Kt, Kt), dtype=torch.float64, device="cuda")
                    Glo = torch.empty_like(Ghi)
                    import ctypes
                    e = (ctypes.c_void_p * len(ext))(*[x.data_ptr() for x in ext])
                    f = (ctypes.c_int * len(ext))(*ew)
                    rc = lib.tb200_weighted_gram(m, k, basis.data.data_ptr(), m, wt.data_ptr() if wt is not None else None,
                                                 len(ext), e, f, Ghi.data_ptr(), Glo.data_ptr(), ws.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream)
                    assert rc == 0

                res = []
                lib.tb200_gram_set_block(4)
                for bulk in (0, 1):
                    lib.tb200_gram_set_bulk(bulk)
                    res.append(timed(launch))
                lib.tb200_gram_set_bulk(1)
                lib.tb200_gram_set_block(2)
                res.append(timed(launch))
                lib.tb200_gram_set_block(0)
                res.append(timed(launch))
                if wt is None:  # panel pass: the last basis column + the extra
                    def panel():
                        Kt = k + len(ext)
                        ws = K.Workspace.get(b.device).gram(Kt)
                        G = torch.zeros((2, Kt, Kt), dtype=torch.float64, device="cuda")
                        import ctypes
                        e = (ctypes.c_void_p * len(ext))(*[x.data_ptr() for x in ext])
                        f = (ctypes.c_int * len(ext))(*ew)
                        rc = lib.tb200_weighted_gram_panel(m, k, basis.data.data_ptr(), m, None, len(ext), e, f, k - 1,
                                                           G[0].data_ptr(), G[1].data_ptr(), ws.data_ptr(),
                                                           torch.cuda.current_stream().cuda_stream)
                        assert rc == 0
                    tp = timed(panel)
                    print(f"{m} | {k} | panel (last column + extra) | - | {tp:.1f} | - | "
                          f"{8 * m * (k + 1) / hbm * 1e6:.1f} | {8 * m * (k + 1) / hbm * 1e6 / tp:.2f}", flush=True)
                Kt = k + len(ext)
                t_fp = Kt * (Kt + 1) / 2 * 10 * m / peak * 1e6  # 10 fp64 instructions per needed entry and row
                t_hbm = 8 * m * (Kt + (wt is not None)) / hbm * 1e6
                print(f"{m} | {k} | {wt is not None} | {res[0]:.1f} | {res[1]:.1f} | {res[2]:.1f} | {res[3]:.1f} | {t_fp:.1f} | {t_hbm:.1f} | "
                      f"{max(t_fp, t_hbm) / res[3]:.2f}", flush=True)
        del basis


if __name__ == "__main__":
    main()
