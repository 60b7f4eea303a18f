#!/usr/bin/env python
"""Small driver for `ncu`: a few launches of the matrix-free projectors (and the stored SELL SpMVs) on a quarter of
the headline problem's angles.  Usage: ncu --set full -k regex:backproject|spmv_sell ... python tools/mf_ncu_target.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trips_b200 as tb  # noqa: E402

nx, views = 2048, 720
sub = np.arange(views)[::4]
mf = tb.ParallelBeamCT(nx, views, angle_subset=sub, layout="implicit")
m, n = mf.shape
x = torch.randn(n, dtype=torch.float64, device="cuda")
u = torch.randn(m, dtype=torch.float64, device="cuda")
pair = torch.zeros(2, dtype=torch.float64, device="cuda")
for _ in range(2):
    mf.adjoint_dev(u, norm_out=pair)
    mf.apply_dev(x, norm_out=pair)
torch.cuda.synchronize()
