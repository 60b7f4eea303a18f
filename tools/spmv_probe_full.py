import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import trips_b200 as tb
from trips_b200 import _lib
from spmv_probe import timeit
A = tb.ParallelBeamCT(2048, 720, layout="sell")
print("padding A %.3f%%  AT %.3f%%" % (100*(A.A_sell.stored/A.nnz-1), 100*(A.AT_sell.stored/A.nnz-1)))
m, n = A.shape
x = torch.randn(n, dtype=torch.float64, device="cuda"); u = torch.randn(m, dtype=torch.float64, device="cuda")
y = torch.empty(m, dtype=torch.float64, device="cuda"); z = torch.empty(n, dtype=torch.float64, device="cuda")
gb = 12 * A.nnz / 1e9
for w in (0, 1, 2, 3):
    _lib.check(_lib.lib().tb200_spmv_set_variant(256 * w))
    tA = timeit(lambda: A.apply_dev(x, out=y), 4); tT = timeit(lambda: A.adjoint_dev(u, out=z), 4)
    print(f"full SELL warps/CTA {1 << w if w else 4}: A {tA:.2f} ms {gb/tA*1e3:6.0f} GB/s   AT {tT:.2f} ms {gb/tT*1e3:6.0f} GB/s", flush=True)
del A
torch.cuda.empty_cache()
A = tb.ParallelBeamCT(2048, 720, layout='csr')
tr = A.with_order("tree")
tA = timeit(lambda: tr.apply_dev(x, out=y), 4); tT = timeit(lambda: tr.adjoint_dev(u, out=z), 4)
print(f"full tree: A {tA:.2f} ms {gb/tA*1e3:6.0f}   AT {tT:.2f} ms {gb/tT*1e3:6.0f}")
