"""Shared plumbing for the solver drivers: lazy iterate history, device/host conversion, error tracking."""
import numpy as np
import torch

from .. import kernels as K


def single_threaded_host_blas(fn):
    """Decorator of the solver drivers: the host side of these solvers is LAPACK on (k+1) x k matrices with k <= ~100
    (SVD, QR, least squares - NumPy / SciPy, as in the reference).  A multi-threaded OpenBLAS is slower on such sizes and
    has pathological cases (measured: the discrepancy-principle solve at k = 50 takes 54 ms with 8 threads, 1.0 ms with
    one), so the BLAS pool is limited to one thread for the duration of the call (threadpoolctl; no-op if unavailable).
    The long vectors never touch the host BLAS."""
    import functools

    @functools.wraps(fn)
    def run(*args, **kwargs):
        try:
            from threadpoolctl import threadpool_limits
        except Exception:  # noqa: BLE001
            return fn(*args, **kwargs)
        with threadpool_limits(limits=1, user_api="blas"):
            return fn(*args, **kwargs)

    return run


class LazyHistory:
    """List-like `info['xHistory']`.

    The reference appends every iterate as a host (n,1) array (Hybrid_LSQR.py:107, CGLS.py:66) - 33.5 MB per
    iteration at 2048^2, which would rival the GPU iteration itself.  Here an entry is a recipe (either a device
    copy of x, or the k coefficients y with x = V y re-lifted on demand from the retained basis); indexing or
    iterating materialises (n,1) NumPy arrays exactly like the reference's list elements."""

    def __init__(self):
        self._items = []

    def append_device(self, x_dev):
        self._items.append(("x", x_dev))

    def append_lift(self, basis, k, y_host):
        self._items.append(("lift", basis, int(k), np.array(y_host, dtype=np.float64).reshape(-1)))

    def append(self, x):  # plain host array
        self._items.append(("host", np.asarray(x)))

    def __len__(self):
        return len(self._items)

    def _get(self, item):
        kind = item[0]
        if kind == "host":
            return item[1]
        if kind == "x":
            return item[1].cpu().numpy().reshape(-1, 1)
        _, basis, k, y = item
        yd = torch.from_numpy(y).to(basis.data.device)
        return K.basis_combine(basis, k, yd).cpu().numpy().reshape(-1, 1)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._get(it) for it in self._items[i]]
        return self._get(self._items[i])

    def __iter__(self):
        return (self._get(it) for it in self._items)


class ErrorTracker:
    """||x - x_true|| / ||x_true|| per iterate, computed on the device (one fused difference-norm pass).
    The reference recomputes the whole history at every iteration (Hybrid_LSQR.py:108-111); the final list is the same."""

    def __init__(self, x_true, device, comm=None):
        self.enabled = x_true is not None
        self.comm = comm
        if self.enabled:
            from ..operators import to_device_vector

            self.xt = to_device_vector(x_true, device)
            pair = K.vec_norm2(self.xt)
            if comm is not None:
                comm.sync_norm_(pair, "model")
            self.xt_norm = float(pair.cpu()[1])
            self._pairs = []

    def add(self, x_dev):
        if self.enabled:
            pair = K.vec_diffnorm2(x_dev, self.xt)
            if self.comm is not None:
                self.comm.sync_norm_(pair, "model")
            self._pairs.append(pair)

    def values(self, denominators=None):
        if not self.enabled or not self._pairs:
            return []
        num = torch.stack(self._pairs)[:, 1].cpu().numpy()
        if denominators is None:
            return [float(v) / self.xt_norm for v in num]
        return [float(v) / float(d) for v, d in zip(num, denominators)]


def need_delta(regparam, kwargs, tail):
    """The reference's argument check (Hybrid_LSQR.py:55-61, Hybrid_GMRES.py:25-31, GKS.py:29-34)."""
    delta = kwargs["delta"] if ("delta" in kwargs) else None
    dp_stop = kwargs["dp_stop"] if ("dp_stop" in kwargs) else False
    if (isinstance(regparam, str) and regparam == "dp" or dp_stop is not False) and delta is None:
        raise Exception("""A value for the noise level delta was not provided and the discrepancy principle cannot be applied.
                    Please supply a value of delta based on the estimated noise level of the problem, or choose the regularization parameter according to """ + tail)
    return delta, dp_stop


def tikhonov_projected(B, R_L, rhs, lam):
    """y = argmin ||B y - rhs||^2 + lam ||R_L y||^2 via the stacked least-squares problem, as the reference does
    (Hybrid_LSQR.py:104, Hybrid_GMRES.py:76, GKS.py:74, MMGKS.py:106)."""
    rhs = np.asarray(rhs, dtype=np.float64).reshape(-1, 1)
    stacked = np.vstack((B, np.sqrt(lam) * R_L))
    return np.linalg.lstsq(stacked, np.vstack((rhs, np.zeros((R_L.shape[0], 1)))), rcond=None)[0]


def host_column(x_dev):
    return x_dev.cpu().numpy().reshape(-1, 1)


def dev_scalar(y_host, device):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(y_host, dtype=np.float64).reshape(-1))).to(device)
