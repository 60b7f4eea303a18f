"""TEST INFRASTRUCTURE - loads the REAL, unmodified reference (mpasha3/trips-py).

Search order: $TRIPS_REFERENCE_ROOT, then oracle/_ref/ (the git-ignored pip install made by oracle/make_ref.sh; it
travels to the GPU box), then /root/reference (build container only).  It is used to
  (1) pin oracle/trips_oracle.py bit for bit against the reference's own functions (tests/test_oracle.py),
  (2) generate the golden vectors committed under tests/golden/ (oracle/make_golden.py),
  (3) run the reference's own solver loops on trips_b200 operators (the drop-in claim, tests/test_gpu_dropin.py),
  (4) time the reference's golub_kahan_update on the host cores (bench.py --impl reference, cpu_baseline).
Nothing under trips-py_b200/ imports this file.

The reference imports pylops, astra, matplotlib, h5py, resizeimage, PIL, requests at module level; none of the
first five is installed and there is no network.  Stand-ins are registered in sys.modules before the import:
empty modules for the ones only touched at import time, and a minimal `pylops` (LinearOperator whose products and
transposes stay LinearOperator instances - the reference tests `isinstance(R_L.T @ R_L, LinearOperator)` at
trips/utilities/reg_param/gcv.py:39 -, Identity, FunctionOperator, FirstDerivative(kind='centered'), Kronecker,
VStack).  pylops semantics are restated from its public documentation (pylops 2.x; the reference does not pin a
version: setup.py:3-12).
"""
import importlib
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    env = os.environ.get("TRIPS_REFERENCE_ROOT")
    for cand in ([env] if env else []) + [os.path.join(_HERE, "_ref"), "/root/reference"]:
        if os.path.isdir(os.path.join(cand, "trips")):
            return cand
    return env or os.path.join(_HERE, "_ref")


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "trips"))


# ---- minimal pylops ---------------------------------------------------------------------------------------------

class LinearOperator:
    def __init__(self, shape=None, dtype="float64"):
        self.shape = tuple(shape) if shape is not None else None
        self.dtype = np.dtype(dtype)

    def _matvec(self, x):
        raise NotImplementedError

    def _rmatvec(self, x):
        raise NotImplementedError

    def matvec(self, x):
        x = np.asarray(x)
        y = np.asarray(self._matvec(x.reshape(-1) if x.ndim == 1 else x))
        return y.reshape(-1) if x.ndim == 1 else y.reshape(-1, 1)

    def rmatvec(self, x):
        x = np.asarray(x)
        y = np.asarray(self._rmatvec(x.reshape(-1) if x.ndim == 1 else x))
        return y.reshape(-1) if x.ndim == 1 else y.reshape(-1, 1)

    def matmat(self, X):
        return np.stack([self.matvec(X[:, j]).reshape(-1) for j in range(X.shape[1])], axis=1)

    def rmatmat(self, X):
        return np.stack([self.rmatvec(X[:, j]).reshape(-1) for j in range(X.shape[1])], axis=1)

    def dot(self, x):
        if isinstance(x, LinearOperator):
            return _Product(self, x)
        if np.isscalar(x):
            return _Scaled(self, x)
        x = np.asarray(x)
        if x.ndim == 2 and x.shape[1] > 1:
            return self.matmat(x)
        return self.matvec(x)

    __matmul__ = dot
    __mul__ = dot

    def __rmul__(self, x):
        return _Scaled(self, x) if np.isscalar(x) else NotImplemented

    def __rmatmul__(self, x):  # ndarray @ op
        x = np.asarray(x)
        return (self.T.dot(x.T)).T

    def __add__(self, other):
        return _Sum(self, other)

    @property
    def T(self):
        return _Transposed(self)

    H = T

    def todense(self):
        return self.matmat(np.eye(self.shape[1]))


class _Transposed(LinearOperator):
    def __init__(self, op):
        super().__init__((op.shape[1], op.shape[0]), op.dtype)
        self.op = op

    def _matvec(self, x):
        return self.op.rmatvec(x)

    def _rmatvec(self, x):
        return self.op.matvec(x)


class _Product(LinearOperator):
    def __init__(self, a, b):
        super().__init__((a.shape[0], b.shape[1]), a.dtype)
        self.a, self.b = a, b

    def _matvec(self, x):
        return self.a.matvec(self.b.matvec(x))

    def _rmatvec(self, x):
        return self.b.rmatvec(self.a.rmatvec(x))


class _Scaled(LinearOperator):
    def __init__(self, a, s):
        super().__init__(a.shape, a.dtype)
        self.a, self.s = a, s

    def _matvec(self, x):
        return self.s * self.a.matvec(x)

    def _rmatvec(self, x):
        return self.s * self.a.rmatvec(x)


class _Sum(LinearOperator):
    def __init__(self, a, b):
        super().__init__(a.shape, a.dtype)
        self.a, self.b = a, b

    def _matvec(self, x):
        return self.a.matvec(x) + self.b.matvec(x)

    def _rmatvec(self, x):
        return self.a.rmatvec(x) + self.b.rmatvec(x)


class Identity(LinearOperator):
    def __init__(self, N, M=None, dtype="float64", **_):
        M = N if M is None else M
        super().__init__((N, M), dtype)

    def _matvec(self, x):
        y = np.zeros(self.shape[0], dtype=self.dtype) if x.ndim == 1 else np.zeros((self.shape[0], 1), dtype=self.dtype)
        k = min(self.shape)
        y[:k] = x[:k]
        return y

    def _rmatvec(self, x):
        y = np.zeros(self.shape[1], dtype=self.dtype) if x.ndim == 1 else np.zeros((self.shape[1], 1), dtype=self.dtype)
        k = min(self.shape)
        y[:k] = x[:k]
        return y


class FunctionOperator(LinearOperator):
    def __init__(self, f, fc, nr, nc=None, dtype="float64", **_):
        nc = nr if nc is None else nc
        super().__init__((nr, nc), dtype)
        self.f, self.fc = f, fc

    def _matvec(self, x):
        return np.squeeze(self.f(x))

    def _rmatvec(self, x):
        return np.squeeze(self.fc(x))

    def matvec(self, x):
        x = np.asarray(x)
        y = np.asarray(self._matvec(x)).reshape(-1)
        return y if x.ndim == 1 else y.reshape(-1, 1)

    def rmatvec(self, x):
        x = np.asarray(x)
        y = np.asarray(self._rmatvec(x)).reshape(-1)
        return y if x.ndim == 1 else y.reshape(-1, 1)


class FirstDerivative(LinearOperator):
    """pylops.FirstDerivative, 1-D, sampling 1, edge=False, default kind='centered' (3-point):
    y[1:-1] = (x[2:] - x[:-2]) / 2, y[0] = y[-1] = 0; result stored in the operator's dtype."""

    def __init__(self, dims, dtype="float64", kind="centered", **_):
        n = int(np.prod(dims))
        super().__init__((n, n), dtype)
        if kind != "centered":
            raise NotImplementedError(kind)

    def _matvec(self, x):
        x = x.reshape(-1)
        y = np.zeros(x.shape, self.dtype)
        y[1:-1] = (0.5 * x[2:] - 0.5 * x[:-2])
        return y

    def _rmatvec(self, x):
        x = x.reshape(-1)
        y = np.zeros(x.shape, self.dtype)
        y[:-2] -= 0.5 * x[1:-1]
        y[2:] += 0.5 * x[1:-1]
        return y


class Kronecker(LinearOperator):
    """kron(A, B) acting on row-major vec: Y = A X B^T with X reshaped (A.shape[1], B.shape[1])."""

    def __init__(self, A, B, dtype="float64"):
        super().__init__((A.shape[0] * B.shape[0], A.shape[1] * B.shape[1]), dtype)
        self.A, self.B = A, B

    def _matvec(self, x):
        X = x.reshape(self.A.shape[1], self.B.shape[1])
        Y = self.B.matmat(X.T).T
        Y = self.A.matmat(Y)
        return Y.reshape(-1)

    def _rmatvec(self, x):
        X = x.reshape(self.A.shape[0], self.B.shape[0])
        Y = self.B.rmatmat(X.T).T
        Y = self.A.rmatmat(Y)
        return Y.reshape(-1)


class VStack(LinearOperator):
    def __init__(self, ops, dtype="float64"):
        self.ops = list(ops)
        super().__init__((sum(o.shape[0] for o in self.ops), self.ops[0].shape[1]), dtype)

    def _matvec(self, x):
        return np.concatenate([o.matvec(x.reshape(-1)) for o in self.ops])

    def _rmatvec(self, x):
        x = x.reshape(-1)
        out, off = 0, 0
        for o in self.ops:
            out = out + o.rmatvec(x[off:off + o.shape[0]])
            off += o.shape[0]
        return out


class BlockDiag(LinearOperator):
    def __init__(self, ops, dtype="float64"):
        self.ops = list(ops)
        super().__init__((sum(o.shape[0] for o in self.ops), sum(o.shape[1] for o in self.ops)), dtype)

    def _matvec(self, x):
        x = x.reshape(-1)
        out, off = [], 0
        for o in self.ops:
            out.append(o.matvec(x[off:off + o.shape[1]]))
            off += o.shape[1]
        return np.concatenate(out)

    def _rmatvec(self, x):
        x = x.reshape(-1)
        out, off = [], 0
        for o in self.ops:
            out.append(o.rmatvec(x[off:off + o.shape[0]]))
            off += o.shape[0]
        return np.concatenate(out)


def _install_shims():
    if "pylops" not in sys.modules:
        pl = types.ModuleType("pylops")
        for name in ("LinearOperator", "Identity", "FunctionOperator", "FirstDerivative", "Kronecker", "VStack", "BlockDiag"):
            setattr(pl, name, globals()[name])
        sys.modules["pylops"] = pl
    for name in ("astra", "h5py", "resizeimage", "resizeimage.resizeimage", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.transforms", "cil"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:  # noqa: BLE001
                sys.modules[name] = types.ModuleType(name)
    if hasattr(sys.modules.get("resizeimage"), "__dict__") and not hasattr(sys.modules["resizeimage"], "resizeimage"):
        sys.modules["resizeimage"].resizeimage = sys.modules["resizeimage.resizeimage"]
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(mpl, "transforms"):
        mpl.transforms = sys.modules["matplotlib.transforms"]


_loaded = None


def load():
    """Import the reference package `trips` from REFERENCE_ROOT and return a namespace of its hot-path functions."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _install_shims()
    import tqdm

    # silence the progress bars of the reference loops
    _orig = tqdm.tqdm

    def _quiet(it=None, *a, **k):
        k["disable"] = True
        return _orig(it, *a, **k)

    tqdm.tqdm = _quiet
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.decompositions = importlib.import_module("trips.utilities.decompositions")
    ns.gcv = importlib.import_module("trips.utilities.reg_param.gcv")
    ns.dp = importlib.import_module("trips.utilities.reg_param.discrepancy_principle")
    ns.l_curve = importlib.import_module("trips.utilities.reg_param.l_curve")
    ns.weights = importlib.import_module("trips.utilities.weights")
    ns.phantoms = importlib.import_module("trips.utilities.phantoms")
    ns.CGLS = importlib.import_module("trips.solvers.CGLS").CGLS
    ns.Hybrid_LSQR = importlib.import_module("trips.solvers.Hybrid_LSQR").Hybrid_LSQR
    ns.Hybrid_GMRES = importlib.import_module("trips.solvers.Hybrid_GMRES").Hybrid_GMRES
    ns.GKS = importlib.import_module("trips.solvers.GKS").GKS
    ns.MMGKS = importlib.import_module("trips.solvers.MMGKS").MMGKS
    ns.Deblurring2D = importlib.import_module("trips.test_problems.Deblurring2D").Deblurring2D
    ns.pylops = sys.modules["pylops"]
    _loaded = ns
    return ns
