"""Executable specification (CPU, NumPy) of the matrix-free back-projector's algorithm - csrc/ct_project.cu.

The kernel never sees the matrix: per (pixel, angle) it brackets the projection with the 1.5*2^52 rounding trick, tests
the two bracketing detector bins with the exact footprint predicate and adds the hits in ascending (angle, detector)
order.  The claim that this reproduces `A.T @ u` of the stored matrix BIT FOR BIT (same pattern, same values, same
summation order) is checked here step by step in NumPy, on many geometries, without a GPU; the GPU tests then check that
the CUDA code equals the same stored-matrix product (tests/test_gpu_kernels.py)."""
import numpy as np
import pytest

import trips_oracle as O

MAGIC = 6755399441055744.0  # 1.5 * 2^52


def backproject_bracket(nx, ny, n_det, theta, u):
    """NumPy transcription of ct_backproject_kernel (CHECKED path), vectorised over the pixels."""
    cx = np.arange(nx) - 0.5 * (nx - 1)
    cy = np.arange(ny) - 0.5 * (ny - 1)
    CX, CY = np.meshgrid(cx, cy)  # (ny, nx): pixel = iy*nx + ix
    dc = 0.5 * (n_det - 1)
    dcm = dc - 0.5
    acc = np.zeros((ny, nx))
    hits_per_pixel = np.zeros((ny, nx), dtype=np.int64)
    for a, th in enumerate(theta):
        c, s = np.cos(th), np.sin(th)
        hi, lo = max(abs(c), abs(s)), min(abs(c), abs(s))
        d2 = 0.5 * (hi + lo)
        with np.errstate(divide="ignore"):
            inv_hi, inv_hilo = 1.0 / hi, 1.0 / (hi * lo)
        proj = CX * c + CY * s
        w = (proj + dcm) + MAGIC
        d0 = (w - MAGIC).astype(np.int64)  # the kernel reads the low word of w; same integer
        for cand in (d0, d0 + 1):  # ascending detector order
            sd = cand - dc
            e = d2 - np.abs(sd - proj)
            inside = (cand >= 0) & (cand < n_det)
            hit = inside & (e > 0)
            with np.errstate(invalid="ignore", over="ignore"):
                slope = e * inv_hilo
                val = np.where(slope < inv_hi, slope, inv_hi)
            uu = u[a * n_det + np.clip(cand, 0, n_det - 1)]
            acc = np.where(hit, acc + val * uu, acc)
            hits_per_pixel += hit
    return acc.reshape(-1), hits_per_pixel.reshape(-1)


@pytest.mark.parametrize("nx,ny,views,n_det", [(16, 16, 12, None), (21, 13, 9, None), (24, 24, 7, 20), (10, 30, 16, 33),
                                               (32, 32, 5, 46)])
def test_bracket_back_projection_equals_the_stored_transpose_product_bitwise(nx, ny, views, n_det):
    rng = np.random.default_rng(nx * 100 + views)
    n_det = O.ct_num_detectors(nx) if n_det is None else n_det
    theta = np.concatenate((O.ct_angles(views), [np.pi / 4, 3 * np.pi / 4, 1e-9]))  # + degenerate trapezoids
    A = O.ct_matrix(nx, theta, ny=ny, n_det=n_det)
    u = rng.standard_normal(A.shape[0])
    got, hits = backproject_bracket(nx, ny, n_det, theta, u)
    want = A.T @ u  # scipy csc_matvec on the CSR matrix: per pixel, ascending (angle, detector)
    assert np.array_equal(hits, np.diff(A.tocsc().indptr))  # the bracket finds exactly the stored pattern
    assert np.array_equal(got, np.asarray(want).ravel())  # ... and the same bits


def test_integer_to_double_by_mantissa_insertion_is_exact():
    """centred_coord / the back-projector's (d - dc): 2^51 + k built by writing 2k into the mantissa of 2^51, minus
    (2^51 + (n-1)/2), equals k - (n-1)/2 exactly (what `(double)k - 0.5*(n-1)` gives)."""
    for n in (2, 3, 724, 2896, 65535):
        k = np.arange(0, n, max(n // 97, 1), dtype=np.int64)
        bits = (np.int64(0x43200000) << 32) | (k << 1)
        built = bits.view(np.float64) if bits.ndim else bits
        assert np.array_equal(built - (2251799813685248.0 + 0.5 * (n - 1)), k - 0.5 * (n - 1))
