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
        # even detector counts: dc - 1/2 is an integer, MAGIC + (dc - 1/2) is exact and the kernel folds the two additions
        w = proj + (MAGIC + dcm) if n_det % 2 == 0 else (proj + dcm) + MAGIC
        d0 = (w - MAGIC).astype(np.int64)  # the kernel reads the low word of w; same integer
        sd0 = (w - MAGIC) - dc             # the kernel's (d - dc): exact
        for cand, sd in ((d0, sd0), (d0 + 1, sd0 + 1.0)):  # ascending detector order
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


def _magic(d):
    """The host-side constants of tb200_ct_forward_f64 (spmv.cu): shift = floor(log2 d), magic = floor(2^(32+shift)/d)
    clipped to 2^32 - 1."""
    sh = 0
    while (2 << sh) <= d:
        sh += 1
    mg = (1 << (32 + sh)) // d
    return min(mg, 0xFFFFFFFF), sh


def test_index_division_by_multiply_high_with_one_fixup_is_exact():
    """The forward projector splits a column index into (row, column) of the image - or of the transposed image for
    shallow rays - with q = umulhi(c, magic) >> shift, r = c - q*d and ONE conditional correction.  Exact for every
    divisor and every index below 2^31 (checked on edge values and a seeded sample)."""
    rng = np.random.default_rng(0)
    divisors = [1, 2, 3, 5, 7, 24, 255, 256, 257, 724, 1448, 2047, 2048, 2049, 2896, 4096, 5792, 46340, 65535, 65536]
    for d in divisors:
        mg, sh = _magic(d)
        top = min((1 << 31) - 1, d * 65536 - 1)
        c = np.unique(np.concatenate((np.arange(0, min(4 * d, top) + 1), rng.integers(0, top + 1, 20000),
                                      [top, top - 1, top - d, (top // d) * d, (top // d) * d - 1]))).astype(np.uint64)
        q = ((c * np.uint64(mg)) >> np.uint64(32)) >> np.uint64(sh)
        r = c.astype(np.int64) - q.astype(np.int64) * d
        fix = r >= d
        q, r = q.astype(np.int64) + fix, r - fix * d
        assert np.array_equal(q, c.astype(np.int64) // d) and np.array_equal(r, c.astype(np.int64) % d), d
        assert fix.sum() <= c.size  # (the correction is needed at all only for some divisors)


def _sectors_per_entry(rows_cols, skips):
    """Lane-per-ray gathers: at every position j the 32 rays of a slice request one element each; count the distinct
    32-byte sectors (4 doubles) per request, summed over the slice, per stored entry."""
    sectors = entries = 0
    for s0 in range(0, len(rows_cols), 32):
        grp, sk = rows_cols[s0:s0 + 32], skips[s0:s0 + 32]
        width = max((len(c) + k for c, k in zip(grp, sk)), default=0)
        M = np.full((len(grp), width), -1, dtype=np.int64)
        for i, (c, k) in enumerate(zip(grp, sk)):
            M[i, k:k + len(c)] = c
        for j in range(width):
            col = M[:, j]
            col = col[col >= 0]
            if col.size:
                sectors += np.unique(col // 4).size
        entries += sum(len(c) for c in grp)
    return sectors / max(entries, 1)


def test_alignment_and_transposed_addressing_reduce_gather_sectors():
    """Why the index layout of the forward projector looks the way it does (kernels.py CTProjector, NumPy transcription
    of its formulas): leading padding (first row - slice's first row)*(1 + |tan|) for steep rays, and for shallow rays
    addresses in the TRANSPOSED image with the same alignment on the column where the ray enters its first row."""
    nx = ny = 192
    n_det = O.ct_num_detectors(nx)

    def layout(theta, transposed):
        c, s = np.cos(theta), np.sin(theta)
        A = O.ct_matrix(nx, np.array([theta]))
        rows = [A.indices[A.indptr[r]:A.indptr[r + 1]].astype(np.int64) for r in range(n_det)]
        key = []
        for cols in rows:
            if cols.size == 0:
                key.append(None)
                continue
            iy0 = cols[0] // nx
            run = cols[cols // nx == iy0] % nx
            rightwards = (s < 0) != (c < 0)
            key.append(iy0 if not transposed else (run.min() if rightwards else nx - 1 - run.max()))
        if transposed:
            rows = [(cols % nx) * ny + cols // nx for cols in rows]
        rate = 1 + (abs(s / c) if not transposed else abs(c / s))
        skips = []
        for s0 in range(0, n_det, 32):
            ks = [k for k in key[s0:s0 + 32] if k is not None]
            kmin = min(ks) if ks else 0
            skips += [0 if k is None else int(np.floor(rate * (k - kmin))) for k in key[s0:s0 + 32]]
        return rows, skips

    for deg in (20.0, 40.0):  # steep rays: alignment alone
        rows, skips = layout(np.radians(deg), transposed=False)
        plain, aligned = _sectors_per_entry(rows, [0] * n_det), _sectors_per_entry(rows, skips)
        assert aligned < 0.85 * plain, (deg, plain, aligned)
    for deg in (70.0, 110.0):  # shallow rays: in the image every lane sits in its own row; transposed they share sectors
        rows_img, _ = layout(np.radians(deg), transposed=False)
        rows_t, skips_t = layout(np.radians(deg), transposed=True)
        plain, fixed = _sectors_per_entry(rows_img, [0] * n_det), _sectors_per_entry(rows_t, skips_t)
        # (not as low as a steep ray's 0.35: the entries keep their (iy, ix) order, a sawtooth in the transposed image)
        assert plain > 0.75 and fixed < 0.65 * plain, (deg, plain, fixed)


def forward_reevaluated(nx, ny, n_det, theta, x, transpose_shallow=True):
    """NumPy transcription of spmv_sell_kernel<GEOM> + the index-only builder (one ray at a time, entries vectorised):
    the ray's pixels come from the stored pattern (the index stream), their VALUES are re-evaluated from the geometry;
    shallow rays carry indices into the transposed image and gather from it."""
    A = O.ct_matrix(nx, theta, ny=ny, n_det=n_det)
    xT = x.reshape(ny, nx).T.copy().reshape(-1)  # xT[ix*ny + iy] = x[iy*nx + ix]
    bias_x, bias_y = 2251799813685248.0 + 0.5 * (nx - 1), 2251799813685248.0 + 0.5 * (ny - 1)

    def insert(k, bias):  # centred_coord: integer into the mantissa of 2^51, one exact subtraction
        return (((np.int64(0x43200000) << 32) | (k.astype(np.int64) << 1)).view(np.float64)) - bias

    y = np.zeros(A.shape[0])
    for a, th in enumerate(theta):
        c, s = np.cos(th), np.sin(th)
        hi, lo = max(abs(c), abs(s)), min(abs(c), abs(s))
        d2 = 0.5 * (hi + lo)
        with np.errstate(divide="ignore"):
            inv_hi, inv_hilo = 1.0 / hi, 1.0 / (hi * lo)
        tmode = transpose_shallow and abs(s) > abs(c)
        for d in range(n_det):
            row = a * n_det + d
            cols = A.indices[A.indptr[row]:A.indptr[row + 1]].astype(np.int64)
            iy, ix = cols // nx, cols % nx
            stored = ix * ny + iy if tmode else cols  # what tb200_ct_fill_rows_aligned writes
            div = ny if tmode else nx
            mg, sh = _magic(div)
            q = ((stored.astype(np.uint64) * np.uint64(mg)) >> np.uint64(32)) >> np.uint64(sh)
            q = q.astype(np.int64)
            r = stored - q * div
            fix = r >= div
            q, r = q + fix, r - fix * div
            if tmode:   # (q, r) = (ix, iy): proj = r-term + q-term = cy*s + cx*c
                proj = insert(r, bias_y) * s + insert(q, bias_x) * c
                xs = xT[stored]
            else:       # (q, r) = (iy, ix): proj = cx*c + cy*s
                proj = insert(r, bias_x) * c + insert(q, bias_y) * s
                xs = x[stored]
            sd = d - 0.5 * (n_det - 1)
            slope = (d2 - np.abs(sd - proj)) * inv_hilo
            val = np.where(slope < inv_hi, slope, inv_hi)
            acc = 0.0
            for p in val * xs:  # the rounding chain, in index order
                acc = acc + p
            y[row] = acc
    return y, A


@pytest.mark.parametrize("nx,ny,views,n_det", [(16, 16, 10, None), (21, 13, 7, None), (12, 28, 9, 25)])
def test_reevaluated_forward_projection_equals_the_stored_product_bitwise(nx, ny, views, n_det):
    rng = np.random.default_rng(nx + 7 * views)
    n_det = O.ct_num_detectors(nx) if n_det is None else n_det
    theta = np.concatenate((O.ct_angles(views), [np.pi / 4, 3 * np.pi / 4, np.pi / 2 + 1e-9]))
    x = rng.standard_normal(nx * ny)
    for transpose_shallow in (False, True):
        y, A = forward_reevaluated(nx, ny, n_det, theta, x, transpose_shallow)
        assert np.array_equal(y, A @ x), transpose_shallow


# ---- ray-driven forward projector (csrc/ct_forward.cu) -----------------------------------------------------------------

FW_ETA = 1e-6


def forward_rays_spec(nx, ny, n_det, theta, x, run_tan=3.0):
    """NumPy transcription of ct_forward_rays_kernel, vectorised over the detectors of an angle: the lockstep form
    (candidate bracket per image row from the DDA estimate, LMAX = ceil(1 + |s/c| + slack) candidates, exact predicate,
    ascending order) and the run form (|s/c| > 3: per row the pixels lo..hi of an over-estimated run).  Returns
    (y, hits, candidates): the product, the number of entries that passed the predicate, the candidates evaluated."""
    x0, y0 = 0.5 * (nx - 1), 0.5 * (ny - 1)
    sd = np.arange(n_det) - 0.5 * (n_det - 1)
    X = x.reshape(ny, nx)
    y = np.zeros(len(theta) * n_det)
    hits = cands = 0
    for a, th in enumerate(theta):
        c, s = np.cos(th), np.sin(th)
        ac, as_ = abs(c), abs(s)
        hi, lo = max(ac, as_), min(ac, as_)
        d2 = 0.5 * (hi + lo)
        with np.errstate(divide="ignore"):
            inv_hi, inv_hilo = 1.0 / hi, 1.0 / (hi * lo)
        acc = np.zeros(n_det)

        def candidate(acc, ix, Q, ok):
            cx = np.clip(ix, 0, nx - 1) - x0
            t = sd - (cx * c + Q)
            e = d2 - np.abs(t)
            with np.errstate(invalid="ignore", over="ignore"):
                sl = e * inv_hilo
                w = np.where(sl < inv_hi, sl, inv_hi)
            hit = ok & (e > 0)
            return hit, w

        if as_ > run_tan * ac:  # run form
            flat = (ac + as_) >= 0.5 * nx * ac
            inv_c = 0.0 if flat else 1.0 / c
            h = d2 * abs(inv_c)
            slope = -s * inv_c
            E0 = (sd + y0 * s) * inv_c + x0 - h - FW_ETA - 0.5
            width = 2.0 * h + 2.0 * FW_ETA + 1e-7
            wlen = float(nx) if flat else np.ceil(width)
            reach = (ac * (x0 + 1.0) + d2) / as_ + 1e-6
            yc = sd / s
            ra = np.maximum(np.ceil(yc - reach + y0), 0).astype(np.int64)
            rb = np.minimum(np.floor(yc + reach + y0) + 1, ny).astype(np.int64)
            for iy in range(ny):
                rowok = (iy >= ra) & (iy < rb)
                Q = (iy - y0) * s
                if flat:
                    lo_i, hi_i = np.zeros(n_det, dtype=np.int64), np.full(n_det, nx - 1, dtype=np.int64)
                else:
                    e1 = np.clip(E0 + iy * slope, -1073741824.0, 1073741824.0)  # the kernel's DDA, restarted per lane at ra
                    i0 = ((e1 + MAGIC) - MAGIC).astype(np.int64) + 1
                    lo_i, hi_i = np.maximum(i0, 0), np.minimum(i0 + wlen - 1.0, nx - 1.0).astype(np.int64)
                hi_i = np.where(rowok, hi_i, -1)
                live_rows = hi_i >= lo_i
                if not live_rows.any():
                    continue
                for ix in range(int(lo_i[live_rows].min()), int(hi_i.max()) + 1):
                    ok = (ix >= lo_i) & (ix <= hi_i)
                    hit, w = candidate(acc, np.full(n_det, ix), Q, ok)
                    acc = np.where(hit, acc + w * X[iy, ix], acc)
                    hits += int(hit.sum())
                    cands += int(ok.sum())
        else:  # lockstep form
            width = (ac + as_) / ac + 2.0 * FW_ETA + 1e-7
            lmax = min(max(int(np.ceil(width)), 2), 9)
            inv_c = 1.0 / c
            h = d2 * abs(inv_c)
            slope = -s * inv_c
            E0 = (sd + y0 * s) * inv_c + x0 - h - FW_ETA - 0.5
            e1 = E0.copy()  # the kernel starts its DDA at the warp's first active row; the drift is what eta covers
            for iy in range(ny):
                Q = (iy - y0) * s
                i0 = ((e1 + MAGIC) - MAGIC).astype(np.int64) + 1
                e1 = e1 + slope
                for k in range(lmax):
                    ix = i0 + k
                    ok = (ix >= 0) & (ix < nx)
                    hit, w = candidate(acc, ix, Q, ok)
                    acc = np.where(hit, acc + w * X[iy, np.clip(ix, 0, nx - 1)], acc)
                    hits += int(hit.sum())
                    cands += int(ok.sum())
        y[a * n_det:(a + 1) * n_det] = acc
    return y, hits, cands


FORWARD_GEOMETRIES = [(16, 16, 12, None), (21, 13, 9, None), (24, 24, 8, 20), (10, 30, 16, 33), (32, 32, 5, 46),
                      (48, 48, 40, None), (33, 47, 36, 71), (40, 28, 24, 57)]


@pytest.mark.parametrize("run_tan", [3.0, 0.5, 7.9])
@pytest.mark.parametrize("nx,ny,views,n_det", FORWARD_GEOMETRIES)
def test_ray_driven_forward_projection_equals_the_stored_product_bitwise(nx, ny, views, n_det, run_tan):
    """Same pattern (number of entries that pass the predicate == nnz), same values, same order: A @ x bit for bit,
    including axis-aligned angles, 45 / 135 degrees, 30 / 60 degrees (rays through pixel corners), odd and even sizes,
    clipped and over-wide detectors."""
    n_det = O.ct_num_detectors(nx) if n_det is None else n_det
    theta = O.ct_angles(views)
    A = O.ct_matrix(nx, theta, ny=ny, n_det=n_det)
    x = np.random.default_rng(nx * 1000 + ny).standard_normal(nx * ny)
    y, hits, cands = forward_rays_spec(nx, ny, n_det, theta, x, run_tan)
    assert hits == A.nnz
    assert np.array_equal(y, A @ x)
    assert cands < 2.6 * A.nnz  # the brackets stay tight: < 2.6 candidates per stored entry on these small images


def test_ray_driven_forward_projection_awkward_angles():
    """Angles a hair away from the class boundaries (|s/c| = 1 and 3), from the axes, and rays exactly along pixel
    edges (odd detector count on an even image at 0 and 90 degrees)."""
    nx = ny = 20
    n_det = 29
    base = [0.0, np.pi / 2, np.pi / 4, 3 * np.pi / 4, np.arctan(3.0), np.pi - np.arctan(3.0), np.arctan(1.0 / 3.0)]
    theta = np.array(sorted(set(t + d for t in base for d in (0.0, 1e-7, -1e-7, 3e-6, -3e-6, 1e-12) if 0 <= t + d < np.pi)))
    A = O.ct_matrix(nx, theta, ny=ny, n_det=n_det)
    x = np.random.default_rng(7).standard_normal(nx * ny)
    y, hits, _ = forward_rays_spec(nx, ny, n_det, theta, x)
    assert hits == A.nnz
    assert np.array_equal(y, A @ x)


# ---- fan beam: the same walk with per-ray geometry (ct_forward_rays_fan_kernel) ------------------------------------------

def forward_rays_fan_spec(nx, ny, n_det, theta, x, fan, run_tan=7.9):
    """NumPy transcription of ct_forward_rays_fan_kernel: per-ray (c, s, rho) from the builder's ray tables, the form chosen
    per warp of 32 neighbouring rays (run form for the shallow rays of a warp in which some ray has |s/c| > run_tan, lockstep
    with the widest bracket of the warp otherwise)."""
    x0, y0 = 0.5 * (nx - 1), 0.5 * (ny - 1)
    X = x.reshape(ny, nx)
    y = np.zeros(len(theta) * n_det)
    hits = 0
    warp = np.arange(n_det) // 32
    nw = warp.max() + 1
    for a, th in enumerate(theta):
        c, s, rho, d2, inv_hi, inv_hilo = O._ray_tables(np.cos(th), np.sin(th), n_det, fan)
        ac, as_ = np.abs(c), np.abs(s)
        acc = np.zeros(n_det)

        def candidate(ix, Q, ok):
            cx = np.clip(ix, 0, nx - 1) - x0
            t = rho - (cx * c + Q)
            e = d2 - np.abs(t)
            with np.errstate(invalid="ignore", over="ignore"):
                sl = e * inv_hilo
                w = np.where(sl < inv_hi, sl, inv_hi)
            return ok & (e > 0), w

        need = np.zeros(nw, dtype=bool)
        np.logical_or.at(need, warp, as_ > run_tan * ac)
        in_runs = need[warp] & (as_ > ac)
        in_lock = ~in_runs
        with np.errstate(divide="ignore", invalid="ignore"):
            width = (ac + as_) / ac + 2.0 * FW_ETA + 1e-7
            inv_c = np.where(c != 0, 1.0 / c, 0.0)
        lm = np.where(in_lock, np.ceil(np.minimum(width, 64.0)), 2).astype(np.int64)
        lmax_w = np.full(nw, 2, dtype=np.int64)
        np.maximum.at(lmax_w, warp, lm)
        lmax = lmax_w[warp]
        h = d2 * np.abs(inv_c)
        slope = -s * inv_c
        E0 = (rho + y0 * s) * inv_c + x0 - h - FW_ETA - 0.5
        # lockstep lanes
        if in_lock.any():
            e1 = E0.copy()
            for iy in range(ny):
                Q = (iy - y0) * s
                with np.errstate(invalid="ignore"):
                    i0 = ((np.where(in_lock, e1, 0.0) + MAGIC) - MAGIC).astype(np.int64) + 1
                e1 = e1 + slope
                for k in range(int(lmax[in_lock].max())):
                    ix = i0 + k
                    ok = in_lock & (k < lmax) & (ix >= 0) & (ix < nx)
                    hit, w = candidate(ix, Q, ok)
                    acc = np.where(hit, acc + w * X[iy, np.clip(ix, 0, nx - 1)], acc)
                    hits += int(hit.sum())
        # run-form lanes
        if in_runs.any():
            flat = (ac + as_) >= 0.5 * nx * ac
            wlen = np.where(flat, float(nx), np.ceil(2.0 * h + 2.0 * FW_ETA + 1e-7))
            with np.errstate(divide="ignore", invalid="ignore"):
                reach = (ac * (x0 + 1.0) + d2) / as_ + 1e-6
                yc = rho / s
            ra = np.where(in_runs, np.maximum(np.ceil(yc - reach + y0), 0), 0).astype(np.int64)
            rb = np.where(in_runs, np.minimum(np.floor(yc + reach + y0) + 1, ny), 0).astype(np.int64)
            for iy in range(ny):
                rowok = in_runs & (iy >= ra) & (iy < rb)
                if not rowok.any():
                    continue
                Q = (iy - y0) * s
                e1 = np.clip(np.where(flat, 0.0, E0 + iy * slope), -1073741824.0, 1073741824.0)
                i0 = ((e1 + MAGIC) - MAGIC).astype(np.int64) + 1
                lo_i = np.where(flat, 0, np.maximum(i0, 0))
                hi_i = np.where(flat, nx - 1, np.minimum(i0 + wlen - 1.0, nx - 1.0)).astype(np.int64)
                hi_i = np.where(rowok, hi_i, -1)
                live_rows = hi_i >= lo_i
                if not live_rows.any():
                    continue
                for ix in range(int(lo_i[live_rows].min()), int(hi_i.max()) + 1):
                    ok = (ix >= lo_i) & (ix <= hi_i)
                    hit, w = candidate(np.full(n_det, ix), Q, ok)
                    acc = np.where(hit, acc + w * X[iy, ix], acc)
                    hits += int(hit.sum())
        y[a * n_det:(a + 1) * n_det] = acc
    return y, hits


@pytest.mark.parametrize("nx,ny,views,n_det,fan", [(16, 16, 12, None, None), (24, 24, 9, 40, (60.0, 20.0, 1.1)),
                                                   (21, 13, 8, 33, (40.0, 0.0, 1.0)), (32, 32, 10, 70, (70.0, 30.0, 1.3)),
                                                   (20, 36, 16, 45, None)])
def test_ray_driven_fan_beam_forward_projection_equals_the_stored_product_bitwise(nx, ny, views, n_det, fan):
    n_det = O.ct_num_detectors(nx) if n_det is None else n_det
    fan = O.fan_geometry(nx) if fan is None else fan
    theta = O.ct_angles(views)
    A = O.ct_matrix(nx, theta, ny=ny, n_det=n_det, fan=fan)
    x = np.random.default_rng(nx + 7 * ny).standard_normal(nx * ny)
    for run_tan in (7.9, 0.7):
        y, hits = forward_rays_fan_spec(nx, ny, n_det, theta, x, fan, run_tan)
        assert hits == A.nnz
        assert np.array_equal(y, A @ x)
