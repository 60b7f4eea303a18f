// Parallel-beam CT system matrix, built directly on the device in CSR form, together with its explicitly
// stored transpose (also CSR, i.e. one gather row per pixel).
//
// Role in the reference: the tomography operator comes from ASTRA (trips/test_problems/Tomography.py:49-88,
// explicit-matrix idiom trips/utilities/cil_io.py:271-275, per-frame parallel beam trips/utilities/io.py:392-400).
// ASTRA is not part of the reference tree, so this is a new synthetic problem generator with the same geometry
// conventions (theta = linspace(0, pi, views, endpoint=False), n_det = int(sqrt(2)*nx), Tomography.py:53-56):
// entry (ray, pixel) = length of the intersection of the ray with the unit pixel ("line" projector model).
//
// The entry definition (trapezoid chord length, every step separately rounded) lives in tb200_ctgeom.cuh.
// Both builders evaluate the SAME device function on the same (ray, pixel) arguments with explicitly rounded
// arithmetic, so A and A^T hold bit-identical values and an identical sparsity pattern: the stored A^T is the
// exact transpose of A, which the parity tests check against scipy's A.T.
//
// Row order of A: (local angle index)*n_det + detector; column = iy*nx + ix.  Both matrices have sorted indices.
//
// Two storage layouts for the same rows (argument `sell`):
//   CSR      entry j of row r at rowptr[r] + j
//   SELL-32-4 ("row-interleaved CSR", see spmv.cu): rows in slices of 32, entry j of row r at
//            sliceptr[r/32] + (j/4)*128 + (r%32)*4 + j%4 ; the caller zero-fills the arrays (padding = 0.0 * x[0]).
#include "tb200_common.cuh"
#include "tb200_ctgeom.cuh"

namespace tb200 {

// inclusive pixel range [lo, hi] of image row iy crossed by ray (g, sd); empty when hi < lo
__device__ __forceinline__ void row_range(const RayGeom& g, double sd, int iy, int nx, int ny, int& lo, int& hi) {
  const double cy = (double)iy - 0.5 * (double)(ny - 1);
  const double x0 = 0.5 * (double)(nx - 1);
  auto pred = [&](int ix) { return hits(g, ray_pixel_t(g, sd, (double)ix - x0, cy)); };
  if (fabs(g.c) < 1e-9) {
    // ray (numerically) parallel to the image rows: the footprint does not move along the row
    lo = 0;
    hi = nx - 1;
    if (pred(0) && pred(nx - 1)) return;
  } else {
    const double q = sd - cy * g.s;
    double e0 = (q - g.d2) / g.c, e1 = (q + g.d2) / g.c;
    if (e0 > e1) {
      const double tmp = e0;
      e0 = e1;
      e1 = tmp;
    }
    e0 = fmin(fmax(e0 + x0, -2.0), (double)nx + 1.0);
    e1 = fmin(fmax(e1 + x0, -2.0), (double)nx + 1.0);
    lo = (int)ceil(e0);
    hi = (int)floor(e1);
    if (lo < 0) lo = 0;
    if (hi > nx - 1) hi = nx - 1;
  }
  // settle the estimate on the exact predicate (its true set is an interval: t is monotone in ix)
  while (lo > 0 && pred(lo - 1)) --lo;
  while (hi < nx - 1 && pred(hi + 1)) ++hi;
  while (lo <= hi && !pred(lo)) ++lo;
  while (hi >= lo && !pred(hi)) --hi;
}

// inclusive detector range hit by pixel (cx, cy) at angle g
__device__ __forceinline__ void det_range(const RayGeom& g, double cx, double cy, int n_det, int& lo, int& hi) {
  const double dc = 0.5 * (double)(n_det - 1);
  auto pred = [&](int d) { return hits(g, ray_pixel_t(g, (double)d - dc, cx, cy)); };
  const double p = cx * g.c + cy * g.s + dc;
  double e0 = fmin(fmax(p - g.d2, -2.0), (double)n_det + 1.0);
  double e1 = fmin(fmax(p + g.d2, -2.0), (double)n_det + 1.0);
  lo = (int)ceil(e0);
  hi = (int)floor(e1);
  if (lo < 0) lo = 0;
  if (hi > n_det - 1) hi = n_det - 1;
  while (lo > 0 && pred(lo - 1)) --lo;
  while (hi < n_det - 1 && pred(hi + 1)) ++hi;
  while (lo <= hi && !pred(lo)) ++lo;
  while (hi >= lo && !pred(hi)) --hi;
}

// fan beam: inclusive detector range hit by pixel (cx, cy) at the angle (cosa, sina).  The estimate is the bin under
// the perspective projection of the pixel centre; the exact per-ray predicate settles it (the shadow is an interval).
__device__ __forceinline__ bool fan_hits(const Beam& bm, double cosa, double sina, int d, int n_det, double cx, double cy) {
  RayGeom g;
  double rho;
  ray_geometry(bm, cosa, sina, d, n_det, g, rho);
  return hits(g, ray_pixel_t(g, rho, cx, cy));
}
__device__ __forceinline__ void det_range_fan(const Beam& bm, double cosa, double sina, double cx, double cy, int n_det,
                                              int& lo, int& hi) {
  const double qx = cx - bm.so * sina, qy = cy + bm.so * cosa;       // pixel relative to the source
  const double depth = -qx * sina + qy * cosa, lateral = qx * cosa + qy * sina;
  double est = lateral * (bm.so + bm.dd) / (depth * bm.dps) + 0.5 * (double)(n_det - 1);
  est = fmin(fmax(est, -2.0), (double)n_det + 1.0);
  auto pred = [&](int d) { return fan_hits(bm, cosa, sina, d, n_det, cx, cy); };
  lo = (int)floor(est);
  hi = lo + 1;
  if (lo < 0) lo = 0;
  if (hi > n_det - 1) hi = n_det - 1;
  while (lo > 0 && pred(lo - 1)) --lo;
  while (hi < n_det - 1 && pred(hi + 1)) ++hi;
  while (lo <= hi && !pred(lo)) ++lo;
  while (hi >= lo && !pred(hi)) --hi;
}

// address of entry j of `row` in either layout (ptr = rowptr for CSR, slice pointers for SELL-32-4)
__device__ __forceinline__ int64_t entry_addr(const int64_t* __restrict__ ptr, int sell, int64_t row, int64_t j) {
  if (!sell) return ptr[row] + j;
  return ptr[row >> 5] + (j >> 2) * 128 + (row & 31) * 4 + (j & 3);
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

// ---- A: one warp per ray -------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256)
ct_rows_kernel(Beam bm, int nx, int ny, int n_det, int n_ang, const double* __restrict__ cosv,
               const double* __restrict__ sinv, int32_t* __restrict__ counts, const int64_t* __restrict__ rowptr, int sell,
               int32_t* __restrict__ col, double* __restrict__ val, int32_t* __restrict__ first_row,
               const int32_t* __restrict__ rowskip, int32_t* __restrict__ first_run, int tshallow) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= (int64_t)n_ang * n_det) return;
  const int a = (int)(ray / n_det), d = (int)(ray % n_det);
  RayGeom g;
  double sd;
  ray_geometry(bm, cosv[a], sinv[a], d, n_det, g, sd);
  const double x0 = 0.5 * (double)(nx - 1);
  int64_t base = (FILL && rowskip != nullptr) ? rowskip[ray] : 0;  // position of the next entry of this row
  int total = 0;
  int first = ny;  // first image row the ray crosses
  int first_lo = 0, first_hi = 0;  // ... and the run of pixels it crosses there
  // index-only layouts can address a shallow ray's pixels in the TRANSPOSED image (ix*ny + iy): see spmv.cu, CtRays::xT
  const bool tmode = tshallow && fabs(g.s) > fabs(g.c);
  for (int iy0 = 0; iy0 < ny; iy0 += 32) {
    const int iy = iy0 + lane;
    int lo = 0, hi = -1;
    if (iy < ny) row_range(g, sd, iy, nx, ny, lo, hi);
    const int cnt = (hi >= lo) ? (hi - lo + 1) : 0;
    if (cnt > 0 && iy < first) first = iy, first_lo = lo, first_hi = hi;
    if (FILL) {
      int chunk;
      const int off = warp_excl_scan(cnt, lane, chunk);
      const double cy = (double)iy - 0.5 * (double)(ny - 1);
      int64_t j = base + off;
      for (int ix = lo; ix <= hi; ++ix, ++j) {
        const int64_t pos = entry_addr(rowptr, sell, ray, j);
        col[pos] = tmode ? ix * ny + iy : iy * nx + ix;
        if (val != nullptr) val[pos] = chord(g, ray_pixel_t(g, sd, (double)ix - x0, cy));
      }
      base += chunk;
    } else {
      total += cnt;
    }
  }
  if (!FILL) {
    const int mine = first;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      total += __shfl_xor_sync(0xffffffffu, total, o);
      first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    }
    if (lane == 0) {
      counts[ray] = total;
      if (first_row != nullptr) first_row[ray] = first;
    }
    // the lane that saw the first row reports where the ray enters it, as a distance travelled along x: the ray moves
    // towards +x with increasing row index when s/c < 0, towards -x otherwise
    if (first_run != nullptr && first < ny && mine == first) {
      const bool rightwards = (g.s < 0.0) != (g.c < 0.0);
      first_run[ray] = rightwards ? first_lo : (nx - 1 - first_hi);
    }
  }
}

// ---- A^T: one warp per pixel ---------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256)
ct_cols_kernel(Beam bm, int nx, int ny, int n_det, int n_ang, const double* __restrict__ cosv,
               const double* __restrict__ sinv, int32_t* __restrict__ counts, const int64_t* __restrict__ rowptr, int sell,
               int32_t* __restrict__ col, double* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t pix = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pix >= (int64_t)nx * ny) return;
  const int iy = (int)(pix / nx), ix = (int)(pix % nx);
  const double cx = (double)ix - 0.5 * (double)(nx - 1), cy = (double)iy - 0.5 * (double)(ny - 1);
  const double dc = 0.5 * (double)(n_det - 1);
  int64_t base = 0;  // entries of this row written so far
  int total = 0;
  for (int a0 = 0; a0 < n_ang; a0 += 32) {
    const int a = a0 + lane;
    int lo = 0, hi = -1;
    RayGeom g = make_geom(1.0, 0.0);
    double ca = 1.0, sa = 0.0;
    if (a < n_ang) {
      ca = cosv[a], sa = sinv[a];
      if (bm.fan) {
        det_range_fan(bm, ca, sa, cx, cy, n_det, lo, hi);
      } else {
        g = make_geom(ca, sa);
        det_range(g, cx, cy, n_det, lo, hi);
      }
    }
    const int cnt = (hi >= lo) ? (hi - lo + 1) : 0;
    if (FILL) {
      int chunk;
      const int off = warp_excl_scan(cnt, lane, chunk);
      int64_t j = base + off;
      for (int d = lo; d <= hi; ++d, ++j) {
        const int64_t pos = entry_addr(rowptr, sell, pix, j);
        col[pos] = a * n_det + d;
        double offs = (double)d - dc;
        if (bm.fan) ray_geometry(bm, ca, sa, d, n_det, g, offs);  // every ray of a fan has its own normal
        val[pos] = chord(g, ray_pixel_t(g, offs, cx, cy));
      }
      base += chunk;
    } else {
      total += cnt;
    }
  }
  if (!FILL) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0) counts[pix] = total;
  }
}

}  // namespace tb200

using namespace tb200;

extern "C" {

static int ct_args_ok(int nx, int ny, int n_det, int n_ang, const void* c, const void* s) {
  TB200_REQUIRE(nx > 0 && ny > 0 && n_det > 0 && n_ang >= 0, "bad geometry");
  TB200_REQUIRE((int64_t)nx * ny < ((int64_t)1 << 31) && (int64_t)n_ang * n_det < ((int64_t)1 << 31), "index space exceeds int32");
  TB200_REQUIRE(n_ang == 0 || (c && s), "null angle tables");
  return 0;
}

static int fan_args_ok(double so, double dd, double dps, int nx, int ny) {
  TB200_REQUIRE(so > 0.0 && dd >= 0.0 && dps > 0.0, "fan beam: source distance and bin width must be positive");
  TB200_REQUIRE(so * so > 0.25 * ((double)nx * nx + (double)ny * ny), "fan beam: the source must lie outside the image");
  return 0;
}

static int count_rows(Beam bm, int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv, int32_t* counts,
                      void* stream, int32_t* first_row = nullptr, int32_t* first_run = nullptr) {
  int rc = ct_args_ok(nx, ny, n_det, n_ang, cosv, sinv);
  if (rc) return rc;
  const int64_t rays = (int64_t)n_ang * n_det;
  if (rays == 0) return 0;
  TB200_REQUIRE(counts, "null counts");
  ct_rows_kernel<false><<<(unsigned)((rays * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      bm, nx, ny, n_det, n_ang, cosv, sinv, counts, nullptr, 0, nullptr, nullptr, first_row, nullptr, first_run, 0);
  return check_launch("ct_count_rows");
}

static int fill_rows(Beam bm, int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                     const int64_t* rowptr, int sell, int32_t* colidx, double* vals, void* stream,
                     const int32_t* rowskip = nullptr, int tshallow = 0) {
  int rc = ct_args_ok(nx, ny, n_det, n_ang, cosv, sinv);
  if (rc) return rc;
  const int64_t rays = (int64_t)n_ang * n_det;
  if (rays == 0) return 0;
  TB200_REQUIRE(rowptr && colidx, "null output");
  ct_rows_kernel<true><<<(unsigned)((rays * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      bm, nx, ny, n_det, n_ang, cosv, sinv, nullptr, rowptr, sell, colidx, vals, nullptr, rowskip, nullptr, tshallow);
  return check_launch("ct_fill_rows");
}

static int count_cols(Beam bm, int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv, int32_t* counts,
                      void* stream) {
  int rc = ct_args_ok(nx, ny, n_det, n_ang, cosv, sinv);
  if (rc) return rc;
  TB200_REQUIRE(counts, "null counts");
  const int64_t npix = (int64_t)nx * ny;
  ct_cols_kernel<false><<<(unsigned)((npix * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      bm, nx, ny, n_det, n_ang, cosv, sinv, counts, nullptr, 0, nullptr, nullptr);
  return check_launch("ct_count_cols");
}

static int fill_cols(Beam bm, int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                     const int64_t* rowptr, int sell, int32_t* colidx, double* vals, void* stream) {
  int rc = ct_args_ok(nx, ny, n_det, n_ang, cosv, sinv);
  if (rc) return rc;
  TB200_REQUIRE(rowptr && colidx && vals, "null output");
  const int64_t npix = (int64_t)nx * ny;
  ct_cols_kernel<true><<<(unsigned)((npix * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      bm, nx, ny, n_det, n_ang, cosv, sinv, nullptr, rowptr, sell, colidx, vals);
  return check_launch("ct_fill_cols");
}

static const Beam PARALLEL = {0, 0.0, 0.0, 1.0};

// counts[ray] = number of pixels crossed by ray (ray = angle*n_det + det), for the n_ang angles in cos/sin.
int tb200_ct_count_rows(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv, int32_t* counts,
                        void* stream) {
  return count_rows(PARALLEL, nx, ny, n_det, n_ang, cosv, sinv, counts, stream);
}

// Fills colidx/vals of A.  sell = 0: CSR, ptr = rowptr (exclusive prefix sum of the counts).
// sell = 1: SELL-32-4, ptr = slice pointers; colidx/vals must have been zero-filled by the caller.
// vals may be NULL: only the column indices are written (index-only matrix for tb200_ct_forward_f64).
int tb200_ct_fill_rows(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                       const int64_t* rowptr, int sell, int32_t* colidx, double* vals, void* stream) {
  return fill_rows(PARALLEL, nx, ny, n_det, n_ang, cosv, sinv, rowptr, sell, colidx, vals, stream);
}

// The index-only, ROW-ALIGNED form of A for the matrix-free forward projector (tb200_ct_forward_f64).
// count pass: counts[ray] and first_row[ray] = first image row the ray crosses (ny for a ray that misses the image).
// fill pass: column indices only, SELL-32-4, entry j of ray r at position rowskip[r] + j of its lane: the caller
// chooses rowskip so that the 32 rays of a slice walk through the same image rows at the same positions (their
// x-gathers then share sectors); skipped positions stay zero-filled and are masked by the kernel.
// first_run[ray] (nullable): how far along x, measured from the side it comes from, the ray enters its first row - the
// alignment key of shallow rays.  transpose_shallow != 0: rays with |sin| > |cos| get the column index ix*ny + iy (their
// pixels addressed in the transposed image, which tb200_ct_forward_f64 forms in its xT scratch): neighbouring shallow
// rays then gather neighbouring addresses, exactly as steep rays do in the image itself.
int tb200_ct_count_rows_first(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                              int32_t* counts, int32_t* first_row, int32_t* first_run, void* stream) {
  TB200_REQUIRE(first_row != nullptr || (int64_t)n_ang * n_det == 0, "null first_row");
  return count_rows(PARALLEL, nx, ny, n_det, n_ang, cosv, sinv, counts, stream, first_row, first_run);
}
// the same fill with the VALUES written next to the indices (the stored row-aligned layout of tb200_ct_spmv_sell_f64)
int tb200_ct_fill_rows_aligned_vals(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                                    const int64_t* sliceptr, const int32_t* rowskip, int transpose_shallow, int32_t* colidx,
                                    double* vals, void* stream) {
  TB200_REQUIRE(vals != nullptr || (int64_t)n_ang * n_det == 0, "null vals");
  return fill_rows(PARALLEL, nx, ny, n_det, n_ang, cosv, sinv, sliceptr, 1, colidx, vals, stream, rowskip,
                   transpose_shallow);
}
int tb200_ct_fill_rows_aligned(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                               const int64_t* sliceptr, const int32_t* rowskip, int transpose_shallow, int32_t* colidx,
                               void* stream) {
  return fill_rows(PARALLEL, nx, ny, n_det, n_ang, cosv, sinv, sliceptr, 1, colidx, nullptr, stream, rowskip,
                   transpose_shallow);
}

// counts[pixel] = number of rays crossing the pixel (rows of A^T).
int tb200_ct_count_cols(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv, int32_t* counts,
                        void* stream) {
  return count_cols(PARALLEL, nx, ny, n_det, n_ang, cosv, sinv, counts, stream);
}

// Fills colidx/vals of A^T (rows = pixels; column = angle*n_det + det); layouts as for tb200_ct_fill_rows.
int tb200_ct_fill_cols(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                       const int64_t* rowptr, int sell, int32_t* colidx, double* vals, void* stream) {
  return fill_cols(PARALLEL, nx, ny, n_det, n_ang, cosv, sinv, rowptr, sell, colidx, vals, stream);
}

// Flat-detector fan beam (the geometry the reference asks ASTRA for: 'fanflat' + 'line_fanflat',
// trips/test_problems/Tomography.py:57-67): source at distance so from the rotation centre, detector at distance dd
// behind it, bins of width dps.  Same four passes, same layouts, same entry function (chord of the source->bin ray
// through the unit pixel).
int tb200_ctfan_count_rows(double so, double dd, double dps, int nx, int ny, int n_det, int n_ang, const double* cosv,
                           const double* sinv, int32_t* counts, void* stream) {
  int rc = fan_args_ok(so, dd, dps, nx, ny);
  if (rc) return rc;
  const Beam bm = {1, so, dd, dps};
  return count_rows(bm, nx, ny, n_det, n_ang, cosv, sinv, counts, stream);
}
int tb200_ctfan_fill_rows(double so, double dd, double dps, int nx, int ny, int n_det, int n_ang, const double* cosv,
                          const double* sinv, const int64_t* rowptr, int sell, int32_t* colidx, double* vals, void* stream) {
  int rc = fan_args_ok(so, dd, dps, nx, ny);
  if (rc) return rc;
  const Beam bm = {1, so, dd, dps};
  return fill_rows(bm, nx, ny, n_det, n_ang, cosv, sinv, rowptr, sell, colidx, vals, stream);
}
int tb200_ctfan_count_cols(double so, double dd, double dps, int nx, int ny, int n_det, int n_ang, const double* cosv,
                           const double* sinv, int32_t* counts, void* stream) {
  int rc = fan_args_ok(so, dd, dps, nx, ny);
  if (rc) return rc;
  const Beam bm = {1, so, dd, dps};
  return count_cols(bm, nx, ny, n_det, n_ang, cosv, sinv, counts, stream);
}
int tb200_ctfan_fill_cols(double so, double dd, double dps, int nx, int ny, int n_det, int n_ang, const double* cosv,
                          const double* sinv, const int64_t* rowptr, int sell, int32_t* colidx, double* vals, void* stream) {
  int rc = fan_args_ok(so, dd, dps, nx, ny);
  if (rc) return rc;
  const Beam bm = {1, so, dd, dps};
  return fill_cols(bm, nx, ny, n_det, n_ang, cosv, sinv, rowptr, sell, colidx, vals, stream);
}

}  // extern "C"
