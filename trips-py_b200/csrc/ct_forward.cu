// Matrix-free (ray-driven) parallel-beam forward projection  y = A x - coef*z  for the CT matrix of ct_builder.cu,
// bit-identical to the sequential-order SpMV on the stored matrix (and so to scipy's A @ x on the same matrix).
//
// Role in the reference: A @ v inside golub_kahan_update / CGLS / GKS / MMGKS (trips/utilities/decompositions.py:240,
// trips/solvers/CGLS.py:60, GKS.py:92, MMGKS.py:124); the reference's own tomography operator is matrix-free
// (astra.OpTomo behind a pylops FunctionOperator, trips/test_problems/Tomography.py:73-83).  SURVEY.md 8(f) item 1.
//
// Round 1 re-evaluated the VALUES but still streamed A's column indices (4 B per entry, 15.7 GB at 2048^2 x 720) only to
// learn which pixels a ray meets, in index order.  This kernel enumerates them instead.  Nothing is read but x.
//
// The row of ray (angle a, detector d) must be summed over its pixels in ascending column index iy*nx + ix, every
// product and every addition separately rounded (scipy's csr_matvec).  In image row iy the pixels the ray meets are
// those with |t| < d2, t = sd - (cx*c + cy*s): an interval of ix around xi = ((sd - cy*s)/c) + x0 of half-width
// h = d2/|c| = (1 + |s/c|)/2.  So: walk the rows upwards, and in each row test LMAX = ceil(2h + slack) consecutive
// candidates starting at the first integer above xi - h - eta, with the SAME predicate on the SAME separately rounded
// t as the builder: a superset of candidates + the exact predicate = the stored pattern, in the stored order.
//   * lane = ray; the 32 lanes of a warp are neighbouring detectors of one angle and walk the rows in lockstep, so at
//     every step their gathers fall into 32/|c| consecutive pixels of ONE image row (3-4 cache lines for steep rays).
//   * the bracket position is a DDA: e += slope per row, converted with the 1.5*2^52 trick; its rounding drift
//     (< 1e-9 pixels over 16k rows) is covered by the slack eta = 1e-6.
//   * rows where every lane's candidates are inside the image run without bounds tests.
//   * per-row term cy*s: a shared-memory table per CTA (one angle per CTA), one broadcast LDS per row.
// Rays with |s/c| > 3 (within 18.4 degrees of the image rows: few rows, long runs) take the run form: lane = ray, each
// lane walks the runs of its own rows with 256-bit loads of four consecutive pixels.
// fp64 instructions per candidate: 9 (coordinate, cx*c, +cy*s, sd-, margin, slope, compare, product, add) + 2 per row.
#include "tb200_common.cuh"
#include "tb200_ctgeom.cuh"
#include "tb200_dd.cuh"

namespace tb200 {

constexpr double FW_ETA = 1e-6;    // slack of the candidate bracket, in pixels
constexpr int FW_WARPS = 4;        // 128 rays of one angle per CTA (64 when the launch has few CTAs: angle shards)
constexpr double FW_RUN_TAN_MAX = 7.9;  // lockstep form is instantiated for up to 9 candidates per row
// measured at 2048^2 x 720 (tools/fw_tune.py, profiles/): run_tan 3 / 5 / 7.9 -> 4.87 / 4.50 / 4.37 ms with 64 registers
// (8 CTAs per SM), 5.47 / 4.89 / 4.70 ms with 90 registers (5 CTAs per SM)
static double g_fw_run_tan = 7.9;       // |s/c| above this: run form (tuning knob, tb200_ct_forward_set_tuning)
static int g_fw_minb = 8;               // resident CTAs (of 128 threads) per SM the compiler must allow for (register cap 128 / 64)
static int g_fw_warps_override = 0;     // 0 = choose 4 or 2 warps per CTA from the launch size

__device__ __forceinline__ double fw_add_if_positive(double acc, double p, int flag) {
  asm("{\n\t.reg .pred q;\n\tsetp.gt.s32 q, %2, 0;\n\t@q add.rn.f64 %0, %0, %1;\n\t}" : "+d"(acc) : "d"(p), "r"(flag));
  return acc;
}

struct FwRay {
  double sd, c, d2, inv_hi, inv_hilo, biasx;
};

// one candidate pixel (ix in [0, nx) guaranteed by the caller): acc += chord * x when the ray meets the pixel
__device__ __forceinline__ double fw_candidate_cx(double acc, const FwRay& r, double Q, double cx, double xv) {
  const double t = __dsub_rn(r.sd, __dadd_rn(__dmul_rn(cx, r.c), Q));
  const double e = __dsub_rn(r.d2, fabs(t));  // > 0 inside the footprint (never denormal: |t|, d2 = O(1))
  const double w = chord_from_margin(e, r.inv_hi, r.inv_hilo);
  return fw_add_if_positive(acc, __dmul_rn(w, xv), __double2hiint(e));
}
__device__ __forceinline__ double fw_candidate(double acc, const FwRay& r, double Q, int ix, double xv) {
  const double cx = centred_coord(ix, r.biasx);
  const double t = __dsub_rn(r.sd, __dadd_rn(__dmul_rn(cx, r.c), Q));
  const double e = __dsub_rn(r.d2, fabs(t));  // > 0 inside the footprint (never denormal: |t|, d2 = O(1))
  const double w = chord_from_margin(e, r.inv_hi, r.inv_hilo);
  return fw_add_if_positive(acc, __dmul_rn(w, xv), __double2hiint(e));
}

// rows [r0, r1) of the lockstep walk; CHECKED: candidates may fall outside [0, nx)
template <int LMAX, bool CHECKED, bool QTAB>
__device__ __forceinline__ void fw_rows(double& acc, double& e1, int r0, int r1, const FwRay& r, double slope, double s,
                                        double biasy, const double* __restrict__ qtab, const double* __restrict__ x, int nx,
                                        uint64_t pol) {
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  const double* xr = x + (int64_t)r0 * nx;
#pragma unroll(LMAX <= 2 ? 4 : 2)
  for (int iy = r0; iy < r1; ++iy, xr += nx) {
    const double Q = QTAB ? qtab[iy] : __dmul_rn(centred_coord(iy, biasy), s);
    const int i0 = __double2loint(__dadd_rn(e1, MAGIC)) + 1;  // first integer above the bracket's left end (or one below)
    e1 = __dadd_rn(e1, slope);
    if (CHECKED) {
#pragma unroll
      for (int k = 0; k < LMAX; ++k) {
        const int ix = i0 + k;
        if ((unsigned)ix < (unsigned)nx) acc = fw_candidate(acc, r, Q, ix, ld_gather_f64(xr + ix, pol));
      }
    } else {
      double xv[LMAX];
#pragma unroll
      for (int k = 0; k < LMAX; ++k) xv[k] = ld_gather_f64(xr + (i0 + k), pol);
      // coordinates of consecutive pixels differ by exactly 1 (half-integers far below 2^52: the additions are exact),
      // so cx of candidate k is the same double as centred_coord(i0 + k)
      double cx = centred_coord(i0, r.biasx);
#pragma unroll
      for (int k = 0; k < LMAX; ++k) {
        acc = fw_candidate_cx(acc, r, Q, cx, xv[k]);
        if (k + 1 < LMAX) cx = __dadd_rn(cx, 1.0);
      }
    }
  }
}

// rows of [0, ny) on which lo <= E0 + iy*slope <= hi  ->  [a, b) (empty: a >= b)
__device__ __forceinline__ void fw_row_window(double E0, double slope, double lo, double hi, int ny, int& a, int& b) {
  if (slope == 0.0) {
    const bool in = (E0 >= lo) && (E0 <= hi);
    a = 0;
    b = in ? ny : 0;
    return;
  }
  double ta = (lo - E0) / slope, tb = (hi - E0) / slope;
  if (ta > tb) {
    const double tmp = ta;
    ta = tb;
    tb = tmp;
  }
  ta = fmin(fmax(ta, -1.0), (double)ny + 1.0);
  tb = fmin(fmax(tb, -1.0), (double)ny + 1.0);
  a = max((int)ceil(ta), 0);
  b = min((int)floor(tb) + 1, ny);
}

template <int LMAX, bool QTAB>
__device__ __forceinline__ double fw_lockstep(const FwRay& r, double s, bool live, int nx, int ny, double biasy,
                                              const double* __restrict__ qtab, const double* __restrict__ x, uint64_t pol) {
  // bracket of row iy: candidates i0 .. i0 + LMAX - 1, i0 = rn(e1) + 1, e1 = xi - h - eta - 1/2 (see the header)
  const double inv_c = 1.0 / r.c;
  const double h = r.d2 * fabs(inv_c);
  const double x0 = 0.5 * (double)(nx - 1), y0 = 0.5 * (double)(ny - 1);
  const double slope = -s * inv_c;
  const double E0 = (r.sd + y0 * s) * inv_c + x0 - h - FW_ETA - 0.5;
  // rows on which some candidate can be inside the image, and rows on which all of them surely are
  int a0 = 0, a1 = 0, i0r = 0, i1r = 0;
  if (live) {
    fw_row_window(E0, slope, -(double)(LMAX + 1), (double)nx, ny, a0, a1);
    fw_row_window(E0, slope, 0.0, (double)(nx - LMAX) - 1.5, ny, i0r, i1r);
  }
  const bool has = live && a0 < a1;
  const unsigned FULL = 0xffffffffu;
  int wa0 = has ? a0 : ny, wa1 = has ? a1 : 0;
  int wi0 = has ? i0r : 0, wi1 = has ? i1r : ny;  // lanes without rows do not constrain the interior (they are masked below)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wa0 = min(wa0, __shfl_xor_sync(FULL, wa0, o));
    wa1 = max(wa1, __shfl_xor_sync(FULL, wa1, o));
    wi0 = max(wi0, __shfl_xor_sync(FULL, wi0, o));
    wi1 = min(wi1, __shfl_xor_sync(FULL, wi1, o));
  }
  if (wa0 >= wa1) return 0.0;
  if (!__all_sync(FULL, has)) wi0 = wi1 = wa0;  // a lane without rows would gather out of bounds in the unchecked loop
  wi0 = min(max(wi0, wa0), wa1);
  wi1 = min(max(wi1, wi0), wa1);
  double acc = 0.0;
  double e1 = __fma_rn((double)wa0, slope, E0);
  if (has) {
    fw_rows<LMAX, true, QTAB>(acc, e1, wa0, wi0, r, slope, s, biasy, qtab, x, nx, pol);
    fw_rows<LMAX, false, QTAB>(acc, e1, wi0, wi1, r, slope, s, biasy, qtab, x, nx, pol);
    fw_rows<LMAX, true, QTAB>(acc, e1, wi1, wa1, r, slope, s, biasy, qtab, x, nx, pol);
  }
  return acc;
}

// Run form: few rows per ray, long runs of consecutive pixels in each.  Every lane walks its OWN rows (row t of the
// walk is image row ra + t of that lane); the run of a row is bracketed by the same DDA as above, LRUN = ceil(2h + slack)
// pixels from the first integer above the bracket's left end, and read as aligned groups of four pixels (one 256-bit
// load each).  Trip counts are uniform over the warp (rows: max over lanes; groups per row: a constant of the angle).
__device__ __forceinline__ double fw_runs(const FwRay& r, double s, bool live, int nx, int ny, double biasy,
                                          const double* __restrict__ x, int64_t n, bool vec4, uint64_t pol) {
  const unsigned FULL = 0xffffffffu;
  const double MAGIC = 6755399441055744.0;
  const double x0 = 0.5 * (double)(nx - 1), y0 = 0.5 * (double)(ny - 1);
  const double ac = fabs(r.c), as = fabs(s);
  // runs of at least half a row (incl. c == 0): every pixel of the rows in reach is a candidate - always a superset, and
  // it keeps the DDA's magnitudes (~ |s/c| * n_det) far below the 2^30 clamp and its rounding far below eta
  const bool flat = (ac + as) >= 0.5 * (double)nx * ac;
  const double inv_c = flat ? 0.0 : 1.0 / r.c;
  const double h = r.d2 * fabs(inv_c);
  const double slope = -s * inv_c;
  const double E0 = (r.sd + y0 * s) * inv_c + x0 - h - FW_ETA - 0.5;
  const double width = 2.0 * h + 2.0 * FW_ETA + 1e-7;
  const double wlen = flat ? (double)nx : ceil(width);        // pixels per run (before clipping to the image: may be huge)
  const int lrun = (int)fmin(wlen, (double)nx);               // ... of which at most nx are inside
  const int ngroups = (lrun + 2) / 4 + 1;  // aligned groups of four that a run of lrun pixels can touch
  // rows the ray can meet: |cy - (sd - cx*c)/s| < d2/|s| for some |cx| <= x0 + 1/2
  int ra = 0, rb = 0;
  if (live) {
    const double reach = (ac * (x0 + 1.0) + r.d2) / as + 1e-6;
    const double yc = r.sd / s;
    ra = max((int)ceil(yc - reach + y0), 0);
    rb = min((int)floor(yc + reach + y0) + 1, ny);
  }
  int nrows = max(rb - ra, 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nrows = max(nrows, __shfl_xor_sync(FULL, nrows, o));
  double acc = 0.0;
  double e1 = __fma_rn((double)ra, slope, E0);
  for (int t = 0; t < nrows; ++t) {
    const int iy = ra + t;
    const double Q = __dmul_rn(centred_coord(iy, biasy), s);
    int lo = 0, hi = nx - 1;
    if (!flat) {
      const double ec = fmin(fmax(e1, -1073741824.0), 1073741824.0);
      const int i0 = __double2loint(__dadd_rn(ec, MAGIC)) + 1;
      e1 = __dadd_rn(e1, slope);
      lo = max(i0, 0);
      hi = (int)fmin((double)i0 + wlen - 1.0, (double)(nx - 1));  // may be < lo: the run lies outside the image
    }
    if (iy >= rb) hi = -1, lo = 0;
    const int64_t rowbase = (int64_t)iy * nx;
    const int64_t g0 = (rowbase + lo) & ~(int64_t)3;
    const int64_t last = rowbase + hi;
    for (int q4 = 0; q4 < ngroups; ++q4) {
      const int64_t g = g0 + 4 * (int64_t)q4;
      if (hi >= lo && g <= last) {
        double xv[4];
        if (vec4 && g + 3 < n) {
          asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                       : "=d"(xv[0]), "=d"(xv[1]), "=d"(xv[2]), "=d"(xv[3])
                       : "l"(x + g), "l"(pol));
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) xv[k] = (g + k < n) ? ld_gather_f64(x + g + k, pol) : 0.0;
        }
        const int ixb = (int)(g - rowbase);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int ix = ixb + k;
          if (ix >= lo && ix <= hi) acc = fw_candidate(acc, r, Q, ix, xv[k]);
        }
      }
    }
  }
  return acc;
}

template <bool QTAB, int MINB, int W>
__global__ void __launch_bounds__(W * 32, MINB * 4 / W)
ct_forward_rays_kernel(int nx, int ny, int n_det, int n_ang, int nblk, const double* __restrict__ geom,
                       const double* __restrict__ x, double* __restrict__ y, double coef_host,
                       const double* __restrict__ coef_dev, const double* __restrict__ z, double* __restrict__ partials,
                       int vec4, double run_tan, PeerOut po, const int32_t* __restrict__ cta_order) {
  extern __shared__ __align__(16) double qtab[];  // QTAB: cy*s for every image row of this CTA's angle
  const int lane = threadIdx.x & 31;
  // CTA -> (angle, block of W*32 detectors).  With cta_order: entry b = angle * nblk + block of the b-th heaviest CTA
  // (longest-processing-time-first: the SMs drain evenly).  Without: blocks from the detector centre outwards, the long
  // central rays of every angle first, the short peripheral ones in the tail.
  int a, blk;
  bool blk_ok = true;
  if (cta_order != nullptr) {
    const int p = cta_order[blockIdx.x];
    a = p / nblk, blk = p - a * nblk;
  } else {
    a = blockIdx.x % n_ang;
    const int rank = blockIdx.x / n_ang;
    const int mid = nblk >> 1;
    blk = (rank & 1) ? mid + ((rank + 1) >> 1) : mid - (rank >> 1);
    blk_ok = blk >= 0 && blk < nblk;  // (nblk even: one rank of the sequence falls outside)
  }
  const int d = blk * (W * 32) + threadIdx.x;
  const bool live = blk_ok && d < n_det;
  const double* gp = geom + 6 * (int64_t)a;
  const double c = gp[0], s = gp[1];
  FwRay r;
  r.c = c, r.d2 = gp[2], r.inv_hi = gp[3], r.inv_hilo = gp[4];
  r.sd = (double)d - 0.5 * (double)(n_det - 1);
  r.biasx = centred_bias(nx);
  const double biasy = centred_bias(ny);
  const uint64_t pol = policy_evict_last();
  const double ac = fabs(c), as = fabs(s);
  const bool runs = as > run_tan * ac;  // also c == 0
  if (QTAB && !runs) {
    for (int i = threadIdx.x; i < ny; i += W * 32) qtab[i] = __dmul_rn(centred_coord(i, biasy), s);
    __syncthreads();
  }
  double acc;
  if (runs) {
    acc = fw_runs(r, s, live, nx, ny, biasy, x, (int64_t)nx * ny, vec4 != 0, pol);
  } else {
    // candidates per row: all integers of an open interval of length 2h + 2 eta (+ DDA drift): ceil of it
    const double width = (ac + as) / ac + 2.0 * FW_ETA + 1e-7;
    if (width <= 2.0) acc = fw_lockstep<2, QTAB>(r, s, live, nx, ny, biasy, qtab, x, pol);
    else if (width <= 3.0) acc = fw_lockstep<3, QTAB>(r, s, live, nx, ny, biasy, qtab, x, pol);
    else if (width <= 4.0) acc = fw_lockstep<4, QTAB>(r, s, live, nx, ny, biasy, qtab, x, pol);
    else if (width <= 5.0) acc = fw_lockstep<5, QTAB>(r, s, live, nx, ny, biasy, qtab, x, pol);
    else if (width <= 6.0) acc = fw_lockstep<6, QTAB>(r, s, live, nx, ny, biasy, qtab, x, pol);
    else if (width <= 7.0) acc = fw_lockstep<7, QTAB>(r, s, live, nx, ny, biasy, qtab, x, pol);
    else if (width <= 8.0) acc = fw_lockstep<8, QTAB>(r, s, live, nx, ny, biasy, qtab, x, pol);
    else acc = fw_lockstep<9, QTAB>(r, s, live, nx, ny, biasy, qtab, x, pol);
  }
  const double coef = (z != nullptr) ? (coef_dev ? *coef_dev : coef_host) : 0.0;
  dd_t nrm = dd_zero();
  if (live) {
    const int64_t row = (int64_t)a * n_det + d;
    if (z != nullptr) acc = __dsub_rn(acc, __dmul_rn(coef, z[row]));
    y[row] = acc;
#pragma unroll
    for (int p = 0; p < 15; ++p)  // the other GPUs' copies of the vector, over NVLink (unrolled: parameters stay in the constant bank)
      if (p < po.n) po.p[p][row] = acc;
    nrm = dd_fma(nrm, acc, acc);
  }
  if (po.n > 0) __threadfence_system();
  if (partials != nullptr) {
    // one double-double partial per WARP (no CTA-wide barrier: the warps of a CTA finish at very different times)
    const dd_t tot2 = dd_warp_sum(nrm);
    if (lane == 0) {
      const int64_t w = (int64_t)blockIdx.x * W + (threadIdx.x >> 5);
      partials[2 * w] = tot2.hi;
      partials[2 * w + 1] = tot2.lo;
    }
  }
}

// Fan beam (flat detector, tb200_ctgeom.cuh Beam): the same walk with PER-RAY geometry - every ray of a fan has its own
// normal (c, s) and offset rho, computed once per lane by the builder's ray_geometry(), so cy*s is formed per lane and the
// form (lockstep with LMAX candidates / run form) is chosen per WARP from the widest bracket among its 32 rays (a wider
// bracket is still a superset).  Replaces the forward product of astra's 'line_fanflat' projector
// (trips/test_problems/Tomography.py:57-67, 73-83).
__global__ void __launch_bounds__(FW_WARPS * 32, 5)
ct_forward_rays_fan_kernel(Beam bm, int nx, int ny, int n_det, int n_ang, int nblk, const double* __restrict__ cosv,
                           const double* __restrict__ sinv, const double* __restrict__ x, double* __restrict__ y,
                           double coef_host, const double* __restrict__ coef_dev, const double* __restrict__ z,
                           double* __restrict__ partials, int vec4, double run_tan) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int a = blockIdx.x % n_ang;
  const int rank = blockIdx.x / n_ang;
  const int mid = nblk >> 1;
  const int blk = (rank & 1) ? mid + ((rank + 1) >> 1) : mid - (rank >> 1);
  const bool blk_ok = blk >= 0 && blk < nblk;
  const int d = blk * (FW_WARPS * 32) + threadIdx.x;
  const bool live = blk_ok && d < n_det;
  RayGeom g;
  double rho;
  ray_geometry(bm, cosv[a], sinv[a], live ? d : 0, n_det, g, rho);
  FwRay r;
  r.c = g.c, r.d2 = g.d2, r.inv_hi = g.inv_hi, r.inv_hilo = g.inv_hilo, r.sd = rho;
  r.biasx = centred_bias(nx);
  const double s = g.s;
  const double biasy = centred_bias(ny);
  const uint64_t pol = policy_evict_last();
  const double ac = fabs(r.c), as = fabs(s);
  // If some ray of the warp needs the run form, all its shallow rays (|s| > |c|) take it; the others - and every ray of a
  // warp without such a ray - walk in lockstep with the widest bracket among them.  (The 32 rays of a warp are nearly
  // parallel, so in practice a warp is all of one kind; the split only keeps degenerate fans correct.)
  const bool need_runs = __any_sync(FULL, live && as > run_tan * ac);
  const bool in_runs = need_runs && live && as > ac;
  const bool in_lock = live && !in_runs;
  const double width = (ac + as) / ac + 2.0 * FW_ETA + 1e-7;
  int lmax = in_lock ? (int)ceil(fmin(width, 64.0)) : 2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = max(lmax, __shfl_xor_sync(FULL, lmax, o));
  double acc_r = 0.0, acc_l = 0.0;
  if (need_runs) acc_r = fw_runs(r, s, in_runs, nx, ny, biasy, x, (int64_t)nx * ny, vec4 != 0, pol);
  if (__any_sync(FULL, in_lock)) {
    if (lmax <= 2) acc_l = fw_lockstep<2, false>(r, s, in_lock, nx, ny, biasy, nullptr, x, pol);
    else if (lmax == 3) acc_l = fw_lockstep<3, false>(r, s, in_lock, nx, ny, biasy, nullptr, x, pol);
    else if (lmax == 4) acc_l = fw_lockstep<4, false>(r, s, in_lock, nx, ny, biasy, nullptr, x, pol);
    else if (lmax == 5) acc_l = fw_lockstep<5, false>(r, s, in_lock, nx, ny, biasy, nullptr, x, pol);
    else if (lmax == 6) acc_l = fw_lockstep<6, false>(r, s, in_lock, nx, ny, biasy, nullptr, x, pol);
    else if (lmax == 7) acc_l = fw_lockstep<7, false>(r, s, in_lock, nx, ny, biasy, nullptr, x, pol);
    else if (lmax == 8) acc_l = fw_lockstep<8, false>(r, s, in_lock, nx, ny, biasy, nullptr, x, pol);
    else acc_l = fw_lockstep<9, false>(r, s, in_lock, nx, ny, biasy, nullptr, x, pol);  // (lmax <= 9: |s/c| <= run_tan <= 7.9)
  }
  double acc = in_runs ? acc_r : acc_l;
  const double coef = (z != nullptr) ? (coef_dev ? *coef_dev : coef_host) : 0.0;
  dd_t nrm = dd_zero();
  if (live) {
    const int64_t row = (int64_t)a * n_det + d;
    if (z != nullptr) acc = __dsub_rn(acc, __dmul_rn(coef, z[row]));
    y[row] = acc;
    nrm = dd_fma(nrm, acc, acc);
  }
  if (partials != nullptr) {
    const dd_t tot2 = dd_warp_sum(nrm);
    if (lane == 0) {
      const int64_t w = (int64_t)blockIdx.x * FW_WARPS + (threadIdx.x >> 5);
      partials[2 * w] = tot2.hi;
      partials[2 * w + 1] = tot2.lo;
    }
  }
}

// warps per CTA the launcher uses for a problem of this size
static int fw_warps_for(int n_det, int n_ang) {
  int warps = FW_WARPS;
  const int nb4 = (n_det + FW_WARPS * 32 - 1) / (FW_WARPS * 32);
  if ((int64_t)nb4 * n_ang < (int64_t)6 * sm_count() * g_fw_minb) warps = 2;
  if (g_fw_warps_override) warps = g_fw_warps_override;
  return warps;
}

}  // namespace tb200

using namespace tb200;

extern "C" {

// Doubles of workspace tb200_ct_forward_rays_f64 needs for its fused norm (one double-double partial per CTA).
int64_t tb200_ct_forward_rays_workspace_len(int n_det, int n_ang) {
  const int64_t nblk2 = (n_det + 63) / 64;  // 2-warp CTAs: the larger count of warps
  return 2 * ((nblk2 + 1) * (int64_t)n_ang * 2 + 8);
}

// Tuning knobs of the ray-driven forward projector (results never depend on them): run_tan = |s/c| above which an angle
// takes the run form (0 < run_tan <= 7.9); min_ctas = 4 or 8 resident CTAs per SM to compile for (<= 128 / 64 registers).
int tb200_ct_forward_set_tuning(double run_tan, int min_ctas) {
  TB200_REQUIRE(run_tan > 0.0 && run_tan <= FW_RUN_TAN_MAX, "run_tan out of range");
  TB200_REQUIRE(min_ctas == 4 || min_ctas == 8 || min_ctas == 42 || min_ctas == 44 || min_ctas == 82 || min_ctas == 84,
                "min_ctas must be 4 or 8 (optionally followed by the digit 2 or 4: warps per CTA forced)");
  g_fw_run_tan = run_tan;
  g_fw_warps_override = min_ctas > 9 ? min_ctas % 10 : 0;
  g_fw_minb = min_ctas > 9 ? min_ctas / 10 : min_ctas;
  return 0;
}

// y = A x - coef*z (z nullable; coef from coef_dev if non-null), optional norm_out = (||y||^2, ||y||), for the
// parallel-beam matrix of n_ang angles (geom from tb200_ct_geometry): x row-major (iy*nx + ix), y angle-major
// (angle*n_det + detector).  No matrix and no index array is read.  Same bits as tb200_spmv_sell_f64 on
// tb200_ct_fill_rows' matrix.  ws: tb200_ct_forward_rays_workspace_len(n_det, n_ang) doubles iff norm_out != NULL.
static int forward_rays_launch(int nx, int ny, int n_det, int n_ang, const double* geom, const double* x, double* y,
                               double coef_host, const double* coef_dev, const double* z, double* part, const PeerOut& po,
                               const int32_t* cta_order, cudaStream_t st, int64_t& nparts) {
  TB200_REQUIRE(nx > 0 && ny > 0 && n_det > 0 && n_ang >= 0, "bad geometry");
  TB200_REQUIRE((int64_t)nx * ny < ((int64_t)1 << 31) && (int64_t)n_ang * n_det < ((int64_t)1 << 31), "index space exceeds int32");
  nparts = 0;
  if (n_ang == 0) return 0;
  TB200_REQUIRE(geom && x && y, "null pointer");
  // rays per CTA: 128, or 64 when 128 would leave fewer than ~6 waves of CTAs (an angle shard of a multi-GPU run: the
  // long central CTAs then finish together and the SMs idle behind them; measured 0.83 -> see profiles/)
  const int warps = fw_warps_for(n_det, n_ang);
  const int nblk = (n_det + warps * 32 - 1) / (warps * 32);
  const int ranks = nblk + ((nblk & 1) ? 0 : 1);  // centre-out sequence mid, mid+1, mid-1, ...: covers [0, nblk) in `ranks` steps
  const int64_t nctas = (int64_t)(cta_order ? nblk : ranks) * n_ang;
  TB200_REQUIRE(nctas < ((int64_t)1 << 31), "too many CTAs");
  const int vec4 = ((uintptr_t)x % 32) == 0;
  const size_t qbytes = (size_t)ny * sizeof(double);
  const double run_tan = g_fw_run_tan;
#define FW_LAUNCH(QT, MB, WW, SMEM)                                                                                   \
  ct_forward_rays_kernel<QT, MB, WW><<<(unsigned)nctas, WW * 32, SMEM, st>>>(nx, ny, n_det, n_ang, nblk, geom, x, y, coef_host, \
                                                                             coef_dev, z, part, vec4, run_tan, po, cta_order)
#define FW_LAUNCH_W(QT, MB, SMEM)        \
  do {                                   \
    if (warps == 2) FW_LAUNCH(QT, MB, 2, SMEM); \
    else FW_LAUNCH(QT, MB, 4, SMEM);     \
  } while (0)
  if (qbytes <= 96 * 1024) {
    if (qbytes > 48 * 1024) {
      static thread_local int configured_dev = -1;
      int dev = 0;
      cudaGetDevice(&dev);
      if (dev != configured_dev) {
        cudaFuncSetAttribute(ct_forward_rays_kernel<true, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(ct_forward_rays_kernel<true, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(ct_forward_rays_kernel<true, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(ct_forward_rays_kernel<true, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        configured_dev = dev;
      }
    }
    if (g_fw_minb == 8) FW_LAUNCH_W(true, 8, qbytes);
    else FW_LAUNCH_W(true, 4, qbytes);
  } else {
    if (g_fw_minb == 8) FW_LAUNCH_W(false, 8, 0);
    else FW_LAUNCH_W(false, 4, 0);
  }
#undef FW_LAUNCH_W
#undef FW_LAUNCH
  nparts = nctas * warps;
  return check_launch("ct_forward_rays");
}

// CTA shape of the launch for this problem size: *rays_per_cta (64 or 128) and *blocks_per_angle.  A caller that wants the
// CTAs scheduled heaviest first passes cta_order: n_ang * blocks_per_angle int32, entry b = angle * blocks_per_angle + block.
int tb200_ct_forward_rays_plan(int n_det, int n_ang, int* rays_per_cta, int* blocks_per_angle) {
  TB200_REQUIRE(n_det > 0 && n_ang >= 0 && rays_per_cta && blocks_per_angle, "bad argument");
  const int warps = fw_warps_for(n_det, n_ang);
  *rays_per_cta = warps * 32;
  *blocks_per_angle = (n_det + warps * 32 - 1) / (warps * 32);
  return 0;
}

int tb200_ct_forward_rays_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const double* x, double* y,
                              double coef_host, const double* coef_dev, const double* z, double* norm_out, double* ws,
                              const int32_t* cta_order, void* stream) {
  TB200_REQUIRE(norm_out == nullptr || ws != nullptr, "norm_out requires a workspace");
  cudaStream_t st = (cudaStream_t)stream;
  PeerOut po;
  po.n = 0;
  int64_t nparts = 0;
  int rc = forward_rays_launch(nx, ny, n_det, n_ang, geom, x, y, coef_host, coef_dev, z, norm_out ? ws : nullptr, po, cta_order, st,
                               nparts);
  if (rc) return rc;
  if (norm_out && nparts > 0) {
    finalize_dd_kernel<<<1, 1024, 0, st>>>(ws, nparts, norm_out);
    rc = check_launch("ct_forward_rays finalize");
  }
  return rc;
}

// Fan-beam form: y = A x - coef*z for the flat-detector fan-beam matrix of tb200_ctfan_fill_rows (source at distance so,
// detector at dd, bins of width dps; cosv / sinv: the n_ang angles), matrix-free and bit-identical to the stored product.
// ws as tb200_ct_forward_rays_f64.
int tb200_ctfan_forward_rays_f64(double so, double dd, double dps, int nx, int ny, int n_det, int n_ang, const double* cosv,
                                 const double* sinv, const double* x, double* y, double coef_host, const double* coef_dev,
                                 const double* z, double* norm_out, double* ws, void* stream) {
  TB200_REQUIRE(nx > 0 && ny > 0 && n_det > 0 && n_ang >= 0, "bad geometry");
  TB200_REQUIRE((int64_t)nx * ny < ((int64_t)1 << 31) && (int64_t)n_ang * n_det < ((int64_t)1 << 31), "index space exceeds int32");
  TB200_REQUIRE(so > 0.0 && dd >= 0.0 && dps > 0.0 && so * so > 0.25 * ((double)nx * nx + (double)ny * ny),
                "fan beam: the source must lie outside the image");
  TB200_REQUIRE(norm_out == nullptr || ws != nullptr, "norm_out requires a workspace");
  if (n_ang == 0) return 0;
  TB200_REQUIRE(cosv && sinv && x && y, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const Beam bm = {1, so, dd, dps};
  const int nblk = (n_det + FW_WARPS * 32 - 1) / (FW_WARPS * 32);
  const int ranks = nblk + ((nblk & 1) ? 0 : 1);
  const int64_t nctas = (int64_t)ranks * n_ang;
  TB200_REQUIRE(nctas < ((int64_t)1 << 31), "too many CTAs");
  ct_forward_rays_fan_kernel<<<(unsigned)nctas, FW_WARPS * 32, 0, st>>>(bm, nx, ny, n_det, n_ang, nblk, cosv, sinv, x, y, coef_host,
                                                                        coef_dev, z, norm_out ? ws : nullptr,
                                                                        ((uintptr_t)x % 32) == 0, g_fw_run_tan);
  int rc = check_launch("ctfan_forward_rays");
  if (rc) return rc;
  if (norm_out) {
    finalize_dd_kernel<<<1, 1024, 0, st>>>(ws, nctas * FW_WARPS, norm_out);
    rc = check_launch("ctfan_forward_rays finalize");
  }
  return rc;
}

// Sharded form (trips-py_b200/dist.py): this rank projects ITS angles (geom: its n_ang angles) from the whole image x.
// y (the rank's chunk of its own copy of the gathered sinogram) receives the rows; the same values are stored into
// peers[0 .. n_peers) (the same chunk of the other ranks' copies) from the epilogue, followed by a system-scope fence.
// partials: tb200_ct_forward_rays_workspace_len doubles, to be summed over the ranks by tb200_comm_allreduce_dd.
int tb200_ct_forward_rays_sharded_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const double* x, double* y,
                                      double* const* peers_host, int n_peers, double coef_host, const double* coef_dev,
                                      const double* z, double* partials, int64_t* n_partials_out, const int32_t* cta_order,
                                      void* stream) {
  TB200_REQUIRE(partials && n_partials_out, "null pointer");
  TB200_REQUIRE(n_peers >= 0 && n_peers <= 15 && (n_peers == 0 || peers_host), "bad peer list");
  PeerOut po;
  po.n = n_peers;
  for (int i = 0; i < n_peers; ++i) po.p[i] = peers_host[i];
  return forward_rays_launch(nx, ny, n_det, n_ang, geom, x, y, coef_host, coef_dev, z, partials, po, cta_order,
                             (cudaStream_t)stream, *n_partials_out);
}

}  // extern "C"
