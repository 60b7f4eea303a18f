// Matrix-free parallel-beam back-projection  y = A^T u - coef*z  for the CT matrix of ct_builder.cu, bit-identical to
// the sequential-order SpMV on the stored transpose (and so to scipy's A.T @ u on the same matrix).
//
// Role in the reference: the tomography operator is matrix-free there too (astra.OpTomo behind a pylops
// FunctionOperator, trips/test_problems/Tomography.py:73-83); SURVEY.md section 8(f) item 1.
//
// Why: streaming the stored transpose costs 12 bytes per entry and bounds a Golub-Kahan step at the HBM roofline.
// An entry of this matrix is a closed-form function of (pixel, angle, detector) - tb200_ctgeom.cuh - and a pixel
// meets at most TWO detector bins per angle (footprint (|c|+|s|) <= sqrt(2) bins wide), always the two that bracket
// its projection.  So one thread per pixel walks the angles in order, evaluates the two candidate entries
// (~20 fp64 instructions per angle), gathers the two neighbouring sinogram samples (L1/L2 resident: the sinogram is
// 16.7 MB at 2048^2 x 720) and adds the products that lie inside the footprint, in ascending (angle, detector) order -
// exactly the row order of the stored transpose.  No matrix bytes are read at all; the kernel is fp64-pipe bound.
//
// Pattern equivalence.  Stored pattern of pixel p at angle a: { d in [0, n_det) : |t_d| < d2 },  t_d = (d - dc) - proj.
// With d0 = any integer within 1 of floor(proj + dc), every d outside {d0, d0+1} that is not one of the bracketing
// pair has |t_d| >= 1 - 1e-12 > d2 (d2 <= 0.7072), so testing the pair {d0, d0+1} with the SAME predicate on the SAME
// separately rounded t_d reproduces the pattern; rounding in the estimate of d0 can only swap in a neighbour that fails
// the predicate.
#include "tb200_common.cuh"
#include "tb200_ctgeom.cuh"
#include "tb200_dd.cuh"

namespace tb200 {

__global__ void ct_geometry_kernel(int n_ang, const double* __restrict__ cosv, const double* __restrict__ sinv,
                                   double* __restrict__ geom) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_ang) return;
  const RayGeom g = make_geom(cosv[a], sinv[a]);
  double* o = geom + 6 * (int64_t)a;
  o[0] = g.c, o[1] = g.s, o[2] = g.d2, o[3] = g.inv_hi, o[4] = g.inv_hilo, o[5] = 0.0;
}

constexpr int BP_TA = 256;  // angles per shared-memory tile of the geometry table (12 KB)

// acc + p when flag > 0, else acc: one predicated DADD (the C form compiles to an add and two selects)
__device__ __forceinline__ double add_if_positive(double acc, double p, int flag) {
  asm("{\n\t.reg .pred q;\n\tsetp.gt.s32 q, %2, 0;\n\t@q add.rn.f64 %0, %0, %1;\n\t}" : "+d"(acc) : "d"(p), "r"(flag));
  return acc;
}

// One tile of angles for the two pixels of a thread ((ix, iy) and (ix, iy + 4): they share cx*c, the geometry loads and
// the loop overhead).  CHECKED = false is the path of warps whose pixels project at least two bins inside the detector at
// every angle (all but the image corners): no bounds tests, unconditional gathers.
// OFFS: the sinogram rows of angle j start at the int64 stored (bit pattern) in gtab[6j + 5] instead of at
// (a0 + j) * n_det - the sharded layout, where the angles of a gathered sinogram are grouped by owner rank.
// FOLD (even detector counts): dc - 1/2 is an integer, MAGIC + (dc - 1/2) is exact and the two additions that round
// proj + dc - 1/2 to an integer fold into one (same integer except on ties of the first rounding, where either neighbour
// is a valid bracket origin - see the header).
template <bool CHECKED, int UNROLL, bool OFFS, bool FOLD>
__device__ __forceinline__ void bp_tile(double (&acc)[2], const double* __restrict__ gtab, int na, const double* __restrict__ u,
                                        int row0, int n_det, double cx, const double (&cy)[2], double dc, double kmagic,
                                        uint64_t pol_keep) {
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: (v + MAGIC) - MAGIC = v rounded to an integer
  const double dcm = dc - 0.5;               // exact: multiples of 0.5
  const unsigned und = (unsigned)n_det;
#pragma unroll UNROLL
  for (int j = 0; j < na; ++j) {
    const double2 cs = reinterpret_cast<const double2*>(gtab)[3 * j];      // c, s
    const double2 dh = reinterpret_cast<const double2*>(gtab)[3 * j + 1];  // d2, 1/hi
    const double inv_hilo = gtab[6 * j + 4];
    const int rowj = OFFS ? (int)__double_as_longlong(gtab[6 * j + 5]) : row0 + j * n_det;
    const double pc = __dmul_rn(cx, cs.x);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const double proj = __dadd_rn(pc, __dmul_rn(cy[q], cs.y));
      // d0 = round(proj + dc - 1/2) = floor(proj + dc) up to ties, as MAGIC + d0 in one double
      const double w = FOLD ? __dadd_rn(proj, kmagic) : __dadd_rn(__dadd_rn(proj, dcm), MAGIC);
      const int d0 = __double2loint(w);  // low word of MAGIC is 0: the integer, in two's complement
      // (d - dc) for d0 and d0 + 1: w - MAGIC is the integer d0 as a double (exact), minus the half-integer dc (exact)
      const double sd0 = __dsub_rn(__dsub_rn(w, MAGIC), dc);
      const double sd1 = __dadd_rn(sd0, 1.0);
      const double e0 = __dsub_rn(dh.x, fabs(__dsub_rn(sd0, proj)));  // d2 - |t|: positive inside the footprint
      const double e1 = __dsub_rn(dh.x, fabs(__dsub_rn(sd1, proj)));
      // e > 0 (never denormal here: |t| and d2 are O(1)) <=> the high word, read as an int, is positive
      if (CHECKED) {
        const double* up = u + ((int64_t)rowj + d0);
        const bool in0 = (unsigned)d0 < und, in1 = (unsigned)(d0 + 1) < und;
        const double u0 = in0 ? ld_gather_f64(up, pol_keep) : 0.0;
        const double u1 = in1 ? ld_gather_f64(up + 1, pol_keep) : 0.0;
        const double p0 = __dmul_rn(chord_from_margin(e0, dh.y, inv_hilo), u0);
        const double p1 = __dmul_rn(chord_from_margin(e1, dh.y, inv_hilo), u1);
        if (in0 && __double2hiint(e0) > 0) acc[q] = __dadd_rn(acc[q], p0);
        if (in1 && __double2hiint(e1) > 0) acc[q] = __dadd_rn(acc[q], p1);
      } else {
        const double* up = u + (unsigned)(rowj + d0);  // n_ang * n_det < 2^31 is checked at launch
        const double p0 = __dmul_rn(chord_from_margin(e0, dh.y, inv_hilo), ld_gather_f64(up, pol_keep));
        const double p1 = __dmul_rn(chord_from_margin(e1, dh.y, inv_hilo), ld_gather_f64(up + 1, pol_keep));
        acc[q] = add_if_positive(acc[q], p0, __double2hiint(e0));
        acc[q] = add_if_positive(acc[q], p1, __double2hiint(e1));
      }
    }
  }
}

// CTA = 4 warps side by side, each warp two 8 x 4 pixel tiles one above the other (32 x 8 pixels per CTA, two pixels per
// thread): a compact tile keeps the two gather requests of an angle within 2-3 sectors whatever the angle.
constexpr int BP_ROWS = 8;  // image rows per CTA
#ifndef BP_MIN_CTAS
#define BP_MIN_CTAS 6  // measured: 7 CTAs per SM (72 registers) is 2 % slower, 8 spills
#endif

template <int UNROLL, bool OFFS, bool FOLD>
__global__ void __launch_bounds__(128, BP_MIN_CTAS)
ct_backproject_kernel(int nx, int ny, int iy_begin, int iy_end, int n_det, int n_ang, const double* __restrict__ geom,
                      const double* __restrict__ u, double* __restrict__ y, double coef_host,
                      const double* __restrict__ coef_dev, const double* __restrict__ z, int64_t z_offset,
                      double* __restrict__ partials, PeerOut po) {
  __shared__ __align__(16) double gtab[BP_TA * 6];
  __shared__ double red[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ix = blockIdx.x * 32 + warp * 8 + (lane & 7);
  const int iy0 = iy_begin + blockIdx.y * BP_ROWS + (lane >> 3);
  const uint64_t pol_keep = policy_evict_last();
  const double cx = (double)ix - 0.5 * (double)(nx - 1);
  const double cy[2] = {(double)iy0 - 0.5 * (double)(ny - 1), (double)(iy0 + 4) - 0.5 * (double)(ny - 1)};
  const double dc = 0.5 * (double)(n_det - 1);
  const double kmagic = 6755399441055744.0 + (dc - 0.5);  // FOLD (even n_det): dc - 1/2 is an integer, the sum is exact
  // |proj| <= |(cx, cy)|: with two bins of slack every candidate of every angle is a valid detector index
  const double cymax = fmax(fabs(cy[0]), fabs(cy[1]));
  const bool interior = __all_sync(0xffffffffu, sqrt(cx * cx + cymax * cymax) + 2.5 <= dc);
  double acc[2] = {0.0, 0.0};

  for (int a0 = 0; a0 < n_ang; a0 += BP_TA) {
    const int na = min(BP_TA, n_ang - a0);
    __syncthreads();
    for (int i = threadIdx.x; i < na * 3; i += 128)
      reinterpret_cast<double2*>(gtab)[i] = reinterpret_cast<const double2*>(geom + 6 * (int64_t)a0)[i];
    __syncthreads();
    if (interior) bp_tile<false, UNROLL, OFFS, FOLD>(acc, gtab, na, u, a0 * n_det, n_det, cx, cy, dc, kmagic, pol_keep);
    else bp_tile<true, UNROLL, OFFS, FOLD>(acc, gtab, na, u, a0 * n_det, n_det, cx, cy, dc, kmagic, pol_keep);
  }

  const double coef = (z != nullptr) ? (coef_dev ? *coef_dev : coef_host) : 0.0;
  dd_t nrm = dd_zero();
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int iy = iy0 + 4 * q;
    if (ix < nx && iy < iy_end) {
      const int64_t pix = (int64_t)iy * nx + ix;
      double v = acc[q];
      if (z != nullptr) v = __dsub_rn(v, __dmul_rn(coef, z[pix - z_offset]));
      y[pix] = v;
#pragma unroll
      for (int p = 0; p < 15; ++p)  // the other GPUs' copies of the vector, over NVLink (unrolled: parameters stay in the constant bank)
        if (p < po.n) po.p[p][pix] = v;
      nrm = dd_fma(nrm, v, v);
    }
  }
  if (po.n > 0) __threadfence_system();
  if (partials != nullptr) {
    const dd_t tot2 = dd_block_sum(nrm, red);
    if (threadIdx.x == 0) {
      const int64_t b = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
      partials[2 * b] = tot2.hi;
      partials[2 * b + 1] = tot2.lo;
    }
  }
}

}  // namespace tb200

using namespace tb200;

extern "C" {

// geom[6*a .. 6*a+5] = (c, s, d2, 1/hi, 1/(hi*lo), 0) for angle a: the per-angle constants of the entry function.
int tb200_ct_geometry(int n_ang, const double* cosv, const double* sinv, double* geom, void* stream) {
  TB200_REQUIRE(n_ang >= 0, "bad size");
  if (n_ang == 0) return 0;
  TB200_REQUIRE(cosv && sinv && geom, "null pointer");
  TB200_REQUIRE(((uintptr_t)geom % 16) == 0, "geom must be 16-byte aligned");
  ct_geometry_kernel<<<(n_ang + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_ang, cosv, sinv, geom);
  return check_launch("ct_geometry");
}

// Doubles of workspace tb200_ct_backproject_f64 needs for its fused norm (one double-double partial per CTA).
int64_t tb200_ct_backproject_workspace_len(int nx, int ny) {
  return 2 * ((int64_t)((nx + 31) / 32) * ((ny + 3) / 4) + 8);
}

// y = A^T u - coef*z (z nullable; coef from coef_dev if non-null), optional norm_out = (||y||^2, ||y||), for the
// parallel-beam matrix of n_ang angles (geom from tb200_ct_geometry), u angle-major (angle*n_det + detector),
// y row-major (iy*nx + ix).  No matrix is read.  Same bits as tb200_spmv_sell_f64 on tb200_ct_fill_cols' matrix.
// The _rows form computes only image rows [iy_begin, iy_end) (y, z still indexed by the global pixel number; the norm
// covers those rows): a sharded caller back-projects the image in bands and all-reduces each band while the next
// one is being computed.
int tb200_ct_backproject_rows_f64(int nx, int ny, int iy_begin, int iy_end, int n_det, int n_ang, const double* geom,
                                  const double* u, double* y, double coef_host, const double* coef_dev, const double* z,
                                  double* norm_out, double* ws, void* stream);

int tb200_ct_backproject_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const double* u, double* y,
                             double coef_host, const double* coef_dev, const double* z, double* norm_out, double* ws,
                             void* stream) {
  return tb200_ct_backproject_rows_f64(nx, ny, 0, ny, n_det, n_ang, geom, u, y, coef_host, coef_dev, z, norm_out, ws, stream);
}

static int backproject_launch(int nx, int ny, int iy_begin, int iy_end, int n_det, int n_ang, const double* geom, const double* u,
                              double* y, double coef_host, const double* coef_dev, const double* z, int64_t z_offset,
                              double* partials, bool offsets, const PeerOut& po, cudaStream_t st, dim3& grid) {
  TB200_REQUIRE(nx > 0 && ny > 0 && n_det > 0 && n_ang >= 0, "bad geometry");
  TB200_REQUIRE(0 <= iy_begin && iy_begin <= iy_end && iy_end <= ny, "bad row band");
  TB200_REQUIRE((int64_t)nx * ny < ((int64_t)1 << 31) && (int64_t)n_ang * n_det < ((int64_t)1 << 31), "index space exceeds int32");
  TB200_REQUIRE(y && (n_ang == 0 || (geom && u)), "null pointer");
  TB200_REQUIRE(((uintptr_t)geom % 16) == 0, "geom must be 16-byte aligned");
  grid = dim3((unsigned)((nx + 31) / 32), (unsigned)((iy_end - iy_begin + BP_ROWS - 1) / BP_ROWS));
  TB200_REQUIRE(grid.y <= 65535u, "ny too large for this launch shape");
#define BP_LAUNCH(OFFS, FOLD)                                                                                             \
  ct_backproject_kernel<4, OFFS, FOLD><<<grid, 128, 0, st>>>(nx, ny, iy_begin, iy_end, n_det, n_ang, geom, u, y, coef_host, coef_dev, \
                                                             z, z_offset, partials, po)
  const bool fold = (n_det & 1) == 0;
  if (offsets) {
    if (fold) BP_LAUNCH(true, true);
    else BP_LAUNCH(true, false);
  } else {
    if (fold) BP_LAUNCH(false, true);
    else BP_LAUNCH(false, false);
  }
#undef BP_LAUNCH
  return check_launch("ct_backproject");
}

int tb200_ct_backproject_rows_f64(int nx, int ny, int iy_begin, int iy_end, int n_det, int n_ang, const double* geom,
                                  const double* u, double* y, double coef_host, const double* coef_dev, const double* z,
                                  double* norm_out, double* ws, void* stream) {
  TB200_REQUIRE(norm_out == nullptr || ws != nullptr, "norm_out requires a workspace");
  if (iy_begin == iy_end) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid;
  PeerOut po;
  po.n = 0;
  int rc = backproject_launch(nx, ny, iy_begin, iy_end, n_det, n_ang, geom, u, y, coef_host, coef_dev, z, 0, norm_out ? ws : nullptr,
                              false, po, st, grid);
  if (rc) return rc;
  if (norm_out) {
    finalize_dd_kernel<<<1, 1024, 0, st>>>(ws, (int64_t)grid.x * grid.y, norm_out);
    rc = check_launch("ct_backproject finalize");
  }
  return rc;
}

// Sharded form (trips-py_b200/dist.py): this rank back-projects image rows [iy_begin, iy_end) from the WHOLE sinogram.
//  * u holds all ranks' angles grouped by owner; geom[6a + 5] carries (as an int64 bit pattern) the row at which global
//    angle a starts in u, so the pixel sums still run over the angles in global order: same bits as one GPU.
//  * y is indexed by the global pixel number (the rank's copy of the full-length vector); the same values are stored
//    into peers[0 .. n_peers) (the other ranks' copies) from the epilogue, followed by a system-scope fence.
//  * z (nullable) is the rank's slice of the previous basis vector: z[pix - iy_begin*nx].
//  * partials: 2 doubles per CTA (tb200_ct_backproject_workspace_len), to be summed over the ranks by
//    tb200_comm_allreduce_dd; *n_partials_out receives their number.
int tb200_ct_backproject_sharded_f64(int nx, int ny, int iy_begin, int iy_end, int n_det, int n_ang, const double* geom,
                                     const double* u, double* y, double* const* peers_host, int n_peers, double coef_host,
                                     const double* coef_dev, const double* z, double* partials, int64_t* n_partials_out,
                                     void* stream) {
  TB200_REQUIRE(partials && n_partials_out, "null pointer");
  TB200_REQUIRE(n_peers >= 0 && n_peers <= 15 && (n_peers == 0 || peers_host), "bad peer list");
  PeerOut po;
  po.n = n_peers;
  for (int i = 0; i < n_peers; ++i) po.p[i] = peers_host[i];
  dim3 grid(0, 0);
  *n_partials_out = 0;
  if (iy_begin == iy_end) return 0;
  int rc = backproject_launch(nx, ny, iy_begin, iy_end, n_det, n_ang, geom, u, y, coef_host, coef_dev, z, (int64_t)iy_begin * nx,
                              partials, true, po, (cudaStream_t)stream, grid);
  *n_partials_out = (int64_t)grid.x * grid.y;
  return rc;
}

// One Golub-Kahan step (the reference's golub_kahan_update, trips/utilities/decompositions.py:230-255) on the
// matrix-free CT operator, enqueued as 6 kernels (7 with the image transpose) on `stream`, every scalar on the device:
//   v = A^T u_k - beta_prev * v_prev ; alpha = ||v|| ; v /= alpha ; u = A v - alpha * u_k ; beta = ||u|| ; u /= beta
// Arguments as tb200_gk_step_sell_f64, the operator as tb200_ct_forward_f64 / tb200_ct_backproject_f64.
// colidx == NULL selects the ray-driven forward projector (tb200_ct_forward_rays_f64): sliceptr, rowlen, rowskip, xT_scratch
// unused; cta_order then is that projector's (nullable) heaviest-first CTA list.
// ws: max(tb200_spmv_workspace_len(n_ang*n_det), tb200_ct_backproject_workspace_len(nx, ny),
//         tb200_ct_forward_rays_workspace_len(n_det, n_ang)) doubles.
int tb200_vec_div(int64_t n, const double* x, double d_host, const double* d_dev, double* out, void* stream);
int tb200_ct_forward_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const int64_t* sliceptr,
                         const int32_t* rowlen, const int32_t* rowskip, const int32_t* colidx, const int32_t* cta_order,
                         double* xT_scratch, const double* x, double* y, double coef_host, const double* coef_dev,
                         const double* z, double* norm_out, double* ws, void* stream);

int tb200_ct_forward_rays_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const double* x, double* y,
                              double coef_host, const double* coef_dev, const double* z, double* norm_out, double* ws,
                              const int32_t* cta_order, void* stream);

int tb200_gk_step_ct_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const int64_t* sliceptr,
                         const int32_t* rowlen, const int32_t* rowskip, const int32_t* colidx, const int32_t* cta_order,
                         double* xT_scratch, const double* u_k, const double* v_prev, const double* beta_prev_dev,
                         double* v_out, double* u_out, double* alpha_pair, double* beta_pair, double* ws,
                         void* const* events_host, void* stream) {
  TB200_REQUIRE(u_k && v_out && u_out && alpha_pair && beta_pair && ws, "null pointer");
  TB200_REQUIRE((v_prev == nullptr) == (beta_prev_dev == nullptr), "v_prev and beta_prev_dev go together");
  cudaStream_t st = (cudaStream_t)stream;
  auto mark = [&](int i) {
    if (events_host != nullptr && events_host[i] != nullptr) cudaEventRecord((cudaEvent_t)events_host[i], st);
  };
  const int64_t m = (int64_t)n_ang * n_det, n = (int64_t)nx * ny;
  mark(0);
  int rc = tb200_ct_backproject_f64(nx, ny, n_det, n_ang, geom, u_k, v_out, 0.0, beta_prev_dev, v_prev, alpha_pair, ws, stream);
  mark(1);
  if (rc) return rc;
  rc = tb200_vec_div(n, v_out, 0.0, alpha_pair + 1, v_out, stream);
  if (rc) return rc;
  mark(2);
  if (colidx == nullptr)  // fully matrix-free: the ray-driven forward projector (ct_forward.cu), no index arrays
    rc = tb200_ct_forward_rays_f64(nx, ny, n_det, n_ang, geom, v_out, u_out, 0.0, alpha_pair + 1, u_k, beta_pair, ws, cta_order, stream);
  else
    rc = tb200_ct_forward_f64(nx, ny, n_det, n_ang, geom, sliceptr, rowlen, rowskip, colidx, cta_order, xT_scratch, v_out, u_out, 0.0,
                              alpha_pair + 1, u_k, beta_pair, ws, stream);
  mark(3);
  if (rc) return rc;
  return tb200_vec_div(m, u_out, 0.0, beta_pair + 1, u_out, stream);
}

}  // extern "C"
