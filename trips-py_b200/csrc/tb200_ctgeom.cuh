// Parallel-beam line-model geometry shared by the matrix builder (ct_builder.cu) and the kernels that re-evaluate
// matrix entries on the fly (ct_project.cu).  ONE definition of an entry, every step a separately rounded IEEE
// operation, so stored and recomputed values are the same bits (oracle statement: oracle/trips_oracle.py ct_matrix).
//
// For a ray with unit normal (c, s) at signed distance t from the pixel centre the chord length through a unit
// square is the trapezoid   len(t) = 1/hi                     |t| <= (hi-lo)/2
//                                   ((hi+lo)/2 - |t|)/(hi*lo)  (hi-lo)/2 < |t| < (hi+lo)/2,   hi/lo = max/min(|c|,|s|)
// evaluated as   min(1/hi, (d2 - |t|) * (1/(hi*lo))),   d2 = (hi+lo)/2:   the plateau is where the slope line exceeds
// it.  No division and no branch per entry - a handful of fp64 instructions instead of 8 streamed bytes.
// (lo == 0: 1/(hi*lo) = inf and the min returns the plateau for every |t| < d2.)
#pragma once
#include <cstdint>

namespace tb200 {

struct RayGeom {
  double c, s, d2, inv_hi, inv_hilo;
};

__device__ __forceinline__ RayGeom make_geom(double c, double s) {
  RayGeom g;
  g.c = c;
  g.s = s;
  const double ac = fabs(c), as = fabs(s);
  const double hi = fmax(ac, as), lo = fmin(ac, as);
  g.d2 = __dmul_rn(0.5, __dadd_rn(hi, lo));
  g.inv_hi = __ddiv_rn(1.0, hi);
  g.inv_hilo = __ddiv_rn(1.0, __dmul_rn(hi, lo));
  return g;
}

// signed distance (along the detector axis) between ray offset sd and the projection of pixel centre (cx, cy)
__device__ __forceinline__ double ray_pixel_t(const RayGeom& g, double sd, double cx, double cy) {
  return __dsub_rn(sd, __dadd_rn(__dmul_rn(cx, g.c), __dmul_rn(cy, g.s)));
}
__device__ __forceinline__ bool hits(const RayGeom& g, double t) { return fabs(t) < g.d2; }
// chord length for |t| < d2 (meaningless outside the footprint)
// (a compare + select, not fmin(): there is no NaN to honour inside the footprint and fmin() costs four extra
// integer instructions per entry on sm_100)
__device__ __forceinline__ double chord_from_margin(double margin, double inv_hi, double inv_hilo) {
  const double s = __dmul_rn(margin, inv_hilo);  // margin = d2 - |t| > 0
  return (s < inv_hi) ? s : inv_hi;
}
__device__ __forceinline__ double chord(const RayGeom& g, double t) {
  return chord_from_margin(__dsub_rn(g.d2, fabs(t)), g.inv_hi, g.inv_hilo);
}

// Beam geometry.  fan == 0: parallel beam, ray (angle, d) has normal (cos, sin) and offset d - (n_det-1)/2.
// fan == 1: flat-detector fan beam in the conventions of ASTRA's 'fanflat' geometry that the reference builds
// (trips/test_problems/Tomography.py:57-67): source at so*(sin, -cos), detector centre at dd*(-sin, cos), detector
// axis (cos, sin), bins of width dps.  The ray from the source to the centre of bin d is a line with unit normal
// (c, s) and signed offset rho = (c, s).source; the entry is the same chord function of t = rho - (c, s).pixel.
struct Beam {
  int fan;
  double so, dd, dps;
};

__device__ __forceinline__ void ray_geometry(const Beam& bm, double cosa, double sina, int d, int n_det, RayGeom& g,
                                             double& offset) {
  const double k = (double)d - 0.5 * (double)(n_det - 1);
  if (!bm.fan) {
    g = make_geom(cosa, sina);
    offset = k;
    return;
  }
  const double sx = __dmul_rn(bm.so, sina), sy = -__dmul_rn(bm.so, cosa);
  const double px = __dadd_rn(-__dmul_rn(bm.dd, sina), __dmul_rn(k, __dmul_rn(cosa, bm.dps)));
  const double py = __dadd_rn(__dmul_rn(bm.dd, cosa), __dmul_rn(k, __dmul_rn(sina, bm.dps)));
  const double ex = __dsub_rn(px, sx), ey = __dsub_rn(py, sy);
  const double len = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
  const double c = __ddiv_rn(ey, len), s = -__ddiv_rn(ex, len);
  g = make_geom(c, s);
  offset = __dadd_rn(__dmul_rn(c, sx), __dmul_rn(s, sy));
}

// k - 0.5*(n-1) for 0 <= k < 2^31, exact (both operands are multiples of 0.5 well inside 2^52): the integer is
// dropped into the mantissa of 2^51 (ulp 0.5) and one subtraction removes the bias together with the centre offset.
// Equal to `(double)k - 0.5*(double)(n-1)` (which is exact too) without the int->fp64 conversion.
__device__ __forceinline__ double centred_coord(int k, double bias_plus_half_n1) {
  return __dsub_rn(__hiloint2double(0x43200000, k << 1), bias_plus_half_n1);
}
__device__ __forceinline__ double centred_bias(int n) { return 2251799813685248.0 + 0.5 * (double)(n - 1); }  // 2^51 + (n-1)/2

}  // namespace tb200
