// Double-double accumulation helpers (device).  Used so that every reduction the Krylov recurrences feed back on
// (norms, dots) is the CORRECTLY ROUNDED value of the exact sum: that makes the result independent of the reduction
// tree, of the grid size and of the number of GPUs, and lets the CPU oracle reproduce it bit for bit
// (math.fsum over exact products).  OpenBLAS' ddot, which the reference uses through np.linalg.norm, has a
// machine- and thread-count-dependent summation order and cannot be matched bit for bit by anything.
#pragma once
#include "tb200_common.cuh"

namespace tb200 {

struct dd_t {
  double hi, lo;
};

__device__ __forceinline__ dd_t dd_zero() { return dd_t{0.0, 0.0}; }

__device__ __forceinline__ dd_t dd_add(dd_t a, dd_t b) {
  // TwoSum on the high parts, then fold the low parts and renormalise (FastTwoSum)
  const double s = __dadd_rn(a.hi, b.hi);
  const double bb = __dsub_rn(s, a.hi);
  double e = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(b.hi, bb));
  e = __dadd_rn(e, __dadd_rn(a.lo, b.lo));
  const double hi = __dadd_rn(s, e);
  const double lo = __dsub_rn(e, __dsub_rn(hi, s));
  return dd_t{hi, lo};
}

// acc += x*y exactly (TwoProduct via FMA)
__device__ __forceinline__ dd_t dd_fma(dd_t acc, double x, double y) {
  const double p = __dmul_rn(x, y);
  const double pe = __fma_rn(x, y, -p);
  return dd_add(acc, dd_t{p, pe});
}

__device__ __forceinline__ dd_t dd_warp_sum(dd_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dd_t w;
    w.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
    w.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
    v = dd_add(v, w);
  }
  return v;
}

// Block reduction; smem needs 64 doubles; result valid in thread 0.
__device__ __forceinline__ dd_t dd_block_sum(dd_t v, double* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = dd_warp_sum(v);
  __syncthreads();
  if (lane == 0) {
    smem[2 * wid] = v.hi;
    smem[2 * wid + 1] = v.lo;
  }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? dd_t{smem[2 * threadIdx.x], smem[2 * threadIdx.x + 1]} : dd_zero();
  if (wid == 0) v = dd_warp_sum(v);
  return v;
}

// Single-CTA finalize of per-CTA double-double partials (hi, lo interleaved): out[0] = round(sum), out[1] = sqrt(out[0]).
static __global__ void __launch_bounds__(1024) finalize_dd_kernel(const double* __restrict__ partials, int64_t n,
                                                                  double* __restrict__ out) {
  __shared__ double red[64];
  dd_t acc = dd_zero();
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc = dd_add(acc, dd_t{partials[2 * i], partials[2 * i + 1]});
  const dd_t tot = dd_block_sum(acc, red);
  if (threadIdx.x == 0) {
    out[0] = tot.hi;
    out[1] = sqrt(tot.hi);
  }
}

}  // namespace tb200
