// Tall-skinny basis kernels: the (re)orthogonalisation and lift GEMVs of the Krylov cores.
//
// Reference call sites (all NumPy/OpenBLAS on the host there):
//   h = V.T @ r ; r = r - V @ h (x2 in MMGKS, x3 in GKS)   trips/solvers/MMGKS.py:119-120, GKS.py:86-88
//   MGS loop h_j = v_j . w ; w -= h_j v_j                   trips/utilities/decompositions.py:216-218
//   x = V @ y, AV @ y, LV @ y (lifts)                        Hybrid_LSQR.py:105, Hybrid_GMRES.py:77, GKS.py:76-83
//   U.T @ b                                                  trips/utilities/reg_param/discrepancy_principle.py:34
//
// Layout: a basis is a pre-allocated buffer of kmax columns, column j contiguous at V + j*ld (ld >= n).
// Appending a vector is a pointer bump (the reference re-copies the whole basis with np.hstack every step).
//
// Both kernels are pure HBM streams of V (8*n*k bytes); nothing here is a dense contraction worth tensor
// cores.  basis_dots streams V once for h = V^T w with the w segment resident in L1; basis_combine streams V
// once for  out = w + sign * V h  with the squared norm of the result fused into the same pass.
#include "tb200_common.cuh"
#include "tb200_dd.cuh"

namespace tb200 {

constexpr int kBThreads = 256;
constexpr int kBMaxBlocks = 1184;  // 148 x 8; fixed => the reduction tree depends on (n, k) only
constexpr int kJT = 8;             // columns per register tile in basis_dots

__device__ __forceinline__ double ld_stream_f64(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

static inline int basis_grid(int64_t n) { return grid_for(n, kBThreads * 4, kBMaxBlocks); }

// partials[j * gridDim.x + blockIdx.x] = sum over this CTA's segment of V[:, j] * w
__global__ void __launch_bounds__(kBThreads)
basis_dots_kernel(int64_t n, int k, const double* __restrict__ V, int64_t ld, const double* __restrict__ w,
                  double* __restrict__ partials) {
  __shared__ double red[kJT][kBThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // contiguous segment per CTA so the w re-reads (once per column tile) stay in this SM's L1
  const int64_t seg = ((n + gridDim.x - 1) / gridDim.x + kBThreads - 1) / kBThreads * kBThreads;
  const int64_t lo = (int64_t)blockIdx.x * seg;
  int64_t hi = lo + seg;
  if (hi > n) hi = n;
  for (int j0 = 0; j0 < k; j0 += kJT) {
    const int jn = (k - j0 < kJT) ? (k - j0) : kJT;
    double acc[kJT];
#pragma unroll
    for (int t = 0; t < kJT; ++t) acc[t] = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += kBThreads) {
      const double wi = __ldg(w + i);
      if (jn == kJT) {
        double v[kJT];
#pragma unroll
        for (int t = 0; t < kJT; ++t) v[t] = ld_stream_f64(V + (int64_t)(j0 + t) * ld + i);
#pragma unroll
        for (int t = 0; t < kJT; ++t) acc[t] = fma(v[t], wi, acc[t]);
      } else {
        for (int t = 0; t < jn; ++t) acc[t] = fma(ld_stream_f64(V + (int64_t)(j0 + t) * ld + i), wi, acc[t]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < kJT; ++t) {
      const double s = warp_sum(acc[t]);
      if (lane == 0) red[t][wid] = s;
    }
    __syncthreads();
    if (threadIdx.x < jn) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < kBThreads / 32; ++q) s += red[threadIdx.x][q];
      partials[(int64_t)(j0 + threadIdx.x) * gridDim.x + blockIdx.x] = s;
    }
  }
}

// h[j] = sum_b partials[j*nb + b] in fixed order; one CTA per column.
__global__ void __launch_bounds__(256) basis_dots_finalize_kernel(int nb, const double* __restrict__ partials,
                                                                  double* __restrict__ h) {
  __shared__ double red[32];
  const int j = blockIdx.x;
  double acc = 0.0;
  for (int b = threadIdx.x; b < nb; b += 256) acc += partials[(int64_t)j * nb + b];
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) h[j] = tot;
}

// out = (w ? w : 0) + sign * sum_j h[j] V[:, j]; optional partial ||out||^2 per CTA.
// k <= kmax columns of coefficients are staged in shared memory.
__global__ void __launch_bounds__(kBThreads)
basis_combine_kernel(int64_t n, int k, const double* __restrict__ V, int64_t ld, const double* __restrict__ h,
                     const double* __restrict__ w, double sign, double* __restrict__ out, double* __restrict__ partials) {
  extern __shared__ double hs[];
  __shared__ double red[64];
  for (int j = threadIdx.x; j < k; j += kBThreads) hs[j] = h[j];
  __syncthreads();
  dd_t nrm = dd_zero();
  for (int64_t i = (int64_t)blockIdx.x * kBThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBThreads) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j = 0;
    for (; j + 4 <= k; j += 4) {
      const double v0 = ld_stream_f64(V + (int64_t)(j + 0) * ld + i);
      const double v1 = ld_stream_f64(V + (int64_t)(j + 1) * ld + i);
      const double v2 = ld_stream_f64(V + (int64_t)(j + 2) * ld + i);
      const double v3 = ld_stream_f64(V + (int64_t)(j + 3) * ld + i);
      a0 = fma(v0, hs[j + 0], a0);
      a1 = fma(v1, hs[j + 1], a1);
      a2 = fma(v2, hs[j + 2], a2);
      a3 = fma(v3, hs[j + 3], a3);
    }
    for (; j < k; ++j) a0 = fma(ld_stream_f64(V + (int64_t)j * ld + i), hs[j], a0);
    const double vh = (a0 + a1) + (a2 + a3);
    double r = (sign < 0.0) ? -vh : vh;
    if (w != nullptr) r = (sign < 0.0) ? __dsub_rn(w[i], vh) : __dadd_rn(w[i], vh);
    out[i] = r;
    if (partials != nullptr) nrm = dd_fma(nrm, r, r);
  }
  if (partials != nullptr) {
    const dd_t tot = dd_block_sum(nrm, red);
    if (threadIdx.x == 0) {
      partials[2 * blockIdx.x] = tot.hi;
      partials[2 * blockIdx.x + 1] = tot.lo;
    }
  }
}

}  // namespace tb200

using namespace tb200;

extern "C" {

// Doubles of workspace needed by tb200_basis_dots for k columns (and by basis_combine's fused norm).
int64_t tb200_basis_workspace_len(int64_t k) { return (int64_t)kBMaxBlocks * (k > 2 ? k : 2); }

// h[0..k) = V[:, 0..k)^T w.  Deterministic (fixed two-stage tree).  2 launches.
int tb200_basis_dots(int64_t n, int64_t k, const double* V, int64_t ld, const double* w, double* h, double* ws,
                     void* stream) {
  TB200_REQUIRE(n >= 0 && k >= 0 && ld >= n, "bad size");
  if (k == 0) return 0;
  TB200_REQUIRE(V && w && h && ws, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = basis_grid(n);
  basis_dots_kernel<<<g, kBThreads, 0, st>>>(n, (int)k, V, ld, w, ws);
  int rc = check_launch("basis_dots");
  if (rc) return rc;
  basis_dots_finalize_kernel<<<(unsigned)k, 256, 0, st>>>(g, ws, h);
  return check_launch("basis_dots finalize");
}

// out = w + sign * V[:, 0..k) h  (w may be NULL => out = sign * V h, the lift x = V y).
// If norm_out != NULL: norm_out[0] = ||out||^2, norm_out[1] = ||out||.  out may alias w.
int tb200_basis_combine(int64_t n, int64_t k, const double* V, int64_t ld, const double* h, const double* w, double sign,
                        double* out, double* norm_out, double* ws, void* stream) {
  TB200_REQUIRE(n >= 0 && k >= 0 && ld >= n && out, "bad argument");
  TB200_REQUIRE(k == 0 || (V && h), "null pointer");
  TB200_REQUIRE(norm_out == nullptr || ws != nullptr, "norm_out requires a workspace");
  TB200_REQUIRE(k * 8 <= 48 * 1024, "k too large for the coefficient stage");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = basis_grid(n);
  basis_combine_kernel<<<g, kBThreads, (size_t)k * sizeof(double), st>>>(n, (int)k, V, ld, h, w, sign, out,
                                                                         norm_out ? ws : nullptr);
  int rc = check_launch("basis_combine");
  if (rc || !norm_out) return rc;
  finalize_dd_kernel<<<1, 1024, 0, st>>>(ws, g, norm_out);
  return check_launch("basis_combine finalize");
}

}  // extern "C"
