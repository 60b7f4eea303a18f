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

template <int VEC>
struct RowVec;
template <>
struct RowVec<1> {
  double v[1];
  __device__ __forceinline__ void load_stream(const double* p) { v[0] = ld_stream_f64(p); }
  __device__ __forceinline__ void load(const double* p) { v[0] = __ldg(p); }
};
template <>
struct RowVec<2> {  // two consecutive rows, one 16-byte access (the column must be 16-byte aligned)
  double v[2];
  __device__ __forceinline__ void load_stream(const double* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
  }
  __device__ __forceinline__ void load(const double* p) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(p));
    v[0] = t.x, v[1] = t.y;
  }
};

// partials[j * gridDim.x + blockIdx.x] = sum over this CTA's segment of V[:, j] * w.
// VEC rows per thread and step: with VEC = 2 a thread has 8 x 16 bytes of V in flight per step, which is what it takes
// to cover the HBM latency at the 4-8 steps per thread these launches have (VEC = 1 measured 25-45 % of the HBM peak).
template <int VEC>
__global__ void __launch_bounds__(kBThreads)
basis_dots_kernel(int64_t n, int k, const double* __restrict__ V, int64_t ld, const double* __restrict__ w,
                  double* __restrict__ partials) {
  __shared__ double red[kJT][kBThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // contiguous segment per CTA so the w re-reads (once per column tile) stay in this SM's L1
  const int64_t seg = ((n + gridDim.x - 1) / gridDim.x + kBThreads * VEC - 1) / (kBThreads * VEC) * (kBThreads * VEC);
  const int64_t lo = (int64_t)blockIdx.x * seg;
  int64_t hi = lo + seg;
  if (hi > n) hi = n;  // (VEC = 2: n is even, so is hi)
  for (int j0 = 0; j0 < k; j0 += kJT) {
    const int jn = (k - j0 < kJT) ? (k - j0) : kJT;
    double acc[kJT];
#pragma unroll
    for (int t = 0; t < kJT; ++t) acc[t] = 0.0;
    for (int64_t i = lo + (int64_t)threadIdx.x * VEC; i < hi; i += kBThreads * VEC) {
      RowVec<VEC> wi;
      wi.load(w + i);
      if (jn == kJT) {
        RowVec<VEC> v[kJT];
#pragma unroll
        for (int t = 0; t < kJT; ++t) v[t].load_stream(V + (int64_t)(j0 + t) * ld + i);
#pragma unroll
        for (int t = 0; t < kJT; ++t)
#pragma unroll
          for (int r = 0; r < VEC; ++r) acc[t] = fma(v[t].v[r], wi.v[r], acc[t]);
      } else {
        for (int t = 0; t < jn; ++t) {
          RowVec<VEC> v;
          v.load_stream(V + (int64_t)(j0 + t) * ld + i);
#pragma unroll
          for (int r = 0; r < VEC; ++r) acc[t] = fma(v.v[r], wi.v[r], acc[t]);
        }
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
// k <= kmax columns of coefficients are staged in shared memory.  Every row is accumulated in the same order whatever
// VEC is (four interleaved partial sums over the columns, then (a0 + a1) + (a2 + a3)), so the result does not depend on it.
template <int VEC>
__global__ void __launch_bounds__(kBThreads)
basis_combine_kernel(int64_t n, int k, const double* __restrict__ V, int64_t ld, const double* __restrict__ h,
                     const double* __restrict__ w, double sign, double* __restrict__ out, double* __restrict__ partials) {
  extern __shared__ double hs[];
  __shared__ double red[64];
  for (int j = threadIdx.x; j < k; j += kBThreads) hs[j] = h[j];
  __syncthreads();
  dd_t nrm = dd_zero();
  for (int64_t i = ((int64_t)blockIdx.x * kBThreads + threadIdx.x) * VEC; i < n; i += (int64_t)gridDim.x * kBThreads * VEC) {
    double a[4][VEC];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int r = 0; r < VEC; ++r) a[t][r] = 0.0;
    int j = 0;
    for (; j + 8 <= k; j += 8) {
      RowVec<VEC> v[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t].load_stream(V + (int64_t)(j + t) * ld + i);
#pragma unroll
      for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int r = 0; r < VEC; ++r) a[t & 3][r] = fma(v[t].v[r], hs[j + t], a[t & 3][r]);
    }
    for (; j + 4 <= k; j += 4) {
      RowVec<VEC> v[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) v[t].load_stream(V + (int64_t)(j + t) * ld + i);
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int r = 0; r < VEC; ++r) a[t][r] = fma(v[t].v[r], hs[j + t], a[t][r]);
    }
    for (; j < k; ++j) {
      RowVec<VEC> v;
      v.load_stream(V + (int64_t)j * ld + i);
#pragma unroll
      for (int r = 0; r < VEC; ++r) a[0][r] = fma(v.v[r], hs[j], a[0][r]);
    }
    RowVec<VEC> wv, res;
    if (w != nullptr) wv.load(w + i);
#pragma unroll
    for (int r = 0; r < VEC; ++r) {
      const double vh = (a[0][r] + a[1][r]) + (a[2][r] + a[3][r]);
      double x = (sign < 0.0) ? -vh : vh;
      if (w != nullptr) x = (sign < 0.0) ? __dsub_rn(wv.v[r], vh) : __dadd_rn(wv.v[r], vh);
      res.v[r] = x;
      if (partials != nullptr) nrm = dd_fma(nrm, x, x);
    }
    if (VEC == 2)
      *reinterpret_cast<double2*>(out + i) = make_double2(res.v[0], res.v[VEC - 1]);
    else
      out[i] = res.v[0];
  }
  if (partials != nullptr) {
    const dd_t tot = dd_block_sum(nrm, red);
    if (threadIdx.x == 0) {
      partials[2 * blockIdx.x] = tot.hi;
      partials[2 * blockIdx.x + 1] = tot.lo;
    }
  }
}

static bool g_basis_vec2 = true;
// two rows per access need even n and ld and 16-byte aligned bases
static inline bool vec2_ok(int64_t n, int64_t ld, const void* a, const void* b, const void* c) {
  return g_basis_vec2 && n % 2 == 0 && ld % 2 == 0 && (uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0 && (uintptr_t)c % 16 == 0;
}

}  // namespace tb200

using namespace tb200;

extern "C" {

// A/B switch (tests, tuning): 1 (default) = two rows per 16-byte access where the alignment allows, 0 = one row.
void tb200_basis_set_vec2(int on) { g_basis_vec2 = on != 0; }

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
  if (vec2_ok(n, ld, V, w, nullptr))
    basis_dots_kernel<2><<<g, kBThreads, 0, st>>>(n, (int)k, V, ld, w, ws);
  else
    basis_dots_kernel<1><<<g, kBThreads, 0, st>>>(n, (int)k, V, ld, w, ws);
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
  if (vec2_ok(n, ld, V, w, out))
    basis_combine_kernel<2><<<g, kBThreads, (size_t)k * sizeof(double), st>>>(n, (int)k, V, ld, h, w, sign, out,
                                                                              norm_out ? ws : nullptr);
  else
    basis_combine_kernel<1><<<g, kBThreads, (size_t)k * sizeof(double), st>>>(n, (int)k, V, ld, h, w, sign, out,
                                                                              norm_out ? ws : nullptr);
  int rc = check_launch("basis_combine");
  if (rc || !norm_out) return rc;
  finalize_dd_kernel<<<1, 1024, 0, st>>>(ws, g, norm_out);
  return check_launch("basis_combine finalize");
}

}  // extern "C"
