// CSR sparse matrix-vector product for the tomography operator and its explicitly stored transpose.
//
// Replaces the reference's `A @ v` / `A.T @ u` call sites (scipy sparsetools csr_matvec / csc_matvec behind
// trips/utilities/decompositions.py:235-240, trips/solvers/CGLS.py:45-68, GKS.py:82-92, MMGKS.py:43-124).
// A.T is never applied by scatter: the caller stores A^T as a second CSR matrix and calls the same kernel,
// so both directions are gather-only and bitwise reproducible (no atomics).
//
// Kernel shape (HBM-bound, 12 B/nnz streamed once):
//  * CT rows are long (~900-1800 nnz): one warp per row, lanes take groups of four consecutive non-zeros;
//    values arrive as one 256-bit LDG (evict-first in L2, not allocated in L1), column indices as one 128-bit
//    LDG, and the four x[col] gathers go through the read-only path with an evict-last L2 policy so the
//    dense vector (33.5 MB at 2048^2) stays L2/L1 resident while 46 GB of matrix streams past it.
//  * the row start is peeled to a multiple of four non-zeros so the wide loads are aligned for any rowptr.
//  * warps of one CTA work on neighbouring rows (neighbouring rays / pixels) at the same time, which is what
//    makes the x gathers hit L1: adjacent rays cross adjacent pixels.
//  * fused epilogue  y = A x - coef * z  (the Golub-Kahan three-term recurrences) and fused ||y||^2:
//    one partial per CTA, then a single-CTA fixed-order finalize => deterministic norms without atomics.
//  * short-row matrices (generic CSR passed by a user) take a sub-warp path with T threads per row.
#include "tb200_common.cuh"

namespace tb200 {

template <typename VT>
struct StreamVals;

template <>
struct StreamVals<double> {
  __device__ __forceinline__ static void load4(const double* p, uint64_t, double (&v)[4]) { ld_stream_f64x4(p, v); }
};
template <>
struct StreamVals<float> {
  __device__ __forceinline__ static void load4(const float* p, uint64_t pol, double (&v)[4]) {
    float f[4];
    ld_stream_f32x4(p, pol, f);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (double)f[i];
  }
};

// One warp per row; WARPS warps per CTA; each CTA owns `rows_per_cta` consecutive rows, visited so that at any
// time the CTA's warps sit on WARPS adjacent rows.
template <typename VT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
spmv_warp_kernel(int64_t m, int rows_per_cta, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                 const VT* __restrict__ val, const double* __restrict__ x, double* __restrict__ y, double coef_host,
                 const double* __restrict__ coef_dev, const double* __restrict__ z, double* __restrict__ partials) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t pol_stream = policy_evict_first();
  const uint64_t pol_keep = policy_evict_last();
  const double coef = (z != nullptr) ? (coef_dev ? *coef_dev : coef_host) : 0.0;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  double nrm = 0.0;

  for (int r = warp; r < rows_per_cta; r += WARPS) {
    const int64_t row = row0 + r;
    if (row >= m) break;
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    int64_t a = (s + 3) & ~(int64_t)3;
    if (a > e) a = e;
    double acc0 = 0.0, acc1 = 0.0;
    // head: up to three unaligned leading entries
    if (s + lane < a) acc0 = (double)val[s + lane] * ld_gather_f64(x + col[s + lane], pol_keep);
    const int64_t ngrp = (e - a) >> 2;
    int64_t g = lane;
    for (; g + 32 < ngrp; g += 64) {
      int32_t c0[4], c1[4];
      double v0[4], v1[4];
      ld_stream_i32x4(col + a + 4 * g, pol_stream, c0);
      ld_stream_i32x4(col + a + 4 * (g + 32), pol_stream, c1);
      StreamVals<VT>::load4(val + a + 4 * g, pol_stream, v0);
      StreamVals<VT>::load4(val + a + 4 * (g + 32), pol_stream, v1);
      double x0[4], x1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) x0[i] = ld_gather_f64(x + c0[i], pol_keep);
#pragma unroll
      for (int i = 0; i < 4; ++i) x1[i] = ld_gather_f64(x + c1[i], pol_keep);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc0 = fma(v0[i], x0[i], acc0);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc1 = fma(v1[i], x1[i], acc1);
    }
    if (g < ngrp) {
      int32_t c0[4];
      double v0[4];
      ld_stream_i32x4(col + a + 4 * g, pol_stream, c0);
      StreamVals<VT>::load4(val + a + 4 * g, pol_stream, v0);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc0 = fma(v0[i], ld_gather_f64(x + c0[i], pol_keep), acc0);
    }
    // tail: up to three trailing entries
    const int64_t t0 = a + 4 * ngrp;
    if (t0 + lane < e) acc1 = fma((double)val[t0 + lane], ld_gather_f64(x + col[t0 + lane], pol_keep), acc1);

    double sum = warp_sum(acc0 + acc1);
    if (lane == 0) {
      if (z != nullptr) sum = __dsub_rn(sum, __dmul_rn(coef, z[row]));
      y[row] = sum;
      nrm = fma(sum, sum, nrm);
    }
  }
  if (partials != nullptr) {
    const double tot = block_sum(nrm, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
  }
}

// T threads per row (T in {2,4,8,16}); scalar loads. For generic short-row CSR matrices.
template <typename VT, int T>
__global__ void __launch_bounds__(256)
spmv_subwarp_kernel(int64_t m, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    const VT* __restrict__ val, const double* __restrict__ x, double* __restrict__ y, double coef_host,
                    const double* __restrict__ coef_dev, const double* __restrict__ z, double* __restrict__ partials) {
  __shared__ double red[32];
  const int sub = threadIdx.x % T;
  const int64_t row = ((int64_t)blockIdx.x * 256 + threadIdx.x) / T;
  const double coef = (z != nullptr) ? (coef_dev ? *coef_dev : coef_host) : 0.0;
  double acc = 0.0;
  if (row < m) {
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    for (int64_t i = s + sub; i < e; i += T) acc = fma((double)val[i], x[col[i]], acc);
  }
#pragma unroll
  for (int o = T / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  double nrm = 0.0;
  if (row < m && sub == 0) {
    if (z != nullptr) acc = __dsub_rn(acc, __dmul_rn(coef, z[row]));
    y[row] = acc;
    nrm = acc * acc;
  }
  if (partials != nullptr) {
    const double tot = block_sum(nrm, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
  }
}

struct SpmvPlan {
  int threads_per_row;  // 32 => warp kernel
  int rows_per_cta;
  int64_t nblocks;
};

static SpmvPlan make_plan(int64_t m, int64_t nnz) {
  SpmvPlan p;
  const double avg = (m > 0) ? (double)nnz / (double)m : 0.0;
  if (avg >= 48.0) {
    p.threads_per_row = 32;
    p.rows_per_cta = 32;  // 8 warps x 4 rows
  } else {
    int t = 2;
    while (t < 16 && t < avg) t <<= 1;
    p.threads_per_row = t;
    p.rows_per_cta = 256 / t;
  }
  p.nblocks = (m + p.rows_per_cta - 1) / p.rows_per_cta;
  if (p.nblocks < 1) p.nblocks = 1;
  return p;
}

template <typename VT>
static int spmv_launch(int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* col, const VT* val,
                       const double* x, double* y, double coef_host, const double* coef_dev, const double* z,
                       double* norm_out, double* ws, cudaStream_t st) {
  (void)n;
  const SpmvPlan p = make_plan(m, nnz);
  double* partials = norm_out ? ws : nullptr;
  if (p.threads_per_row == 32) {
    spmv_warp_kernel<VT, 8><<<(unsigned)p.nblocks, 256, 0, st>>>(m, p.rows_per_cta, rowptr, col, val, x, y, coef_host,
                                                                 coef_dev, z, partials);
  } else {
    switch (p.threads_per_row) {
      case 2:
        spmv_subwarp_kernel<VT, 2><<<(unsigned)p.nblocks, 256, 0, st>>>(m, rowptr, col, val, x, y, coef_host, coef_dev, z, partials);
        break;
      case 4:
        spmv_subwarp_kernel<VT, 4><<<(unsigned)p.nblocks, 256, 0, st>>>(m, rowptr, col, val, x, y, coef_host, coef_dev, z, partials);
        break;
      case 8:
        spmv_subwarp_kernel<VT, 8><<<(unsigned)p.nblocks, 256, 0, st>>>(m, rowptr, col, val, x, y, coef_host, coef_dev, z, partials);
        break;
      default:
        spmv_subwarp_kernel<VT, 16><<<(unsigned)p.nblocks, 256, 0, st>>>(m, rowptr, col, val, x, y, coef_host, coef_dev, z, partials);
        break;
    }
  }
  int rc = check_launch("spmv");
  if (rc) return rc;
  if (norm_out) {
    finalize_sum_kernel<<<1, 1024, 0, st>>>(ws, p.nblocks, norm_out);
    rc = check_launch("spmv finalize");
  }
  return rc;
}

}  // namespace tb200

using namespace tb200;

extern "C" {

// Number of doubles of workspace a fused-norm SpMV over m rows may need (upper bound over all plans).
int64_t tb200_spmv_workspace_len(int64_t m) { return (m + 7) / 8 + 8; }

// How many of this library's kernels one call enqueues (for launch accounting in bench.py).
int tb200_spmv_launches(int with_norm) { return with_norm ? 2 : 1; }

static int check_spmv_args(int64_t m, int64_t n, int64_t nnz, const void* rowptr, const void* col, const void* val,
                           const void* x, const void* y, const void* norm_out, const void* ws) {
  TB200_REQUIRE(m >= 0 && n >= 0 && nnz >= 0, "negative size");
  TB200_REQUIRE(n < ((int64_t)1 << 31), "n must fit int32 column indices");
  TB200_REQUIRE(rowptr && x && y, "null pointer");
  TB200_REQUIRE(nnz == 0 || (col && val), "null matrix arrays");
  TB200_REQUIRE(((uintptr_t)val % 32) == 0 && ((uintptr_t)col % 16) == 0, "vals must be 32-byte and colidx 16-byte aligned");
  TB200_REQUIRE(norm_out == nullptr || ws != nullptr, "norm_out requires a workspace");
  return 0;
}

int tb200_spmv_csr_f64(int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* colidx,
                       const double* vals, const double* x, double* y, double coef_host, const double* coef_dev,
                       const double* z, double* norm_out, double* ws, void* stream) {
  int rc = check_spmv_args(m, n, nnz, rowptr, colidx, vals, x, y, norm_out, ws);
  if (rc) return rc;
  if (m == 0) return 0;
  return spmv_launch<double>(m, n, nnz, rowptr, colidx, vals, x, y, coef_host, coef_dev, z, norm_out, ws,
                             (cudaStream_t)stream);
}

int tb200_spmv_csr_f32s(int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* colidx,
                        const float* vals, const double* x, double* y, double coef_host, const double* coef_dev,
                        const double* z, double* norm_out, double* ws, void* stream) {
  int rc = check_spmv_args(m, n, nnz, rowptr, colidx, vals, x, y, norm_out, ws);
  if (rc) return rc;
  if (m == 0) return 0;
  return spmv_launch<float>(m, n, nnz, rowptr, colidx, vals, x, y, coef_host, coef_dev, z, norm_out, ws,
                            (cudaStream_t)stream);
}

// out[0] = sum(partials[0..n)), out[1] = sqrt(out[0]); fixed order, one CTA.
int tb200_reduce_finalize(const double* partials, int64_t n, double* out, void* stream) {
  TB200_REQUIRE(partials && out && n >= 0, "bad argument");
  finalize_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, n, out);
  return check_launch("reduce_finalize");
}

}  // extern "C"
