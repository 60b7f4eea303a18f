// Weighted Gram matrix of a tall-skinny basis in one streaming pass, and the k x k factorisation that turns it
// into the triangular factors GKS / MMGKS consume.
//
// Reference: every GKS / MMGKS iteration does a Householder QR of the m x k and p x k matrices on the host,
//   (Q_A, R_A) = la.qr(AV * wf), (Q_L, R_L) = la.qr(LV * wr), Q_A.T @ b        trips/solvers/MMGKS.py:57-59,94-106
//   (Q_A, R_A) = la.qr(AV), _, R_L = la.qr(LV)                                 trips/solvers/GKS.py:54-58
// and only R_A, R_L, Q_A^T b, Q_A^T (wf*b) and ||z - Q_A Q_A^T z|| are used downstream.
// All of those are functions of the Gram matrix of [W*B | extra columns]:  R = chol(G_BB), Q^T z = R^-T G_Bz,
// ||z - QQ^T z||^2 = G_zz - ||Q^T z||^2.  Forming G in working precision would square the condition number
// (error ~ eps*kappa^2), so the pass accumulates every entry in double-double (TwoProduct via FMA + TwoSum,
// "Dot2"), the partials are combined in double-double in a fixed order, and the k x k Cholesky / triangular
// solves run in double-double on the host.  The factors are then accurate to working precision for
// kappa up to ~1e15, i.e. as good as Householder, while the basis is read exactly once (8*m*k bytes).
// The accumulation is ~10 flops per multiply-add: this pass is FP64-pipe-bound for k >~ 24, not HBM-bound.
#include "tb200_common.cuh"
#include <cmath>
#include <vector>

namespace tb200 {

constexpr int kGThreads = 256;
constexpr int kGMaxExtra = 4;
constexpr int kGBlocksX = 296;  // 2 CTAs per SM; fixed => deterministic reduction tree

struct GramExtras {
  const double* ptr[kGMaxExtra];
  int weighted[kGMaxExtra];
  int n;
};

struct dd {
  double hi, lo;
};
__host__ __device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
  s = a + b;
  const double z = s - a;
  e = (a - (s - z)) + (b - z);
}
__device__ __forceinline__ void dd_mac(double a, double b, double& hi, double& lo) {
  const double p = __dmul_rn(a, b);
  const double e = __fma_rn(a, b, -p);
  const double s = __dadd_rn(hi, p);
  const double z = __dsub_rn(s, hi);
  const double err = __dadd_rn(__dsub_rn(hi, __dsub_rn(s, z)), __dsub_rn(p, z));
  hi = s;
  lo = __dadd_rn(lo, __dadd_rn(err, e));
}

// pair index -> (ti, tj) with ti <= tj over nt tiles, row-major upper triangle
__host__ __device__ __forceinline__ void pair_to_tiles(int p, int nt, int& ti, int& tj) {
  int i = 0, rowlen = nt;
  while (p >= rowlen) {
    p -= rowlen;
    ++i;
    --rowlen;
  }
  ti = i;
  tj = i + p;
}

// A CTA streams tiles of `rows` consecutive rows through shared memory; thread (pair, rg) accumulates the 4 x 4 block
// `pair` of the Gram matrix over the rows of group rg of every tile (rows / RG rows per tile).  For the small K of the
// Krylov solvers (k <= ~20) a handful of 4 x 4 blocks exists, so the rows of a tile are split over up to 64 groups to
// keep all 256 threads busy (round 1 used 32-row tiles and left 90 % of the threads idle: 617 us for 1M x 8, now tens).
// partials layout: [blockIdx.x][pair][rg][16 entries][hi, lo]
__global__ void __launch_bounds__(kGThreads, 2)
gram_dd_kernel(int64_t m, int k, const double* __restrict__ B, int64_t ld, const double* __restrict__ w, GramExtras ex,
               int K, int Kpad, int npairs, int RG, int rows, double* __restrict__ partials) {
  extern __shared__ double T[];  // rows x Kpad
  const int nt = (K + 3) / 4;
  const int slot = blockIdx.y * kGThreads + threadIdx.x;  // (pair, rg) assignment
  const bool active = slot < npairs * RG;
  int ti = 0, tj = 0, rg = 0;
  if (active) {
    pair_to_tiles(slot % npairs, nt, ti, tj);
    rg = slot / npairs;
  }
  double hi[16], lo[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) hi[q] = lo[q] = 0.0;
  const int rows_per_group = rows / RG;
  const int64_t nblk = (m + rows - 1) / rows;

  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t row0 = blk * rows;
    __syncthreads();
    // stage the tile: lanes run along rows (contiguous in memory), warps over columns
    for (int r = threadIdx.x; r < rows; r += kGThreads) {  // (no index division: column loop outside, rows strided)
      const int64_t row = row0 + r;
      const bool in = row < m;
      const double wr = (in && w) ? w[row] : 1.0;
      double* Tr = T + r * Kpad;
      for (int j = 0; j < k; ++j) {
        double v = in ? B[(int64_t)j * ld + row] : 0.0;
        if (w) v = __dmul_rn(v, wr);
        Tr[j] = v;
      }
      for (int j = k; j < K; ++j) {
        double v = in ? ex.ptr[j - k][row] : 0.0;
        if (w && ex.weighted[j - k]) v = __dmul_rn(v, wr);
        Tr[j] = v;
      }
      for (int j = K; j < Kpad; ++j) Tr[j] = 0.0;
    }
    __syncthreads();
    if (active) {
      const double* Ta = T + 4 * ti;
      const double* Tb = T + 4 * tj;
      for (int r = rg * rows_per_group; r < (rg + 1) * rows_per_group; ++r) {
        double a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a[q] = Ta[r * Kpad + q];
          b[q] = Tb[r * Kpad + q];
        }
#pragma unroll
        for (int qi = 0; qi < 4; ++qi)
#pragma unroll
          for (int qj = 0; qj < 4; ++qj) dd_mac(a[qi], b[qj], hi[qi * 4 + qj], lo[qi * 4 + qj]);
      }
    }
  }
  if (active) {
    double* out = partials + ((((int64_t)blockIdx.x * npairs + (slot % npairs)) * RG + rg) * 16) * 2;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      out[2 * q] = hi[q];
      out[2 * q + 1] = lo[q];
    }
  }
}

// One CTA per (pair, entry): double-double sum over CTAs and row groups - every thread a fixed strided subset in order,
// then a fixed tree; writes both triangles.
__global__ void __launch_bounds__(128)
gram_finalize_kernel(int nbx, int npairs, int RG, int nt, int K, const double* __restrict__ partials,
                     double* __restrict__ Ghi, double* __restrict__ Glo) {
  __shared__ double sh[128], sl[128];
  const int id = blockIdx.x;
  const int pair = id / 16, q = id % 16;
  int ti, tj;
  pair_to_tiles(pair, nt, ti, tj);
  const int i = 4 * ti + q / 4, j = 4 * tj + q % 4;
  if (i >= K || j >= K) return;  // (uniform over the CTA)
  double shi = 0.0, slo = 0.0;
  const int terms = nbx * RG;
  for (int t = threadIdx.x; t < terms; t += 128) {
    const int bx = t / RG, rg = t % RG;
    const double* p = partials + ((((int64_t)bx * npairs + pair) * RG + rg) * 16 + q) * 2;
    double s, e;
    two_sum(shi, p[0], s, e);
    slo += e + p[1];
    shi = s;
  }
  sh[threadIdx.x] = shi, sl[threadIdx.x] = slo;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      double s, e;
      two_sum(sh[threadIdx.x], sh[threadIdx.x + o], s, e);
      sl[threadIdx.x] = sl[threadIdx.x] + sl[threadIdx.x + o] + e;
      sh[threadIdx.x] = s;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double s, e;
    two_sum(sh[0], sl[0], s, e);
    Ghi[i * K + j] = s;
    Glo[i * K + j] = e;
    Ghi[j * K + i] = s;
    Glo[j * K + i] = e;
  }
}

static void gram_shape(int K, int& Kpad, int& npairs, int& RG, int& gy, int& rows) {
  const int nt = (K + 3) / 4;
  Kpad = 4 * nt + 1;  // odd stride: conflict-free column staging
  npairs = nt * (nt + 1) / 2;
  // as many row groups as fit one CTA next to the 4 x 4 blocks (K <= 4: one block, 256 groups), so that (nearly) every
  // thread accumulates; several CTAs in y only when there are more blocks than threads
  RG = npairs <= kGThreads ? kGThreads / npairs : 1;
  gy = (npairs * RG + kGThreads - 1) / kGThreads;
  // rows per group and tile: ~256 rows per tile (two barriers and one staging pass per tile), within 64 KB of shared
  // memory so that two CTAs stay resident per SM
  int rpg = (256 + RG - 1) / RG;
  if (rpg < 4) rpg = 4;
  while (rpg > 1 && (size_t)RG * rpg * Kpad * sizeof(double) > 64 * 1024) --rpg;
  rows = RG * rpg;
}

// ---- host double-double arithmetic for the k x k factorisation ------------------------------------------
struct hdd {
  double hi, lo;
};
static inline hdd h_norm(double s, double e) {
  double t = s + e;
  return {t, e - (t - s)};
}
static inline hdd h_add(hdd a, hdd b) {
  double s, e;
  two_sum(a.hi, b.hi, s, e);
  e += a.lo + b.lo;
  return h_norm(s, e);
}
static inline hdd h_neg(hdd a) { return {-a.hi, -a.lo}; }
static inline hdd h_mul(hdd a, hdd b) {
  const double p = a.hi * b.hi;
  double e = std::fma(a.hi, b.hi, -p);
  e += a.hi * b.lo + a.lo * b.hi;
  return h_norm(p, e);
}
static inline hdd h_div(hdd a, hdd b) {
  const double q1 = a.hi / b.hi;
  hdd r = h_add(a, h_neg(h_mul(b, {q1, 0.0})));
  const double q2 = r.hi / b.hi;
  r = h_add(r, h_neg(h_mul(b, {q2, 0.0})));
  const double q3 = r.hi / b.hi;
  hdd q = h_norm(q1, q2);
  return h_add(q, {q3, 0.0});
}
static inline hdd h_sqrt(hdd a) {
  if (a.hi <= 0.0) return {0.0, 0.0};
  const double x = std::sqrt(a.hi);
  // one Newton step in double-double: x + (a - x^2) / (2x)
  hdd x2 = h_mul({x, 0.0}, {x, 0.0});
  hdd d = h_add(a, h_neg(x2));
  return h_add({x, 0.0}, {d.hi / (2.0 * x), 0.0});
}

}  // namespace tb200

using namespace tb200;

extern "C" {

// Workspace (doubles) for a Gram pass over K = k + n_extra columns.
int64_t tb200_gram_workspace_len(int64_t K) {
  int Kpad, npairs, RG, gy, rows;
  gram_shape((int)K, Kpad, npairs, RG, gy, rows);
  return (int64_t)kGBlocksX * npairs * RG * 32;
}

// G = M^T M for M = [diag(w) B[:, 0..k) | extras], K = k + n_extra, accumulated in double-double.
// B column j is contiguous at B + j*ld; w (length m) may be NULL; extras: up to 4 column pointers, each with a
// flag saying whether the row weights apply to it.  Ghi/Glo: K x K row-major device outputs (G = Ghi + Glo).
int tb200_weighted_gram(int64_t m, int64_t k, const double* B, int64_t ld, const double* w, int n_extra,
                        const double* const* extras, const int* extra_weighted, double* Ghi, double* Glo, double* ws,
                        void* stream) {
  TB200_REQUIRE(m >= 0 && k >= 0 && ld >= m && n_extra >= 0 && n_extra <= kGMaxExtra, "bad size");
  const int K = (int)k + n_extra;
  TB200_REQUIRE(K >= 1 && K <= 512, "need 1 <= k + n_extra <= 512");
  TB200_REQUIRE((k == 0 || B) && Ghi && Glo && ws, "null pointer");
  GramExtras ex;
  ex.n = n_extra;
  for (int i = 0; i < kGMaxExtra; ++i) {
    ex.ptr[i] = (i < n_extra) ? extras[i] : nullptr;
    ex.weighted[i] = (i < n_extra && extra_weighted) ? extra_weighted[i] : 0;
    TB200_REQUIRE(i >= n_extra || ex.ptr[i], "null extra column");
  }
  int Kpad, npairs, RG, gy, rows;
  gram_shape(K, Kpad, npairs, RG, gy, rows);
  const size_t smem = (size_t)rows * Kpad * sizeof(double);
  cudaStream_t st = (cudaStream_t)stream;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(gram_dd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("weighted_gram: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  dim3 grid(kGBlocksX, gy);
  gram_dd_kernel<<<grid, kGThreads, smem, st>>>(m, (int)k, B, ld, w, ex, K, Kpad, npairs, RG, rows, ws);
  int rc = check_launch("weighted_gram");
  if (rc) return rc;
  gram_finalize_kernel<<<npairs * 16, 128, 0, st>>>(kGBlocksX, npairs, RG, (K + 3) / 4, K, ws, Ghi, Glo);
  return check_launch("weighted_gram finalize");
}

// Host-side, double-double: from the K x K Gram matrix of [B | z_1..z_ne] (K = k + ne, row-major hi/lo parts)
// compute   R (k x k upper triangular, row-major, B = Q R with positive diagonal),
//           C (k x ne, row-major) = Q^T z_e,    resid2[e] = ||z_e - Q Q^T z_e||^2.
// Returns 0, or j+1 if the leading minor of order j+1 is not positive definite.
int tb200_gram_factor_dd(int k, int ne, const double* Ghi, const double* Glo, double* R, double* C, double* resid2) {
  const int K = k + ne;
  std::vector<hdd> Rm((size_t)k * k, hdd{0.0, 0.0});
  auto G = [&](int i, int j) { return hdd{Ghi[(size_t)i * K + j], Glo[(size_t)i * K + j]}; };
  // upper Cholesky: G = R^T R, row by row
  for (int i = 0; i < k; ++i) {
    hdd d = G(i, i);
    for (int p = 0; p < i; ++p) d = h_add(d, h_neg(h_mul(Rm[(size_t)p * k + i], Rm[(size_t)p * k + i])));
    if (!(d.hi > 0.0)) return i + 1;
    const hdd rii = h_sqrt(d);
    Rm[(size_t)i * k + i] = rii;
    for (int j = i + 1; j < k; ++j) {
      hdd s = G(i, j);
      for (int p = 0; p < i; ++p) s = h_add(s, h_neg(h_mul(Rm[(size_t)p * k + i], Rm[(size_t)p * k + j])));
      Rm[(size_t)i * k + j] = h_div(s, rii);
    }
  }
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) R[(size_t)i * k + j] = (j >= i) ? Rm[(size_t)i * k + j].hi : 0.0;
  // C = R^-T G_Bz (forward substitution), residuals
  for (int e = 0; e < ne; ++e) {
    std::vector<hdd> c(k);
    hdd cc{0.0, 0.0};
    for (int i = 0; i < k; ++i) {
      hdd s = G(i, k + e);
      for (int p = 0; p < i; ++p) s = h_add(s, h_neg(h_mul(Rm[(size_t)p * k + i], c[p])));
      c[i] = h_div(s, Rm[(size_t)i * k + i]);
      cc = h_add(cc, h_mul(c[i], c[i]));
      if (C) C[(size_t)i * ne + e] = c[i].hi;
    }
    if (resid2) {
      const hdd r = h_add(G(k + e, k + e), h_neg(cc));
      resid2[e] = r.hi > 0.0 ? r.hi : 0.0;
    }
  }
  return 0;
}

}  // extern "C"
