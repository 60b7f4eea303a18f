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
// The accumulation is 10 fp64 instructions per multiply-add: this pass is bound by the fp64 issue rate, not by HBM
// (DESIGN.md section 4b has the measurements).
#include "tb200_common.cuh"
#include <cmath>
#include <vector>

namespace tb200 {

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

// ---- mbarrier / bulk-copy (TMA, 1-D) primitives ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy (bytes a multiple of 16, both addresses 16-byte aligned), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// A CTA streams tiles of `rows` = RG * rpg consecutive rows through shared memory, column-major (T[column][row], the
// layout the basis has in HBM, so a tile is one contiguous segment per column).  Thread (item, rg) accumulates the
// BH x BW block `item` of the Gram matrix over rows rg, rg + RG, rg + 2 RG, ... of every tile: the lanes of a warp read
// consecutive rows of the same columns (broadcast / conflict-free) as long as they share the item.  For the small K of
// the Krylov solvers only a handful of blocks exists, so a tile is split over up to 512 row groups to keep all threads
// busy; their partial sums are combined inside the CTA before anything is written.
//
// BULK = true (every column 16-byte aligned): full tiles arrive by cp.async.bulk (one per column, issued by warp 0,
// completion on an mbarrier) into one of two buffers, so tile t + 1 is in flight while tile t is accumulated - the
// round-2 profile of the single-buffered kernel showed the fp64 pipe idle 70 % of the time waiting on the staging
// loads.  The row weights travel as one more column and are applied in shared memory before the accumulation.
// BULK = false and the last, partial tile: staged by ordinary loads (zero-filled past m).
//
// BH = BW (2 or 4): an item is a square block (ti, tj), ti <= tj, of the full Gram matrix; 2 x 2 on up to 512 threads is
// the default (gram_launch).  BH = 4, BW = 1 ("panel"): only the columns c >= c0 of G are wanted (a basis that gained
// columns since its Gram matrix was last formed, unweighted: everything left of c0 is unchanged) - an item is the 4 x 1
// block (rows 4 ti .. 4 ti + 3, column c0 + pj), O(K) items instead of O(K^2), and the pass is HBM-bound (it still
// reads every column once).
// NT: upper bound of the CTA size (launch bounds); the launch uses the multiple of 32 that fits the items.
// partials layout: [blockIdx.x][item][BH * BW entries][hi, lo]
template <bool BULK, int BH, int BW, int NT>
__global__ void __launch_bounds__(NT, 2)
gram_dd_kernel(int64_t m, int k, const double* __restrict__ B, int64_t ld, const double* __restrict__ w,
               const __grid_constant__ GramExtras ex, int K, int c0, int npairs, int RG, int rpg, int S, int KC,
               double* __restrict__ partials) {
  constexpr int NE = BH * BW;
  const int kGThreads = blockDim.x;  // (shadows the default) this launch's CTA size: a multiple of 32, <= NT
  extern __shared__ __align__(16) double T[];  // [2][KC][rows]  (one buffer when !BULK)
  __shared__ uint64_t bars[2];
  const int nt = (K + BH - 1) / BH;  // row tiles (= column tiles of the square blocks)
  const int rows = RG * rpg;  // S >= rows: column stride in shared memory (see gram_shape)
  const size_t bufsz = (size_t)KC * S;
  const int Kw = K + (w ? 1 : 0);
  const int slot = blockIdx.y * kGThreads + threadIdx.x;  // (pair, rg) assignment
  const bool active = slot < npairs * RG;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int ti = 0, cb = 0, rg = 0;  // cb: first column of the item's b operand
  if (active) {
    if (BW == BH) {
      int tj;
      pair_to_tiles(slot / RG, nt, ti, tj);
      cb = BW * tj;
    } else {
      ti = (slot / RG) % nt;
      cb = c0 + (slot / RG) / nt;
    }
    rg = slot % RG;
  }
  double hi[NE], lo[NE];
#pragma unroll
  for (int q = 0; q < NE; ++q) hi[q] = lo[q] = 0.0;

  auto colptr = [&](int j) -> const double* { return j < k ? B + (int64_t)j * ld : (j < K ? ex.ptr[j - k] : w); };
  auto weighted = [&](int j) -> bool { return w && (j < k || ex.weighted[j - k]); };
  auto accumulate = [&](const double* Tb_) {
    if (!active) return;
    const double* pa = Tb_ + (size_t)(BH * ti) * S + rg;
    const double* pb = Tb_ + (size_t)cb * S + rg;
#pragma unroll 2
    for (int i = 0; i < rpg; ++i) {
      double a[BH], b[BW];
#pragma unroll
      for (int q = 0; q < BH; ++q) a[q] = pa[q * S];
#pragma unroll
      for (int q = 0; q < BW; ++q) b[q] = pb[q * S];
#pragma unroll
      for (int qi = 0; qi < BH; ++qi)
#pragma unroll
        for (int qj = 0; qj < BW; ++qj) dd_mac(a[qi], b[qj], hi[qi * BW + qj], lo[qi * BW + qj]);
      pa += RG;
      pb += RG;
    }
  };
  // the pad columns K .. BH nt - 1 of the last block are read (their products are discarded): keep them zero
  for (int j = Kw + warp; j < KC; j += kGThreads / 32)
    for (int r = lane; r < rows; r += 32) {
      T[(size_t)j * S + r] = 0.0;
      if (BULK) T[bufsz + (size_t)j * S + r] = 0.0;
    }

  const int64_t nfull = m / rows;
  if (BULK) {
    if (threadIdx.x == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int64_t blk, int buf) {  // warp 0: one bulk copy per column
      if (warp != 0) return;
      if (lane == 0) mbar_arrive_expect_tx(&bars[buf], (uint32_t)Kw * (uint32_t)rows * 8u);
      __syncwarp();
      double* dst = T + buf * bufsz;
      for (int j = lane; j < Kw; j += 32) bulk_g2s(dst + (size_t)j * S, colptr(j) + blk * rows, (uint32_t)rows * 8u, &bars[buf]);
    };
    if ((int64_t)blockIdx.x < nfull) issue(blockIdx.x, 0);
    int t = 0;
    for (int64_t blk = blockIdx.x; blk < nfull; blk += gridDim.x, ++t) {
      const int buf = t & 1;
      if (blk + gridDim.x < nfull) issue(blk + gridDim.x, buf ^ 1);  // (buf^1 was last read before the barrier ending t-1)
      mbar_wait(&bars[buf], (t >> 1) & 1);
      double* Tc = T + buf * bufsz;
      if (w) {
        const double* Tw = Tc + (size_t)K * S;
        for (int j = warp; j < K; j += kGThreads / 32)
          if (weighted(j))
            for (int r = lane; r < rows; r += 32) Tc[(size_t)j * S + r] = __dmul_rn(Tc[(size_t)j * S + r], Tw[r]);
        __syncthreads();
      }
      accumulate(Tc);
      if (w) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my generic writes before the next bulk write
      __syncthreads();
    }
  }
  // tiles staged by ordinary loads: all of them when !BULK, else the partial last tile (owned by the next CTA in turn)
  const int64_t nblk = (m + rows - 1) / rows;
  for (int64_t blk = BULK ? nfull + ((blockIdx.x - nfull % gridDim.x + gridDim.x) % gridDim.x) : (int64_t)blockIdx.x; blk < nblk;
       blk += gridDim.x) {
    const int64_t row0 = blk * rows;
    __syncthreads();
    for (int j = warp; j < K; j += kGThreads / 32) {
      const double* src = colptr(j);
      const bool wt = weighted(j);
      for (int r = lane; r < rows; r += 32) {
        const int64_t row = row0 + r;
        double v = row < m ? src[row] : 0.0;
        if (wt && row < m) v = __dmul_rn(v, w[row]);
        T[(size_t)j * S + r] = v;
      }
    }
    __syncthreads();
    accumulate(T);
  }
  if (RG == 1) {
    if (active) {
      double* out = partials + (((int64_t)blockIdx.x * npairs + slot) * NE) * 2;
#pragma unroll
      for (int q = 0; q < NE; ++q) {
        out[2 * q] = hi[q];
        out[2 * q + 1] = lo[q];
      }
    }
    return;
  }
  // RG > 1 (one CTA in y): the row groups of a block are summed inside the CTA, in double-double and in a fixed order
  // (C threads per entry, strided, then those C in order), so one partial per (CTA, entry) is left for the finalize
  // kernel - with up to 256 groups that kernel used to dominate the small-K launches.
  __syncthreads();
  double* red = T;  // [2 NE][256]: entry component major, slot minor
#pragma unroll
  for (int q = 0; q < NE; ++q) {
    red[(2 * q) * kGThreads + threadIdx.x] = hi[q];
    red[(2 * q + 1) * kGThreads + threadIdx.x] = lo[q];
  }
  __syncthreads();
  const int E = npairs * NE;
  const int C = E >= kGThreads ? 1 : kGThreads / E;
  double* red2 = T + 2 * NE * kGThreads;  // [E * C][2]
  double* outb = partials + (int64_t)blockIdx.x * npairs * NE * 2;
  for (int id = threadIdx.x; id < E * C; id += kGThreads) {
    const int e = id / C, c = id % C, pair = e / NE, q = e % NE;
    const double* ph = red + (2 * q) * kGThreads + pair * RG;
    const double* pl = ph + kGThreads;
    double shi = 0.0, slo = 0.0;
    for (int g = c; g < RG; g += C) {
      double s_, e_;
      two_sum(shi, ph[g], s_, e_);
      slo += e_ + pl[g];
      shi = s_;
    }
    if (C == 1) {
      outb[2 * e] = shi;
      outb[2 * e + 1] = slo;
    } else {
      red2[2 * id] = shi;
      red2[2 * id + 1] = slo;
    }
  }
  if (C > 1) {
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += kGThreads) {
      double shi = 0.0, slo = 0.0;
      for (int c = 0; c < C; ++c) {
        double s_, e_;
        two_sum(shi, red2[2 * (e * C + c)], s_, e_);
        slo += e_ + red2[2 * (e * C + c) + 1];
        shi = s_;
      }
      outb[2 * e] = shi;
      outb[2 * e + 1] = slo;
    }
  }
}

// One CTA per (pair, entry): double-double sum over the CTAs of the pass - every thread a fixed strided subset in order,
// then a fixed tree; writes both triangles.
__global__ void __launch_bounds__(128)
gram_finalize_kernel(int nbx, int npairs, int bs, int nt, int K, int c0, const double* __restrict__ partials,
                     double* __restrict__ Ghi, double* __restrict__ Glo) {
  __shared__ double sh[128], sl[128];
  const int id = blockIdx.x;
  const int NE = c0 < 0 ? bs * bs : 4;
  const int pair = id / NE, q = id % NE;
  int i, j;
  if (c0 < 0) {
    int ti, tj;
    pair_to_tiles(pair, nt, ti, tj);
    i = bs * ti + q / bs, j = bs * tj + q % bs;
  } else {  // panel: item = (ti, pj); the entries below the diagonal belong to the item of the mirrored position
    i = 4 * (pair % nt) + q, j = c0 + pair / nt;
    if (i > j) return;
  }
  if (i >= K || j >= K) return;  // (uniform over the CTA)
  double shi = 0.0, slo = 0.0;
  for (int bx = threadIdx.x; bx < nbx; bx += 128) {
    const double* p = partials + (((int64_t)bx * npairs + pair) * NE + q) * 2;
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

static inline size_t gram_reduce_bytes(int ne, int threads) { return (size_t)(2 * ne + 2) * threads * sizeof(double); }
constexpr size_t kGBufBytes = 54 * 1024;  // one staging buffer; two buffers per CTA, two CTAs per SM

// bs: side of the square blocks of the full pass (4 or 2); the panel pass uses 4 x 1 blocks.  threads: CTA size.
static void gram_shape(int K, bool has_w, int c0, int bs, int max_threads, int& KC, int& npairs, int& RG, int& gy, int& rpg,
                       int& S, int& threads) {
  if (c0 >= 0) bs = 4;
  const int nt = (K + bs - 1) / bs;
  const int Kw = K + (has_w ? 1 : 0);
  KC = bs * nt > Kw ? bs * nt : Kw;
  npairs = c0 < 0 ? nt * (nt + 1) / 2 : nt * (K - c0);  // items: bs x bs blocks, or 4 x 1 blocks of the panel
  // Threads: one per (item, row group).  The CTA size is what the items need, not a fixed 256 / 512: with 276 items a
  // 512-thread CTA would leave 46 % of its threads idle (measured: fp64 pipe 63 % busy at K = 46).  Among the row-group
  // counts that fit, take the one with the best product of (useful lanes / launched lanes) and (balance of the two
  // resident CTAs' warps over the four schedulers of an SM); ties go to the larger CTA.
  RG = 1, threads = 32;
  if (npairs > max_threads) {
    gy = (npairs + max_threads - 1) / max_threads;
    threads = ((npairs + gy - 1) / gy + 31) / 32 * 32;
  } else {
    gy = 1;
    double best = -1.0;
    for (int rg = 1; rg * npairs <= max_threads; ++rg) {
      const int t = (rg * npairs + 31) / 32 * 32;
      const int wps = 2 * t / 32;  // warps of two resident CTAs
      // (a larger CTA also hides more latency: measured at K = 17, 512 threads at 0.97 beat 320 at 0.98)
      const double score = (double)(rg * npairs) / t * (wps / 4.0) / ((wps + 3) / 4) * (0.85 + 0.15 * t / max_threads);
      if (score >= best) best = score, RG = rg, threads = t;
    }
  }
  // ~256 rows per tile (one barrier pair per tile), an even number of rows (16-byte columns), within one buffer
  const int step = (RG & 1) ? 2 : 1;
  rpg = (256 + RG - 1) / RG;
  if (rpg < 4) rpg = 4;
  rpg = (rpg + step - 1) / step * step;
  while (rpg > step && (size_t)(RG * rpg + 2) * KC * sizeof(double) > kGBufBytes) rpg -= step;
  // column stride: even (16-byte columns for the bulk copies) with S / 2 odd - the threads of a half-warp that hold
  // different blocks read the same row of different columns, and a stride that is a multiple of 4 doubles would put
  // every second (S = 8j: every) column on the same bank (measured 2x on the whole pass at K = 56)
  const int rows = RG * rpg;
  S = (rows / 2) % 2 == 1 ? rows : rows + 2;
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

static bool g_gram_bulk = true;
static int g_gram_block = 0;

extern "C" {

// A/B switch for tests and tuning: 0 = stage every tile with ordinary loads (single buffer), 1 = bulk copies (default).
void tb200_gram_set_bulk(int on) { g_gram_bulk = on != 0; }
// Block shape of the full pass: 0 (default) = chosen per K, 2 = 2 x 2 blocks (<= 512 threads), 4 = 4 x 4 blocks (<= 256).
void tb200_gram_set_block(int bs) { g_gram_block = (bs == 4 || bs == 2) ? bs : 0; }

// Workspace (doubles) for a Gram pass over K = k + n_extra columns.
int64_t tb200_gram_workspace_len(int64_t K) {
  const int64_t nt = (K + 3) / 4;
  return (int64_t)kGBlocksX * (nt * (nt + 1) / 2) * 32;  // one (hi, lo) per CTA and entry of every 4 x 4 block
}

}  // extern "C"

static int gram_launch(int64_t m, int64_t k, const double* B, int64_t ld, const double* w, int n_extra,
                       const double* const* extras, const int* extra_weighted, int c0, double* Ghi, double* Glo, double* ws,
                       void* stream) {
  TB200_REQUIRE(m >= 0 && k >= 0 && ld >= m && n_extra >= 0 && n_extra <= kGMaxExtra, "bad size");
  const int K = (int)k + n_extra;
  TB200_REQUIRE(K >= 1 && K <= 512, "need 1 <= k + n_extra <= 512");
  TB200_REQUIRE((k == 0 || B) && Ghi && Glo && ws, "null pointer");
  TB200_REQUIRE(c0 < K, "panel starts past the last column");
  const int nt = (K + 3) / 4;
  if (c0 <= 0 || K - c0 > 2 * (nt + 1)) c0 = -1;  // a wide panel: the full pass is a superset and fits the workspace
  GramExtras ex;
  ex.n = n_extra;
  for (int i = 0; i < kGMaxExtra; ++i) {
    ex.ptr[i] = (i < n_extra) ? extras[i] : nullptr;
    ex.weighted[i] = (i < n_extra && extra_weighted) ? extra_weighted[i] : 0;
    TB200_REQUIRE(i >= n_extra || ex.ptr[i], "null extra column");
  }
  // full pass: 2 x 2 blocks on up to 512 threads (finer items: fewer idle threads and padded entries; measured 10-40 %
  // faster than 4 x 4 blocks on 256 threads for K <= 56) - except where its items fill only ~300 threads of a CTA
  // (K = 45..48: 276 / 300 items, one row group), where four independent accumulations per thread on 9 warps do not keep
  // the pipe busy and the 4 x 4 shape (16 per thread) is 12 % faster.  K > 128: 4 x 4.  Panel: 4 x 1 blocks on 256.
  int bs = 4, max_threads = 256;
  int KC, npairs, RG, gy, rpg, S, threads;
  if (c0 < 0 && K <= 128 && g_gram_block != 4) {
    gram_shape(K, w != nullptr, c0, 2, 512, KC, npairs, RG, gy, rpg, S, threads);
    if (g_gram_block == 2 || threads >= 320) bs = 2, max_threads = 512;
  }
  gram_shape(K, w != nullptr, c0, bs, max_threads, KC, npairs, RG, gy, rpg, S, threads);
  TB200_REQUIRE((size_t)S * KC * sizeof(double) <= kGBufBytes, "k + n_extra too large for one shared-memory tile");
  // bulk copies need every column segment 16-byte aligned (rows is even): B, ld, the extras and w
  bool bulk = g_gram_bulk && ((uintptr_t)B % 16 == 0) && (ld % 2 == 0) && (!w || (uintptr_t)w % 16 == 0);
  for (int i = 0; i < n_extra; ++i) bulk = bulk && ((uintptr_t)ex.ptr[i] % 16 == 0);
  size_t smem = (size_t)S * KC * sizeof(double) * (bulk ? 2 : 1);
  const size_t red_bytes = gram_reduce_bytes(c0 < 0 ? bs * bs : 4, threads);
  if (RG > 1 && smem < red_bytes) smem = red_bytes;  // the in-CTA reduction over row groups reuses the tiles
  cudaStream_t st = (cudaStream_t)stream;
  auto kern = c0 >= 0 ? (bulk ? gram_dd_kernel<true, 4, 1, 256> : gram_dd_kernel<false, 4, 1, 256>)
              : bs == 2 ? (bulk ? gram_dd_kernel<true, 2, 2, 512> : gram_dd_kernel<false, 2, 2, 512>)
                        : (bulk ? gram_dd_kernel<true, 4, 4, 256> : gram_dd_kernel<false, 4, 4, 256>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("weighted_gram: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  dim3 grid(kGBlocksX, gy);
  kern<<<grid, threads, smem, st>>>(m, (int)k, B, ld, w, ex, K, c0, npairs, RG, rpg, S, KC, ws);
  int rc = check_launch("weighted_gram");
  if (rc) return rc;
  gram_finalize_kernel<<<npairs * (c0 < 0 ? bs * bs : 4), 128, 0, st>>>(kGBlocksX, npairs, bs, (K + bs - 1) / bs, K, c0, ws,
                                                                       Ghi, Glo);
  return check_launch("weighted_gram finalize");
}

extern "C" {

// G = M^T M for M = [diag(w) B[:, 0..k) | extras], K = k + n_extra, accumulated in double-double.
// B column j is contiguous at B + j*ld; w (length m) may be NULL; extras: up to 4 column pointers, each with a
// flag saying whether the row weights apply to it.  Ghi/Glo: K x K row-major device outputs (G = Ghi + Glo).
int tb200_weighted_gram(int64_t m, int64_t k, const double* B, int64_t ld, const double* w, int n_extra,
                        const double* const* extras, const int* extra_weighted, double* Ghi, double* Glo, double* ws,
                        void* stream) {
  return gram_launch(m, k, B, ld, w, n_extra, extras, extra_weighted, -1, Ghi, Glo, ws, stream);
}

// The same, but only the columns c >= c0 of G (entries (i, c) and (c, i), i <= c) are computed and written; the rest of
// Ghi/Glo is left untouched.  For a basis whose first c0 columns (and weights) have not changed since their Gram matrix
// was formed: O(K) block products instead of O(K^2) (GKS: AV, LV; MMGKS with pnorm = 2: AV).
int tb200_weighted_gram_panel(int64_t m, int64_t k, const double* B, int64_t ld, const double* w, int n_extra,
                              const double* const* extras, const int* extra_weighted, int64_t c0, double* Ghi, double* Glo,
                              double* ws, void* stream) {
  TB200_REQUIRE(c0 >= 0, "bad panel start");
  return gram_launch(m, k, B, ld, w, n_extra, extras, extra_weighted, (int)c0, Ghi, Glo, ws, stream);
}

}  // extern "C"

extern "C" {

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
