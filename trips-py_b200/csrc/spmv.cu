// CSR sparse matrix-vector product for the tomography operator and its explicitly stored transpose.
//
// Replaces the reference's `A @ v` / `A.T @ u` call sites (scipy sparsetools csr_matvec / csc_matvec behind
// trips/utilities/decompositions.py:235-240, trips/solvers/CGLS.py:45-68, GKS.py:82-92, MMGKS.py:43-124).
// A.T is never applied by scatter: the caller stores A^T as a second CSR matrix and calls the same kernel,
// so both directions are gather-only and bitwise reproducible (no atomics).
//
// Two summation orders (argument `order`):
//
//  order 0, "sequential" (default, the parity build).  Golub-Kahan and CGLS run WITHOUT reorthogonalisation in the
//    reference; on CT problems the bases lose orthogonality within ~15 steps and a 1-ulp change anywhere moves the
//    reference's own iterate by ~1e-3 after 50 iterations (measured).  Matching it to 1e-10 therefore
//    needs the SAME arithmetic, not a similar one.  scipy's csr_matvec adds the products of a row one after the
//    other in index order, each multiply and add rounded separately (baseline x86-64 build: no FMA); csc_matvec
//    (A.T @ u) adds the contributions to an output entry in increasing row index, which is the order of the sorted
//    rows of the stored transpose.  The sequential kernels do exactly that, one thread per row for the summation,
//    and are bit-identical to scipy.
//    Bandwidth shape: a warp owns 32 consecutive rows.  Row chunks of CH non-zeros are copied global->shared with
//    cp.async (LDGSTS, fully coalesced: one row chunk = CH consecutive entries), STAGES chunks deep, into a
//    [32][CH+1] tile; lane r then walks row r of the tile (conflict-free: odd pitch), gathers x[col] through the
//    read-only path and extends its own rounding chain.  Warps never synchronise with each other (per-warp tiles,
//    __syncwarp only), so the 8 warps of an SM are 8 independent copy/compute pipelines.  Neighbouring rows are
//    neighbouring rays (or pixels): the 32 gathers of one step hit neighbouring pixels, i.e. few sectors.
//
//  order 1, "tree".  One warp per row, 128-/256-bit streaming loads, FMA, butterfly reduction.  Fastest; differs
//    from scipy by summation order only (rel. 1e-16 per product).
//
// Both fuse the recurrence epilogue  y = A x - coef * z  (decompositions.py:237,240) and ||y||^2.  The norm is
// accumulated in double-double and is the correctly rounded value of the exact sum of squares (tb200_dd.cuh).
#include "tb200_common.cuh"
#include "tb200_dd.cuh"
#include "tb200_ctgeom.cuh"

namespace tb200 {

// ---------------------------------------------------------------------------------------------------------
// cp.async helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 16-byte global->shared copy, L2 only (no L1 allocation); bytes beyond src_bytes are zero-filled.
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst), "l"(src), "r"(src_bytes),
               "l"(policy)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// order 0: sequential (scipy) summation order
// ---------------------------------------------------------------------------------------------------------
// Tile geometry.  Every row stream is taken from its start rounded DOWN to a multiple of four entries, so that
// column indices (4 B) and values (8 or 4 B) both arrive in 16-byte pieces that are 16-byte aligned in global
// memory; the up to three leading entries that belong to the previous row, and whatever follows the row's end in
// its last piece, are masked to zero in registers (only the first and last chunk of a row take that path).
template <typename VT, int WARPS, int CH, int STAGES>
struct SeqTileCfg {
  static constexpr int VPP = 16 / (int)sizeof(VT);  // values per 16-byte piece
  static constexpr int NC = CH / 4;                 // column pieces per row chunk
  static constexpr int NV = CH / VPP;               // value pieces per row chunk
  static constexpr int PC = NC + 1;                 // row pitch in pieces: odd => 128-bit reads by 8 consecutive
  static constexpr int PV = NV + 1;                 //   lanes (one row each) hit 8 distinct 16-byte bank groups
  static_assert((PC & 1) && (PV & 1), "pitches must be odd");
  static constexpr int CSTAGES = STAGES - 1;        // column tiles are dead once their gathers are issued
  static constexpr size_t kColBytesPerWarp = (size_t)CSTAGES * 32 * PC * 16;
  static constexpr size_t kValBytesPerWarp = (size_t)STAGES * 32 * PV * 16;
  static constexpr int PX = CH / 2 + 1;              // x-transpose tile (row-major gather mode): pitch in pieces, odd
  static_assert(PX & 1, "x tile pitch must be odd");
  static constexpr size_t kXBytesPerWarp = (size_t)32 * PX * 16;
  static constexpr size_t kSmemBytes = (size_t)WARPS * (kColBytesPerWarp + kValBytesPerWarp + kXBytesPerWarp);
};

template <typename VT, int CH>
__device__ __forceinline__ void load_vals(const unsigned char* rowbase, double (&v)[CH]);
template <>
__device__ __forceinline__ void load_vals<double, 8>(const unsigned char* rowbase, double (&v)[8]) {
  const double2* p = reinterpret_cast<const double2*>(rowbase);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double2 t = p[j];
    v[2 * j] = t.x;
    v[2 * j + 1] = t.y;
  }
}
template <>
__device__ __forceinline__ void load_vals<double, 16>(const unsigned char* rowbase, double (&v)[16]) {
  const double2* p = reinterpret_cast<const double2*>(rowbase);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double2 t = p[j];
    v[2 * j] = t.x;
    v[2 * j + 1] = t.y;
  }
}
template <>
__device__ __forceinline__ void load_vals<double, 32>(const unsigned char* rowbase, double (&v)[32]) {
  const double2* p = reinterpret_cast<const double2*>(rowbase);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const double2 t = p[j];
    v[2 * j] = t.x;
    v[2 * j + 1] = t.y;
  }
}
template <>
__device__ __forceinline__ void load_vals<float, 8>(const unsigned char* rowbase, double (&v)[8]) {
  const float4* p = reinterpret_cast<const float4*>(rowbase);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float4 t = p[j];
    v[4 * j] = t.x, v[4 * j + 1] = t.y, v[4 * j + 2] = t.z, v[4 * j + 3] = t.w;
  }
}
template <>
__device__ __forceinline__ void load_vals<float, 16>(const unsigned char* rowbase, double (&v)[16]) {
  const float4* p = reinterpret_cast<const float4*>(rowbase);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = p[j];
    v[4 * j] = t.x, v[4 * j + 1] = t.y, v[4 * j + 2] = t.z, v[4 * j + 3] = t.w;
  }
}
template <>
__device__ __forceinline__ void load_vals<float, 32>(const unsigned char* rowbase, double (&v)[32]) {
  const float4* p = reinterpret_cast<const float4*>(rowbase);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = p[j];
    v[4 * j] = t.x, v[4 * j + 1] = t.y, v[4 * j + 2] = t.z, v[4 * j + 3] = t.w;
  }
}

template <typename VT, int WARPS, int CH, int STAGES>
__global__ void __launch_bounds__(WARPS * 32)
spmv_seq_tile_kernel(int64_t m, int64_t nnz, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                     const VT* __restrict__ val, const double* __restrict__ x, double* __restrict__ y, double coef_host,
                     const double* __restrict__ coef_dev, const double* __restrict__ z, double* __restrict__ partials,
                     int force_mode) {
  using Cfg = SeqTileCfg<VT, WARPS, CH, STAGES>;
  constexpr int NC = Cfg::NC, NV = Cfg::NV, PC = Cfg::PC, PV = Cfg::PV, VPP = Cfg::VPP, CSTAGES = Cfg::CSTAGES;
  constexpr int RPI = 32 / NC;  // rows covered by one warp-wide copy instruction (NC lanes per row)
  constexpr int VH = NV / NC;   // value pieces each lane copies per row (2 for fp64, 1 for fp32)
  static_assert(STAGES >= 3, "the gather prefetch needs the tile of chunk it+1 landed while chunk it is consumed");
  static_assert(32 % NC == 0 && NV % NC == 0, "pieces per row must divide the warp");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // carve: [col tiles of all warps][val tiles of all warps]
  unsigned char* ctile = smem_raw + (size_t)warp * Cfg::kColBytesPerWarp;
  unsigned char* vtile = smem_raw + (size_t)WARPS * Cfg::kColBytesPerWarp + (size_t)warp * Cfg::kValBytesPerWarp;
  unsigned char* xtile = smem_raw + (size_t)WARPS * (Cfg::kColBytesPerWarp + Cfg::kValBytesPerWarp) +
                         (size_t)warp * Cfg::kXBytesPerWarp;
  const uint32_t ctile_s = smem_u32(ctile), vtile_s = smem_u32(vtile);

  const uint64_t pol_keep = policy_evict_last();     // the gathered vector: the only data with reuse
  const uint64_t pol_stream = policy_evict_first();  // the matrix stream: read once, must not displace x in L2
  const double coef = (z != nullptr) ? (coef_dev ? *coef_dev : coef_host) : 0.0;
  const int64_t row = ((int64_t)blockIdx.x * WARPS + warp) * 32 + lane;
  int64_t a = 0;  // row start rounded down to a multiple of 4 entries
  int off = 0;    // leading entries of the stream that belong to the previous row
  int tot = 0;    // stream length = off + row length
  if (row < m) {
    const int64_t s = rowptr[row];
    a = s & ~(int64_t)3;
    off = (int)(s - a);
    tot = (int)(rowptr[row + 1] - a);
    if (tot == off) tot = 0, off = 0;  // empty row
  }
  const int nit = (__reduce_max_sync(0xffffffffu, tot) + CH - 1) / CH;

  // Gather mapping for this group of 32 rows.  L1TEX spends one tag cycle per distinct 128-byte line of a gather
  // request, so the request shape should follow the matrix: if the same-position entries of NEIGHBOURING ROWS are
  // close in x (near-vertical rays, pixels of the transpose), lane r gathers for row r (one request = one position
  // of 32 rows); if CONSECUTIVE ENTRIES OF A ROW are close in x (near-horizontal rays: runs of adjacent pixels), a
  // request covers 32 consecutive entries of a row pair and the values reach their owner lanes through shared
  // memory.  Decided once per group by probing the column stride in both directions in the middle of the rows.
  bool row_major = false;
  {
    const int len = tot - off;
    int c0 = 0, c1 = 0;
    const bool ok = len >= 2;
    if (ok) {
      const int64_t i0 = a + off + ((len - 2) >> 1);
      c0 = col[i0];
      c1 = col[i0 + 1];
    }
    const int cn = __shfl_xor_sync(0xffffffffu, c0, 1);
    const bool okn = __shfl_xor_sync(0xffffffffu, (int)ok, 1) != 0;
    const unsigned along = __ballot_sync(0xffffffffu, ok && (c1 - c0) <= 2);
    const unsigned across = __ballot_sync(0xffffffffu, ok && okn && abs(c0 - cn) <= 8);
    row_major = __popc(along) > __popc(across);
    if (force_mode == 1) row_major = false;
    if (force_mode == 2) row_major = true;
  }

  // Producer role of this lane: piece (lane % NC) of the rows q*RPI + lane/NC, q = 0..NC-1 (the same rows for the
  // column and the value tiles), whose stream descriptors are fetched once from the owning lanes.
  const int pc = lane % NC, rsub = lane / NC;
  int64_t pa[NC];  // aligned stream start of my q-th row
  int ptot[NC];    // its stream length
  int pavail[NC];  // entries left in the arrays from its start (clamped): no read beyond the end of colidx / vals
#pragma unroll
  for (int q = 0; q < NC; ++q) {
    const int r = q * RPI + rsub;
    pa[q] = __shfl_sync(0xffffffffu, a, r);
    ptot[q] = __shfl_sync(0xffffffffu, tot, r);
    const int64_t left = nnz - pa[q];
    pavail[q] = left > (int64_t)0x3fffffff ? 0x3fffffff : (int)left;
  }

  // One chunk = CH stream positions of each of the 32 rows: NC column pieces + NV value pieces per row.
  auto issue = [&](int it) {
    const int cslot = it % CSTAGES, vslot = it % STAGES;
    const int pos0 = it * CH;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      const int r = q * RPI + rsub;
      {
        const int pos = pos0 + 4 * pc;
        const int left = (pavail[q] - pos) * 4;
        const int bytes = (pos < ptot[q]) ? (left < 16 ? left : 16) : 0;
        cp_async_16(ctile_s + (uint32_t)(((cslot * 32 + r) * PC + pc) * 16), col + (bytes ? pa[q] + pos : 0), bytes, pol_stream);
      }
#pragma unroll
      for (int h = 0; h < VH; ++h) {
        const int pv = h * NC + pc;
        const int pos = pos0 + VPP * pv;
        const int left = (pavail[q] - pos) * (int)sizeof(VT);
        const int bytes = (pos < ptot[q]) ? (left < 16 ? left : 16) : 0;
        cp_async_16(vtile_s + (uint32_t)(((vslot * 32 + r) * PV + pv) * 16), val + (bytes ? pa[q] + pos : 0), bytes, pol_stream);
      }
    }
  };
  constexpr int RPR = 32 / CH;  // rows per request in row-major mode (a request = CH consecutive entries of RPR rows)
  auto gather = [&](int it, double (&xv)[CH]) {
    if (force_mode == 3) {  // measurement only: no gathers (pipeline ceiling of the stream + rounding chain)
#pragma unroll
      for (int j = 0; j < CH; ++j) xv[j] = 1.0;
    } else if (!row_major) {
      const int4* tc = reinterpret_cast<const int4*>(ctile + (size_t)(((it % CSTAGES) * 32 + lane) * PC) * 16);
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int4 c = tc[j];
        xv[4 * j + 0] = ld_gather_f64(x + c.x, pol_keep);
        xv[4 * j + 1] = ld_gather_f64(x + c.y, pol_keep);
        xv[4 * j + 2] = ld_gather_f64(x + c.z, pol_keep);
        xv[4 * j + 3] = ld_gather_f64(x + c.w, pol_keep);
      }
    } else {
      // request j: lanes -> entries (lane % CH) of rows j*RPR + lane / CH; xv[j] is delivered to its owner later
      const int32_t* tc = reinterpret_cast<const int32_t*>(ctile + (size_t)((it % CSTAGES) * 32 * PC) * 16);
      const int k = lane % CH, rs = lane / CH;
#pragma unroll
      for (int j = 0; j < 32 / RPR; ++j) {
        const int r = j * RPR + rs;
        if (j < CH) xv[j] = ld_gather_f64(x + tc[r * (PC * 4) + k], pol_keep);
      }
    }
  };
  // row-major mode: hand the gathered values to the lanes that own the rows (through the x tile)
  auto deliver = [&](double (&xv)[CH]) {
    if (row_major) {
      double* xt = reinterpret_cast<double*>(xtile);
      const int k = lane % CH, rs = lane / CH;
#pragma unroll
      for (int j = 0; j < 32 / RPR; ++j)
        if (j < CH) xt[(j * RPR + rs) * (Cfg::PX * 2) + k] = xv[j];
      __syncwarp();
      const double2* xr = reinterpret_cast<const double2*>(xtile + (size_t)(lane * Cfg::PX) * 16);
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) {
        const double2 t = xr[j];
        xv[2 * j] = t.x;
        xv[2 * j + 1] = t.y;
      }
      __syncwarp();
    }
  };

  // Software pipeline, per warp:  cp.async of chunk it+STAGES-1  |  x-gathers of chunk it+1  |  add chain of chunk it.
  // The gathered values of chunk it+1 are live across the loop back-edge, so all CH gathers of a lane are in flight
  // while the (latency-bound, strictly ordered) rounding chain of chunk it runs.  Column tiles need one slot less
  // than value tiles: the columns of chunk it are dead once its gathers were issued (iteration it-1).
#pragma unroll
  for (int p = 0; p < STAGES - 1; ++p) {
    if (p < nit) issue(p);
    cp_async_commit();
  }
  double xn[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) xn[k] = 0.0;
  if (nit > 0) {
    cp_async_wait<STAGES - 2>();  // chunk 0 has landed
    __syncwarp();
    gather(0, xn);
  }
  double acc = 0.0;
  for (int it = 0; it < nit; ++it) {
    // value slot (it-1) % STAGES and column slot it % CSTAGES were fully consumed (trailing __syncwarp of it-1)
    if (it + STAGES - 1 < nit) issue(it + STAGES - 1);
    cp_async_commit();
    deliver(xn);
    double v[CH];
    load_vals<VT, CH>(vtile + (size_t)(((it % STAGES) * 32 + lane) * PV) * 16, v);
    // positions outside [lo, hi) are not this row's: zero their value (adds +-0.0: the running sum keeps its bits)
    const int lo = (it == 0) ? off : 0;
    const int hi = tot - it * CH;
    if (lo > 0 || hi < CH) {
#pragma unroll
      for (int k = 0; k < CH; ++k)
        if (k < lo || k >= hi) v[k] = 0.0, xn[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) v[k] = __dmul_rn(v[k], xn[k]);
    // gathers of chunk it+1 (its tile has landed: at most STAGES-2 younger groups may still be pending)
    cp_async_wait<STAGES - 2>();
    __syncwarp();
    if (it + 1 < nit) gather(it + 1, xn);
    // the rounding chain, in index order
#pragma unroll
    for (int k = 0; k < CH; ++k) acc = __dadd_rn(acc, v[k]);
    __syncwarp();
  }
  cp_async_wait<0>();

  dd_t nrm = dd_zero();
  if (row < m) {
    if (z != nullptr) acc = __dsub_rn(acc, __dmul_rn(coef, z[row]));
    y[row] = acc;
    nrm = dd_fma(nrm, acc, acc);
  }
  if (partials != nullptr) {
    const dd_t tot2 = dd_block_sum(nrm, red);
    if (threadIdx.x == 0) {
      partials[2 * (int64_t)blockIdx.x] = tot2.hi;
      partials[2 * (int64_t)blockIdx.x + 1] = tot2.lo;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// order 0, direct variant: no shared memory.
// ncu on the tile kernel above shows the L1TEX data pipe saturated (92 %), 56 % of its wavefronts being the
// shared-memory round trip of the matrix stream (LDGSTS in, LDS out).  Here lane r streams row r straight into
// registers with 256-bit loads: 32 bytes = one full sector per lane per load (8 column indices or 4 values), L2
// prefetch granularity 128 B so the next three loads of the lane hit L2, no L1 allocation.  That is ~4x fewer data
// pipe wavefronts for the stream; what remains are the x-gathers.  Software pipeline per lane:
//   stream loads of chunk it+2  |  gathers of chunk it+1  |  rounding chain of chunk it        (16 entries each)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld_row_i32x8(const int32_t* p, int32_t* c) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7])
               : "l"(p));
}
template <typename VT>
struct RowVals;
template <>
struct RowVals<double> {
  static constexpr int VPP = 4;  // values per 32-byte piece
  __device__ __forceinline__ static void piece(const double* p, double* v) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                 : "l"(p));
  }
};
template <>
struct RowVals<float> {
  static constexpr int VPP = 8;
  __device__ __forceinline__ static void piece(const float* p, double* v) {
    float f[8];
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7])
                 : "l"(p));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (double)f[i];
  }
};

template <typename VT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
spmv_seq_direct_kernel(int64_t m, int64_t nnz, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                       const VT* __restrict__ val, const double* __restrict__ x, double* __restrict__ y, double coef_host,
                       const double* __restrict__ coef_dev, const double* __restrict__ z, double* __restrict__ partials) {
  constexpr int CH = 16;
  constexpr int VPP = RowVals<VT>::VPP;
  __shared__ double red[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t pol_keep = policy_evict_last();
  const double coef = (z != nullptr) ? (coef_dev ? *coef_dev : coef_host) : 0.0;
  const int64_t row = ((int64_t)blockIdx.x * WARPS + warp) * 32 + lane;
  int64_t a = 0;  // row start rounded down to 8 entries: 32-byte aligned for both arrays
  int off = 0, tot = 0;
  if (row < m) {
    const int64_t s = rowptr[row];
    a = s & ~(int64_t)7;
    off = (int)(s - a);
    tot = (int)(rowptr[row + 1] - a);
    if (tot == off) tot = 0, off = 0;
  }
  const int nit = (__reduce_max_sync(0xffffffffu, tot) + CH - 1) / CH;
  const int32_t* cp = col + a;
  const VT* vp = val + a;
  const int64_t room64 = nnz - a;  // entries between the stream start and the end of the arrays
  const int room = room64 > (int64_t)0x3fffffff ? 0x3fffffff : (int)room64;

  // A piece is loaded when it starts inside the row's stream; positions it covers beyond the row (or, for the very
  // last pieces of the arrays, beyond nnz) are masked by position later, so their content is irrelevant - but no
  // byte outside [0, nnz) is ever read.
  auto load_cols = [&](int it, int32_t (&c)[CH]) {
#pragma unroll
    for (int q = 0; q < CH / 8; ++q) {
      const int pos = it * CH + 8 * q;
      if (pos < tot) {
        if (pos + 8 <= room) {
          ld_row_i32x8(cp + pos, &c[8 * q]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) c[8 * q + j] = (pos + j < room) ? cp[pos + j] : 0;
        }
      }
    }
  };
  auto load_vals = [&](int it, double (&v)[CH]) {
#pragma unroll
    for (int q = 0; q < CH / VPP; ++q) {
      const int pos = it * CH + VPP * q;
      if (pos < tot) {
        if (pos + VPP <= room) {
          RowVals<VT>::piece(vp + pos, &v[VPP * q]);
        } else {
#pragma unroll
          for (int j = 0; j < VPP; ++j) v[VPP * q + j] = (pos + j < room) ? (double)vp[pos + j] : 0.0;
        }
      }
    }
  };

  int32_t c1[CH];
  double v0[CH], v1[CH], x0[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) c1[k] = 0, v0[k] = 0.0, v1[k] = 0.0, x0[k] = 0.0;
  if (nit > 0) {
    load_cols(0, c1);
    load_vals(0, v0);
#pragma unroll
    for (int k = 0; k < CH; ++k) x0[k] = ld_gather_f64(x + c1[k], pol_keep);
  }
  if (nit > 1) {
    load_cols(1, c1);
    load_vals(1, v1);
  }
  double acc = 0.0;
  for (int it = 0; it < nit; ++it) {
    // products of chunk it; positions outside [lo, hi) are not this row's and contribute +0.0
    const int lo = (it == 0) ? off : 0;
    const int hi = tot - it * CH;
    double p[CH];
    if (lo > 0 || hi < CH) {
#pragma unroll
      for (int k = 0; k < CH; ++k) p[k] = (k < lo || k >= hi) ? 0.0 : __dmul_rn(v0[k], x0[k]);
    } else {
#pragma unroll
      for (int k = 0; k < CH; ++k) p[k] = __dmul_rn(v0[k], x0[k]);
    }
    // gathers of chunk it+1 (its columns arrived during the previous iteration)
    if (it + 1 < nit) {
#pragma unroll
      for (int k = 0; k < CH; ++k) x0[k] = ld_gather_f64(x + c1[k], pol_keep);
    }
    // stream loads of chunk it+2
    double v2[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) v2[k] = 0.0;
    if (it + 2 < nit) {
      load_cols(it + 2, c1);
      load_vals(it + 2, v2);
    }
    // the rounding chain, in index order
#pragma unroll
    for (int k = 0; k < CH; ++k) acc = __dadd_rn(acc, p[k]);
#pragma unroll
    for (int k = 0; k < CH; ++k) v0[k] = v1[k], v1[k] = v2[k];
  }

  dd_t nrm = dd_zero();
  if (row < m) {
    if (z != nullptr) acc = __dsub_rn(acc, __dmul_rn(coef, z[row]));
    y[row] = acc;
    nrm = dd_fma(nrm, acc, acc);
  }
  if (partials != nullptr) {
    const dd_t tot2 = dd_block_sum(nrm, red);
    if (threadIdx.x == 0) {
      partials[2 * (int64_t)blockIdx.x] = tot2.hi;
      partials[2 * (int64_t)blockIdx.x + 1] = tot2.lo;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// order 0 on the SELL-32-4 layout ("row-interleaved CSR"): the production path for the tomography matrices.
//
// Measured on the CSR kernels above (profiles/): lane-per-row streaming is what the
// sequential rounding chain wants, but on plain CSR it costs the L1TEX data pipe one 32-byte sector per lane per
// load (scattered rows) or a shared-memory round trip; the pipe moves only ~2 such sectors per clock per SM and
// saturates (92 %) below the HBM roofline.  Interleaving the rows of each group of 32 removes that cost without
// touching the arithmetic: entry j of row r lives at  sliceptr[r/32] + (j/4)*128 + (r%32)*4 + j%4,  so the 32
// lanes of a warp (lane = row) read 512 contiguous bytes of column indices / 1 KB of values per load, every lane
// still walks ITS row in index order, and the result is bit-identical to the CSR kernels and to scipy.  Rows of a
// slice are padded to the slice's longest row rounded up to 4 (zeros; < 2 % for the CT matrices).
// Software pipeline per lane, 16 entries per stage:
//   stream loads of chunk it+2  |  x-gathers of chunk it+1  |  rounding chain of chunk it
// ---------------------------------------------------------------------------------------------------------
template <typename VT>
struct SellVals;
template <>
struct SellVals<double> {
  __device__ __forceinline__ static void piece(const double* p, uint64_t, double* v) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                 : "l"(p));
  }
};
template <>
struct SellVals<float> {
  __device__ __forceinline__ static void piece(const float* p, uint64_t pol, double* v) {
    float f[4];
    ld_stream_f32x4(p, pol, f);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (double)f[i];
  }
};

// GEOM = true: the matrix is the parallel-beam CT matrix A (rows = rays) and its VALUES ARE NOT STORED: an entry is
// re-evaluated from its column index and the ray's geometry (tb200_ctgeom.cuh, ~9 fp64 instructions, same bits as the
// builder writes) while the index stream - 4 bytes per entry instead of 12 - and the gathers are in flight.
// out[c*rows + r] = in[r*cols + c] (rows x cols, row-major), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
transpose_f64_kernel(int rows, int cols, const double* __restrict__ in, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[(int64_t)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(int64_t)c * rows + r] = tile[threadIdx.x][j];
  }
}

struct CtRays {
  const double* geom;  // 6 doubles per angle: c, s, d2, 1/hi, 1/(hi*lo), 0
  const int32_t* cta_order;  // nullable: CTA b works on slice group cta_order[b] (longest first: no straggler tail)
  const int32_t* rowskip;    // nullable: row r's entries start at position rowskip[r] of its lane (row-aligned slices)
  const double* xT;          // nullable: the transposed image; rays with |sin| > |cos| then carry indices ix*ny + iy
  uint32_t ny_magic;         // as nx_magic, for the division by ny of those indices
  int ny_shift;
  int nx, ny, n_det;
  uint32_t nx_magic;   // floor(2^(32+nx_shift) / nx) clipped to 2^32-1: col / nx = umulhi(col, magic) >> shift (+1 fix-up)
  int nx_shift;
};

template <typename VT, int WARPS, bool GEOM, int CH>
__global__ void __launch_bounds__(WARPS * 32, GEOM ? 512 / (WARPS * 32) : (CH == 8 ? 512 / (WARPS * 32) : 1))
spmv_sell_kernel(int64_t m, const int64_t* __restrict__ sliceptr, const int32_t* __restrict__ rowlen,
                 const int32_t* __restrict__ col, const VT* __restrict__ val, const double* __restrict__ x,
                 double* __restrict__ y, double coef_host, const double* __restrict__ coef_dev,
                 const double* __restrict__ z, double* __restrict__ partials, int gather_mode, CtRays ct) {
  // CH = entries per pipeline stage and lane: 16 for stored values (deep stream prefetch, HBM bound), 8 when the
  // values are computed (GEOM: issue bound, wants four times the warps, so half the registers per lane)
  constexpr int RPR = 32 / CH;        // row-major gathers: rows covered by one request
  constexpr int CT_LD = CH + 4;       // ctile row pitch (ints): 16-byte aligned rows
  constexpr int XT_LD = CH + 2;       // xtile row pitch (doubles): odd number of 16-byte pieces
  __shared__ double red[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t pol_keep = policy_evict_last();
  const uint64_t pol_stream = policy_evict_first();
  const double coef = (z != nullptr) ? (coef_dev ? *coef_dev : coef_host) : 0.0;
  const int64_t group = (ct.cta_order != nullptr) ? (int64_t)ct.cta_order[blockIdx.x] : (int64_t)blockIdx.x;
  const int64_t slice = group * WARPS + warp;
  const int64_t nslices = (m + 31) >> 5;
  const int64_t row = slice * 32 + lane;
  int64_t base = 0;
  int w = 0, len = 0, skip = 0;  // slice width (positions per lane, multiple of 4); this lane's entries: [skip, skip+len)
  if (slice < nslices) {
    base = sliceptr[slice];
    w = (int)((sliceptr[slice + 1] - base) >> 5);
    if (row < m) {
      len = rowlen[row];
      if (ct.rowskip != nullptr) skip = ct.rowskip[row];
    }
  }
  const int nit = (w + CH - 1) / CH;
  const int32_t* cp = col + base + lane * 4;
  const VT* vp = val + base + lane * 4;

  // GEOM: this lane's ray.  An index is q*div + r: (q, r) = (iy, ix) in the image, or (ix, iy) for a shallow ray that
  // addresses the transposed image.  proj = cx*c + cy*s is formed as r-term + q-term either way (a + b == b + a).
  RayGeom g = {1.0, 0.0, 0.5, 1.0, 1.0};
  double sd = 0.0, bias_q = 0.0, bias_r = 0.0, coef_q = 0.0, coef_r = 0.0;
  uint32_t dmagic = ct.nx_magic;
  int dshift = ct.nx_shift, ddiv = ct.nx;
  const double* xb = x;  // where this lane's indices point
  if (GEOM) {
    const int64_t r = (row < m) ? row : 0;
    const int a = (int)(r / ct.n_det), d = (int)(r - (int64_t)a * ct.n_det);
    const double* gp = ct.geom + 6 * (int64_t)a;
    g.c = gp[0], g.s = gp[1], g.d2 = gp[2], g.inv_hi = gp[3], g.inv_hilo = gp[4];
    sd = (double)d - 0.5 * (double)(ct.n_det - 1);
    const bool tmode = ct.xT != nullptr && fabs(g.s) > fabs(g.c);
    if (tmode) {
      xb = ct.xT, dmagic = ct.ny_magic, dshift = ct.ny_shift, ddiv = ct.ny;
      bias_q = centred_bias(ct.nx), coef_q = g.c, bias_r = centred_bias(ct.ny), coef_r = g.s;
    } else {
      bias_q = centred_bias(ct.ny), coef_q = g.s, bias_r = centred_bias(ct.nx), coef_r = g.c;
    }
  }
  if (!GEOM && ct.xT != nullptr) {
    // stored values over the row-aligned CT index layout (tb200_ct_spmv_sell_f64): the indices of a shallow ray address
    // the transposed image, exactly as in the GEOM case - only the gather base depends on the ray
    const int64_t r = (row < m) ? row : 0;
    const double* gp = ct.geom + 6 * (r / ct.n_det);
    if (fabs(gp[1]) > fabs(gp[0])) xb = ct.xT;
  }
  auto entry_values = [&](const int32_t (&c)[CH], double (&v)[CH]) {
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      int q = (int)(__umulhi((uint32_t)c[k], dmagic) >> dshift);
      int r = c[k] - q * ddiv;
      if (r >= ddiv) r -= ddiv, ++q;
      const double proj = __dadd_rn(__dmul_rn(centred_coord(r, bias_r), coef_r), __dmul_rn(centred_coord(q, bias_q), coef_q));
      v[k] = chord(g, __dsub_rn(sd, proj));
    }
  };

  auto load_cols = [&](int it, int32_t (&c)[CH]) {
#pragma unroll
    for (int q = 0; q < CH / 4; ++q)
      if (it * CH + 4 * q < w) {
        int32_t t[4];
        ld_stream_i32x4(cp + (int64_t)(it * (CH / 4) + q) * 128, pol_stream, t);
#pragma unroll
        for (int i = 0; i < 4; ++i) c[4 * q + i] = t[i];
      }
  };
  auto load_vals = [&](int it, double (&v)[CH]) {
#pragma unroll
    for (int q = 0; q < CH / 4; ++q)
      if (it * CH + 4 * q < w) SellVals<VT>::piece(vp + (int64_t)(it * (CH / 4) + q) * 128, pol_stream, &v[4 * q]);
  };

  // Gather mapping of this slice.  The L1TEX data pipe pays per distinct 32-byte sector of a gather request, so the
  // request shape should follow the matrix.  Lane-per-row requests (one position of 32 rows) are cheap when the rows
  // are near-vertical rays or pixels of the transpose: neighbouring rows hit neighbouring x.  For near-horizontal rays
  // (runs of adjacent pixels along the row, neighbouring rays in DIFFERENT image rows) they touch 32 sectors each:
  // there a request is re-shaped to 16 consecutive entries of two rows (2-3 sectors) and the values are handed to
  // the owner lanes through a per-warp shared-memory tile.  Arithmetic and order are unchanged (bit-identical).
  // Decided once per slice by probing the column stride along the row and across the rows in the first chunk.
  __shared__ __align__(16) int32_t ctile_all[WARPS][32 * CT_LD];  // [row][CH cols + 4 pad]
  __shared__ __align__(16) double xtile_all[WARPS][32 * XT_LD];   // [row][CH x + 2 pad]
  int32_t* ctile = ctile_all[warp];
  double* xtile = xtile_all[warp];
  const int gk = lane % CH, gr = lane / CH;  // row-major mode: my entry / which row of the request's group
  auto gather_lane_per_row = [&](const int32_t (&c)[CH], double (&xv)[CH]) {
#pragma unroll
    for (int k = 0; k < CH; ++k) xv[k] = ld_gather_f64(xb + c[k], pol_keep);
  };
  auto gather_row_major = [&](const int32_t (&c)[CH], double (&xv)[CH]) {
    // publish my row's columns, then request j gathers entries 0..CH-1 of rows RPR*j .. RPR*j + RPR-1
#pragma unroll
    for (int q = 0; q < CH / 4; ++q)
      *reinterpret_cast<int4*>(ctile + lane * CT_LD + 4 * q) = make_int4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < CH; ++j) xv[j] = ld_gather_f64(x + ctile[(RPR * j + gr) * CT_LD + gk], pol_keep);
    __syncwarp();
  };
  auto deliver_row_major = [&](double (&xv)[CH]) {
    // xv[j] currently holds x for (row RPR*j + gr, entry gk): hand every value to the lane that owns its row
#pragma unroll
    for (int j = 0; j < CH; ++j) xtile[(RPR * j + gr) * XT_LD + gk] = xv[j];
    __syncwarp();
    const double2* xr = reinterpret_cast<const double2*>(xtile + lane * XT_LD);
#pragma unroll
    for (int j = 0; j < CH / 2; ++j) {
      const double2 t = xr[j];
      xv[2 * j] = t.x;
      xv[2 * j + 1] = t.y;
    }
    __syncwarp();
  };

  int32_t c1[CH];
  double v0[CH], v1[CH], x0[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) c1[k] = 0, v0[k] = 0.0, v1[k] = 0.0, x0[k] = 0.0;
  bool row_major = false;
  if (nit > 0) {
    load_cols(0, c1);
    if (!GEOM) load_vals(0, v0);
    {
      constexpr int PB = CH / 2;  // probe position inside the first chunk
      const bool ok = skip == 0 && len >= PB + 2;
      const int cn = __shfl_xor_sync(0xffffffffu, c1[PB], 1);
      const bool okn = __shfl_xor_sync(0xffffffffu, (int)ok, 1) != 0;
      const unsigned along = __ballot_sync(0xffffffffu, ok && (c1[PB + 1] - c1[PB]) <= 2 && (c1[PB] - c1[PB - 1]) <= 2);
      const unsigned across = __ballot_sync(0xffffffffu, ok && okn && abs(c1[PB] - cn) <= 6);
      row_major = __popc(along) > __popc(across) + 8;
      if (gather_mode == 1 || ct.xT != nullptr) row_major = false;  // shallow rays read the transposed image
      if (gather_mode == 2) row_major = true;
    }
    if (row_major) gather_row_major(c1, x0);
    else gather_lane_per_row(c1, x0);
    if (GEOM) entry_values(c1, v0);
  }
  if (nit > 1) {
    load_cols(1, c1);
    if (!GEOM) load_vals(1, v1);
  }
  double acc = 0.0;
  for (int it = 0; it < nit; ++it) {
    if (row_major) deliver_row_major(x0);
    // products of chunk it; positions outside [skip, skip + len) are padding and contribute +0.0
    const int lo = skip - it * CH, hi = skip + len - it * CH;
    double p[CH];
    if (hi < CH || lo > 0) {
#pragma unroll
      for (int k = 0; k < CH; ++k) p[k] = (k >= hi || k < lo) ? 0.0 : __dmul_rn(v0[k], x0[k]);
    } else {
#pragma unroll
      for (int k = 0; k < CH; ++k) p[k] = __dmul_rn(v0[k], x0[k]);
    }
    // gathers of chunk it+1 (its columns arrived during the previous iteration)
    if (it + 1 < nit) {
      if (row_major) gather_row_major(c1, x0);
      else gather_lane_per_row(c1, x0);
      if (GEOM) entry_values(c1, v1);  // values of chunk it+1 while its gathers fly
    }
    // stream loads of chunk it+2
    double v2[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) v2[k] = 0.0;
    if (it + 2 < nit) {
      load_cols(it + 2, c1);
      if (!GEOM) load_vals(it + 2, v2);
    }
    // the rounding chain, in index order
#pragma unroll
    for (int k = 0; k < CH; ++k) acc = __dadd_rn(acc, p[k]);
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      v0[k] = v1[k];
      if (!GEOM) v1[k] = v2[k];
    }
  }

  dd_t nrm = dd_zero();
  if (row < m) {
    if (z != nullptr) acc = __dsub_rn(acc, __dmul_rn(coef, z[row]));
    y[row] = acc;
    nrm = dd_fma(nrm, acc, acc);
  }
  if (partials != nullptr) {
    const dd_t tot2 = dd_block_sum(nrm, red);
    if (threadIdx.x == 0) {
      partials[2 * (int64_t)blockIdx.x] = tot2.hi;
      partials[2 * (int64_t)blockIdx.x + 1] = tot2.lo;
    }
  }
}

// One thread per row, for short rows (generic CSR handed in by a user, e.g. a sparse regularisation matrix).
template <typename VT>
__global__ void __launch_bounds__(256)
spmv_seq_scalar_kernel(int64_t m, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                       const VT* __restrict__ val, const double* __restrict__ x, double* __restrict__ y, double coef_host,
                       const double* __restrict__ coef_dev, const double* __restrict__ z, double* __restrict__ partials) {
  __shared__ double red[64];
  const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const double coef = (z != nullptr) ? (coef_dev ? *coef_dev : coef_host) : 0.0;
  dd_t nrm = dd_zero();
  if (row < m) {
    double acc = 0.0;
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    for (int64_t i = s; i < e; ++i) acc = __dadd_rn(acc, __dmul_rn((double)val[i], x[col[i]]));
    if (z != nullptr) acc = __dsub_rn(acc, __dmul_rn(coef, z[row]));
    y[row] = acc;
    nrm = dd_fma(nrm, acc, acc);
  }
  if (partials != nullptr) {
    const dd_t tot = dd_block_sum(nrm, red);
    if (threadIdx.x == 0) {
      partials[2 * (int64_t)blockIdx.x] = tot.hi;
      partials[2 * (int64_t)blockIdx.x + 1] = tot.lo;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// order 1: tree summation (one warp per row, wide streaming loads)
// ---------------------------------------------------------------------------------------------------------
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

template <typename VT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
spmv_warp_kernel(int64_t m, int rows_per_cta, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                 const VT* __restrict__ val, const double* __restrict__ x, double* __restrict__ y, double coef_host,
                 const double* __restrict__ coef_dev, const double* __restrict__ z, double* __restrict__ partials) {
  __shared__ double red[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t pol_stream = policy_evict_first();
  const uint64_t pol_keep = policy_evict_last();
  const double coef = (z != nullptr) ? (coef_dev ? *coef_dev : coef_host) : 0.0;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  dd_t nrm = dd_zero();

  for (int r = warp; r < rows_per_cta; r += WARPS) {
    const int64_t row = row0 + r;
    if (row >= m) break;
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    int64_t a = (s + 3) & ~(int64_t)3;
    if (a > e) a = e;
    double acc0 = 0.0, acc1 = 0.0;
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
    const int64_t t0 = a + 4 * ngrp;
    if (t0 + lane < e) acc1 = fma((double)val[t0 + lane], ld_gather_f64(x + col[t0 + lane], pol_keep), acc1);
    double sum = warp_sum(acc0 + acc1);
    if (lane == 0) {
      if (z != nullptr) sum = __dsub_rn(sum, __dmul_rn(coef, z[row]));
      y[row] = sum;
      nrm = dd_fma(nrm, sum, sum);
    }
  }
  if (partials != nullptr) {
    const dd_t tot = dd_block_sum(nrm, red);
    if (threadIdx.x == 0) {
      partials[2 * (int64_t)blockIdx.x] = tot.hi;
      partials[2 * (int64_t)blockIdx.x + 1] = tot.lo;
    }
  }
}

// T threads per row (T in {2,4,8,16}); scalar loads; tree order.
template <typename VT, int T>
__global__ void __launch_bounds__(256)
spmv_subwarp_kernel(int64_t m, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    const VT* __restrict__ val, const double* __restrict__ x, double* __restrict__ y, double coef_host,
                    const double* __restrict__ coef_dev, const double* __restrict__ z, double* __restrict__ partials) {
  __shared__ double red[64];
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
  dd_t nrm = dd_zero();
  if (row < m && sub == 0) {
    if (z != nullptr) acc = __dsub_rn(acc, __dmul_rn(coef, z[row]));
    y[row] = acc;
    nrm = dd_fma(nrm, acc, acc);
  }
  if (partials != nullptr) {
    const dd_t tot = dd_block_sum(nrm, red);
    if (threadIdx.x == 0) {
      partials[2 * (int64_t)blockIdx.x] = tot.hi;
      partials[2 * (int64_t)blockIdx.x + 1] = tot.lo;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static int g_seq_variant = 0;  // tuning knob (tb200_spmv_set_variant): which tile configuration order 0 uses
static int g_seq_gather_mode = 0;  // 0 = decide per row group, 1 = always lane-per-row gathers, 2 = always row-major gathers

template <typename VT, int WARPS, int CH, int STAGES>
static int launch_seq_tile(int64_t m, int64_t nnz, const int64_t* rowptr, const int32_t* col, const VT* val, const double* x, double* y,
                           double coef_host, const double* coef_dev, const double* z, double* partials, int64_t* nblocks,
                           cudaStream_t st) {
  using Cfg = SeqTileCfg<VT, WARPS, CH, STAGES>;
  auto kern = spmv_seq_tile_kernel<VT, WARPS, CH, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("spmv_seq_tile: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  const int64_t rows_per_cta = (int64_t)WARPS * 32;
  *nblocks = (m + rows_per_cta - 1) / rows_per_cta;
  kern<<<(unsigned)*nblocks, WARPS * 32, Cfg::kSmemBytes, st>>>(m, nnz, rowptr, col, val, x, y, coef_host, coef_dev, z, partials,
                                                                g_seq_gather_mode);
  return check_launch("spmv_seq_tile");
}

template <typename VT, int WARPS>
static int launch_seq_direct(int64_t m, int64_t nnz, const int64_t* rowptr, const int32_t* col, const VT* val,
                             const double* x, double* y, double coef_host, const double* coef_dev, const double* z,
                             double* partials, int64_t* nblocks, cudaStream_t st) {
  const int64_t rows_per_cta = (int64_t)WARPS * 32;
  *nblocks = (m + rows_per_cta - 1) / rows_per_cta;
  spmv_seq_direct_kernel<VT, WARPS><<<(unsigned)*nblocks, WARPS * 32, 0, st>>>(m, nnz, rowptr, col, val, x, y, coef_host,
                                                                              coef_dev, z, partials);
  return check_launch("spmv_seq_direct");
}

template <typename VT>
static int spmv_launch(int order, int64_t m, int64_t nnz, const int64_t* rowptr, const int32_t* col, const VT* val,
                       const double* x, double* y, double coef_host, const double* coef_dev, const double* z,
                       double* norm_out, double* ws, cudaStream_t st) {
  double* partials = norm_out ? ws : nullptr;
  const double avg = (m > 0) ? (double)nnz / (double)m : 0.0;
  int64_t nblocks = 1;
  int rc = 0;
  if (order == 0) {
    if (avg >= 24.0) {
      switch (g_seq_variant) {
        case 4: rc = launch_seq_direct<VT, 2>(m, nnz, rowptr, col, val, x, y, coef_host, coef_dev, z, partials, &nblocks, st); break;
        case 5: rc = launch_seq_direct<VT, 4>(m, nnz, rowptr, col, val, x, y, coef_host, coef_dev, z, partials, &nblocks, st); break;
        case 6: rc = launch_seq_direct<VT, 1>(m, nnz, rowptr, col, val, x, y, coef_host, coef_dev, z, partials, &nblocks, st); break;
        case 1: rc = launch_seq_tile<VT, 2, 16, 3>(m, nnz, rowptr, col, val, x, y, coef_host, coef_dev, z, partials, &nblocks, st); break;
        case 2: rc = launch_seq_tile<VT, 4, 16, 3>(m, nnz, rowptr, col, val, x, y, coef_host, coef_dev, z, partials, &nblocks, st); break;
        case 3: rc = launch_seq_tile<VT, 1, 8, 4>(m, nnz, rowptr, col, val, x, y, coef_host, coef_dev, z, partials, &nblocks, st); break;
        default: rc = launch_seq_tile<VT, 1, 16, 3>(m, nnz, rowptr, col, val, x, y, coef_host, coef_dev, z, partials, &nblocks, st); break;
      }
    } else {
      nblocks = (m + 255) / 256;
      spmv_seq_scalar_kernel<VT><<<(unsigned)nblocks, 256, 0, st>>>(m, rowptr, col, val, x, y, coef_host, coef_dev, z, partials);
      rc = check_launch("spmv_seq_scalar");
    }
  } else if (avg >= 48.0) {
    const int rows_per_cta = 32;
    nblocks = (m + rows_per_cta - 1) / rows_per_cta;
    spmv_warp_kernel<VT, 8><<<(unsigned)nblocks, 256, 0, st>>>(m, rows_per_cta, rowptr, col, val, x, y, coef_host, coef_dev, z, partials);
    rc = check_launch("spmv_warp");
  } else {
    int t = 2;
    while (t < 16 && t < avg) t <<= 1;
    nblocks = (m * t + 255) / 256;
    switch (t) {
      case 2: spmv_subwarp_kernel<VT, 2><<<(unsigned)nblocks, 256, 0, st>>>(m, rowptr, col, val, x, y, coef_host, coef_dev, z, partials); break;
      case 4: spmv_subwarp_kernel<VT, 4><<<(unsigned)nblocks, 256, 0, st>>>(m, rowptr, col, val, x, y, coef_host, coef_dev, z, partials); break;
      case 8: spmv_subwarp_kernel<VT, 8><<<(unsigned)nblocks, 256, 0, st>>>(m, rowptr, col, val, x, y, coef_host, coef_dev, z, partials); break;
      default: spmv_subwarp_kernel<VT, 16><<<(unsigned)nblocks, 256, 0, st>>>(m, rowptr, col, val, x, y, coef_host, coef_dev, z, partials); break;
    }
    rc = check_launch("spmv_subwarp");
  }
  if (rc) return rc;
  if (norm_out) {
    finalize_dd_kernel<<<1, 1024, 0, st>>>(ws, nblocks, norm_out);
    rc = check_launch("spmv finalize");
  }
  return rc;
}

static int g_sell_warps = 4;  // tuning knob (tb200_spmv_set_variant, bits 8-9: 0/2 -> 4, 1 -> 2, 3 -> 1 warps per CTA)
static int g_sell_ch8 = 0;    // tuning knob (bit 10): stored values with 8 entries per pipeline stage and 16 warps per SM

template <typename VT, bool GEOM = false, int CH = 16>
static int sell_launch(int64_t m, const int64_t* sliceptr, const int32_t* rowlen, const int32_t* col, const VT* val,
                       const double* x, double* y, double coef_host, const double* coef_dev, const double* z,
                       double* norm_out, double* ws, cudaStream_t st, CtRays ct = CtRays()) {
  double* partials = norm_out ? ws : nullptr;
  const int64_t nslices = (m + 31) / 32;
  const int warps = g_sell_warps;
  const int64_t nblocks = (nslices + warps - 1) / warps;
  switch (warps) {
    case 1: spmv_sell_kernel<VT, 1, GEOM, CH><<<(unsigned)nblocks, 32, 0, st>>>(m, sliceptr, rowlen, col, val, x, y, coef_host, coef_dev, z, partials, g_seq_gather_mode, ct); break;
    case 2: spmv_sell_kernel<VT, 2, GEOM, CH><<<(unsigned)nblocks, 64, 0, st>>>(m, sliceptr, rowlen, col, val, x, y, coef_host, coef_dev, z, partials, g_seq_gather_mode, ct); break;
    default: spmv_sell_kernel<VT, 4, GEOM, CH><<<(unsigned)nblocks, 128, 0, st>>>(m, sliceptr, rowlen, col, val, x, y, coef_host, coef_dev, z, partials, g_seq_gather_mode, ct); break;
  }
  int rc = check_launch("spmv_sell");
  if (rc) return rc;
  if (norm_out) {
    finalize_dd_kernel<<<1, 1024, 0, st>>>(ws, nblocks, norm_out);
    rc = check_launch("spmv_sell finalize");
  }
  return rc;
}

}  // namespace tb200

using namespace tb200;

extern "C" {

// Number of doubles of workspace a fused-norm SpMV over m rows may need (upper bound over all plans):
// one double-double partial per CTA, smallest CTA footprint is 16 rows.
int64_t tb200_spmv_workspace_len(int64_t m) { return 2 * ((m + 15) / 16 + 8); }  // >= one dd partial per 32 rows

// How many of this library's kernels one call enqueues (for launch accounting in bench.py).
int tb200_spmv_launches(int with_norm) { return with_norm ? 2 : 1; }

// Tuning knob for the order-0 tile kernel (warps per CTA x entries per chunk x stages): 0 = 1 x 16 x 3 (default),
// 1 = 2 x 16 x 3, 2 = 4 x 16 x 3, 3 = 1 x 8 x 4; 4/5/6 = direct (no shared memory) kernel with 2/4/1 warps per CTA;
// plus 8 * gather mode of the tile kernel (0 = per row group, 1 = lane-per-row, 2 = row-major, 3 = no gathers:
// measurement only).  Results are bit-identical across variants (except gather mode 3).
int tb200_spmv_set_variant(int v) {
  TB200_REQUIRE(v >= 0 && (v & 7) <= 6 && ((v >> 3) & 3) <= 3 && (v >> 8) <= 7,
                "variant = CSR kernel (0..6) + 8 * gather mode (0..3) + 256 * log2(SELL warps per CTA) (0..3) + 1024 * (SELL 8-entry stages)");
  g_seq_variant = v & 7;
  g_seq_gather_mode = (v >> 3) & 3;
  g_sell_warps = (((v >> 8) & 3) == 1) ? 2 : (((v >> 8) & 3) == 3) ? 1 : 4;
  g_sell_ch8 = (v >> 10) & 1;
  return 0;
}

static int check_spmv_args(int order, int64_t m, int64_t n, int64_t nnz, const void* rowptr, const void* col,
                           const void* val, const void* x, const void* y, const void* norm_out, const void* ws) {
  TB200_REQUIRE(order == 0 || order == 1, "order must be 0 (sequential) or 1 (tree)");
  TB200_REQUIRE(m >= 0 && n >= 0 && nnz >= 0, "negative size");
  TB200_REQUIRE(n < ((int64_t)1 << 31), "n must fit int32 column indices");
  TB200_REQUIRE(rowptr && x && y, "null pointer");
  TB200_REQUIRE(nnz == 0 || (col && val), "null matrix arrays");
  TB200_REQUIRE(((uintptr_t)val % 32) == 0 && ((uintptr_t)col % 16) == 0, "vals must be 32-byte and colidx 16-byte aligned");
  TB200_REQUIRE(norm_out == nullptr || ws != nullptr, "norm_out requires a workspace");
  return 0;
}

int tb200_spmv_csr_f64(int order, int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* colidx,
                       const double* vals, const double* x, double* y, double coef_host, const double* coef_dev,
                       const double* z, double* norm_out, double* ws, void* stream) {
  int rc = check_spmv_args(order, m, n, nnz, rowptr, colidx, vals, x, y, norm_out, ws);
  if (rc) return rc;
  if (m == 0) return 0;
  return spmv_launch<double>(order, m, nnz, rowptr, colidx, vals, x, y, coef_host, coef_dev, z, norm_out, ws,
                             (cudaStream_t)stream);
}

int tb200_spmv_csr_f32s(int order, int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* colidx,
                        const float* vals, const double* x, double* y, double coef_host, const double* coef_dev,
                        const double* z, double* norm_out, double* ws, void* stream) {
  int rc = check_spmv_args(order, m, n, nnz, rowptr, colidx, vals, x, y, norm_out, ws);
  if (rc) return rc;
  if (m == 0) return 0;
  return spmv_launch<float>(order, m, nnz, rowptr, colidx, vals, x, y, coef_host, coef_dev, z, norm_out, ws,
                            (cudaStream_t)stream);
}

// SELL-32-4 ("row-interleaved CSR") SpMV, scipy summation order: y = A x - coef*z, optional fused ||y||^2.
// sliceptr: int64[ceil(m/32)+1] (multiples of 128), rowlen: int32[m], colidx/vals: sliceptr[last] entries, entry j of
// row r at sliceptr[r/32] + (j/4)*128 + (r%32)*4 + j%4, padding zero-filled.  colidx 512-byte, vals 1024-byte aligned.
static int check_sell_args(int64_t m, int64_t n, const void* sliceptr, const void* rowlen, const void* col,
                           const void* val, const void* x, const void* y, const void* norm_out, const void* ws) {
  TB200_REQUIRE(m >= 0 && n >= 0 && n < ((int64_t)1 << 31), "bad size");
  TB200_REQUIRE(sliceptr && rowlen && x && y, "null pointer");
  TB200_REQUIRE(((uintptr_t)val % 32) == 0 && ((uintptr_t)col % 16) == 0, "vals must be 32-byte and colidx 16-byte aligned");
  TB200_REQUIRE(norm_out == nullptr || ws != nullptr, "norm_out requires a workspace");
  return 0;
}

int tb200_spmv_sell_f64(int64_t m, int64_t n, const int64_t* sliceptr, const int32_t* rowlen, const int32_t* colidx,
                        const double* vals, const double* x, double* y, double coef_host, const double* coef_dev,
                        const double* z, double* norm_out, double* ws, void* stream) {
  int rc = check_sell_args(m, n, sliceptr, rowlen, colidx, vals, x, y, norm_out, ws);
  if (rc) return rc;
  if (m == 0) return 0;
  if (g_sell_ch8)
    return sell_launch<double, false, 8>(m, sliceptr, rowlen, colidx, vals, x, y, coef_host, coef_dev, z, norm_out, ws,
                                         (cudaStream_t)stream);
  return sell_launch<double>(m, sliceptr, rowlen, colidx, vals, x, y, coef_host, coef_dev, z, norm_out, ws,
                             (cudaStream_t)stream);
}

int tb200_spmv_sell_f32s(int64_t m, int64_t n, const int64_t* sliceptr, const int32_t* rowlen, const int32_t* colidx,
                         const float* vals, const double* x, double* y, double coef_host, const double* coef_dev,
                         const double* z, double* norm_out, double* ws, void* stream) {
  int rc = check_sell_args(m, n, sliceptr, rowlen, colidx, vals, x, y, norm_out, ws);
  if (rc) return rc;
  if (m == 0) return 0;
  return sell_launch<float>(m, sliceptr, rowlen, colidx, vals, x, y, coef_host, coef_dev, z, norm_out, ws,
                            (cudaStream_t)stream);
}

// Parallel-beam CT forward projection y = A x - coef*z with the VALUES OF A RE-EVALUATED ON THE FLY (see CtRays above):
// only the SELL-32-4 column indices of A are read.  geom: tb200_ct_geometry output for the n_ang angles of A's rows
// (row = angle*n_det + detector).  Bit-identical to tb200_spmv_sell_f64 on the matrix tb200_ct_fill_rows writes.
// cta_order (nullable): a permutation of the ceil(ceil(m/32)/4) groups of four slices, heaviest first, so that the
// long central rays are scheduled before the short peripheral ones.  rowskip (nullable): leading padding of each row
// inside its lane (tb200_ct_fill_rows_aligned).  xT_scratch: nx*ny doubles, REQUIRED iff the indices were filled with
// transpose_shallow != 0 (the image is transposed into it first), NULL otherwise.
static int ct_rays_setup(int nx, int ny, int n_det, const double* geom, const int32_t* rowskip, const int32_t* cta_order,
                         double* xT_scratch, const double* x, void* stream, CtRays& ct) {
  ct.geom = geom;
  ct.cta_order = (g_sell_warps == 4) ? cta_order : nullptr;  // the order is a permutation of groups of FOUR slices
  ct.rowskip = rowskip;
  ct.xT = xT_scratch;
  int shy = 0;
  while ((2u << shy) <= (uint32_t)ny) ++shy;
  const uint64_t mgy = (((uint64_t)1) << (32 + shy)) / (uint64_t)ny;
  ct.ny_magic = (uint32_t)(mgy > 0xffffffffull ? 0xffffffffull : mgy);
  ct.ny_shift = shy;
  if (xT_scratch != nullptr) {  // xT[ix*ny + iy] = x[iy*nx + ix]
    const dim3 tg((unsigned)((nx + 31) / 32), (unsigned)((ny + 31) / 32));
    transpose_f64_kernel<<<tg, dim3(32, 8), 0, (cudaStream_t)stream>>>(ny, nx, x, xT_scratch);
    int rc = check_launch("ct_forward transpose");
    if (rc) return rc;
  }
  ct.nx = nx, ct.ny = ny, ct.n_det = n_det;
  int sh = 0;
  while ((2u << sh) <= (uint32_t)nx) ++sh;  // floor(log2(nx))
  const uint64_t mg = (((uint64_t)1) << (32 + sh)) / (uint64_t)nx;
  ct.nx_magic = (uint32_t)(mg > 0xffffffffull ? 0xffffffffull : mg);
  ct.nx_shift = sh;
  return 0;
}

int tb200_ct_forward_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const int64_t* sliceptr,
                         const int32_t* rowlen, const int32_t* rowskip, const int32_t* colidx, const int32_t* cta_order,
                         double* xT_scratch, const double* x, double* y, double coef_host, const double* coef_dev,
                         const double* z, double* norm_out, double* ws, void* stream) {
  TB200_REQUIRE(nx > 0 && ny > 0 && n_det > 0 && n_ang >= 0 && (int64_t)nx * ny < ((int64_t)1 << 31), "bad geometry");
  const int64_t m = (int64_t)n_ang * n_det;
  int rc = check_sell_args(m, (int64_t)nx * ny, sliceptr, rowlen, colidx, nullptr, x, y, norm_out, ws);
  if (rc) return rc;
  TB200_REQUIRE(geom != nullptr || m == 0, "null geometry table");
  if (m == 0) return 0;
  CtRays ct;
  rc = ct_rays_setup(nx, ny, n_det, geom, rowskip, cta_order, xT_scratch, x, stream, ct);
  if (rc) return rc;
  return sell_launch<double, true, 8>(m, sliceptr, rowlen, colidx, (const double*)nullptr, x, y, coef_host, coef_dev, z,
                                   norm_out, ws, (cudaStream_t)stream, ct);
}

// y = A x - coef*z with the STORED values of the parallel-beam CT matrix over the same row-aligned index layout
// (tb200_ct_fill_rows_aligned_vals): the stored SELL-32-4 product whose x-gathers share sectors the way the index-only
// projector's do - rowskip aligns the rays of a slice on the image rows, shallow rays address the transposed image formed
// in xT_scratch (REQUIRED iff filled with transpose_shallow != 0), cta_order schedules the long slices first.
// vals: double, or float when vals_f32 != 0 (fp32 storage, fp64 accumulation).  Same bits as tb200_spmv_sell_f64 on the
// plain layout: the order in which a row's entries are added is the same.
int tb200_ct_spmv_sell_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const int64_t* sliceptr,
                           const int32_t* rowlen, const int32_t* rowskip, const int32_t* colidx, const void* vals,
                           int vals_f32, const int32_t* cta_order, double* xT_scratch, const double* x, double* y,
                           double coef_host, const double* coef_dev, const double* z, double* norm_out, double* ws,
                           void* stream) {
  TB200_REQUIRE(nx > 0 && ny > 0 && n_det > 0 && n_ang >= 0 && (int64_t)nx * ny < ((int64_t)1 << 31), "bad geometry");
  const int64_t m = (int64_t)n_ang * n_det;
  int rc = check_sell_args(m, (int64_t)nx * ny, sliceptr, rowlen, colidx, vals, x, y, norm_out, ws);
  if (rc) return rc;
  TB200_REQUIRE(geom != nullptr || m == 0, "null geometry table");
  if (m == 0) return 0;
  CtRays ct;
  rc = ct_rays_setup(nx, ny, n_det, geom, rowskip, cta_order, xT_scratch, x, stream, ct);
  if (rc) return rc;
  if (vals_f32)
    return sell_launch<float>(m, sliceptr, rowlen, colidx, (const float*)vals, x, y, coef_host, coef_dev, z, norm_out, ws,
                              (cudaStream_t)stream, ct);
  return sell_launch<double>(m, sliceptr, rowlen, colidx, (const double*)vals, x, y, coef_host, coef_dev, z, norm_out, ws,
                             (cudaStream_t)stream, ct);
}

// One Golub-Kahan step (the reference's golub_kahan_update, trips/utilities/decompositions.py:230-255) on SELL-32-4
// matrices, enqueued as 6 kernels on `stream` with every scalar kept on the device:
//   v = A^T u_k - beta_prev * v_prev ; alpha = ||v|| ; v /= alpha ; u = A v - alpha * u_k ; beta = ||u|| ; u /= beta
// v_prev / beta_prev_dev are NULL on the first step.  alpha_pair / beta_pair: 2 doubles each (sum of squares, norm).
// ws: tb200_spmv_workspace_len(max(m, n)) doubles.  events_host (nullable): host array of four cudaEvent_t recorded on
// `stream` before/after the A^T launch (+ its norm finalize) and before/after the A launch, for live per-launch timing.
int tb200_vec_div(int64_t n, const double* x, double d_host, const double* d_dev, double* out, void* stream);

int tb200_gk_step_sell_f64(int64_t m, int64_t n, const int64_t* a_sliceptr, const int32_t* a_rowlen, const int32_t* a_col,
                           const double* a_val, const int64_t* at_sliceptr, const int32_t* at_rowlen,
                           const int32_t* at_col, const double* at_val, const double* u_k, const double* v_prev,
                           const double* beta_prev_dev, double* v_out, double* u_out, double* alpha_pair,
                           double* beta_pair, double* ws, void* const* events_host, void* stream) {
  TB200_REQUIRE(u_k && v_out && u_out && alpha_pair && beta_pair && ws, "null pointer");
  TB200_REQUIRE((v_prev == nullptr) == (beta_prev_dev == nullptr), "v_prev and beta_prev_dev go together");
  cudaStream_t st = (cudaStream_t)stream;
  auto mark = [&](int i) {
    if (events_host != nullptr && events_host[i] != nullptr) cudaEventRecord((cudaEvent_t)events_host[i], st);
  };
  mark(0);
  int rc = tb200_spmv_sell_f64(n, m, at_sliceptr, at_rowlen, at_col, at_val, u_k, v_out, 0.0, beta_prev_dev, v_prev,
                               alpha_pair, ws, stream);
  mark(1);
  if (rc) return rc;
  rc = tb200_vec_div(n, v_out, 0.0, alpha_pair + 1, v_out, stream);
  if (rc) return rc;
  mark(2);
  rc = tb200_spmv_sell_f64(m, n, a_sliceptr, a_rowlen, a_col, a_val, v_out, u_out, 0.0, alpha_pair + 1, u_k, beta_pair, ws,
                           stream);
  mark(3);
  if (rc) return rc;
  return tb200_vec_div(m, u_out, 0.0, beta_pair + 1, u_out, stream);
}

// The same Golub-Kahan step with A in the row-aligned CT layout (tb200_ct_spmv_sell_f64) and A^T as a plain SELL matrix.
int tb200_gk_step_sell_ct_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const int64_t* a_sliceptr,
                              const int32_t* a_rowlen, const int32_t* a_rowskip, const int32_t* a_col, const double* a_val,
                              const int32_t* a_cta_order, double* xT_scratch, const int64_t* at_sliceptr,
                              const int32_t* at_rowlen, const int32_t* at_col, const double* at_val, const double* u_k,
                              const double* v_prev, const double* beta_prev_dev, double* v_out, double* u_out,
                              double* alpha_pair, double* beta_pair, double* ws, void* const* events_host, void* stream) {
  TB200_REQUIRE(u_k && v_out && u_out && alpha_pair && beta_pair && ws, "null pointer");
  TB200_REQUIRE((v_prev == nullptr) == (beta_prev_dev == nullptr), "v_prev and beta_prev_dev go together");
  const int64_t m = (int64_t)n_ang * n_det, n = (int64_t)nx * ny;
  cudaStream_t st = (cudaStream_t)stream;
  auto mark = [&](int i) {
    if (events_host != nullptr && events_host[i] != nullptr) cudaEventRecord((cudaEvent_t)events_host[i], st);
  };
  mark(0);
  int rc = tb200_spmv_sell_f64(n, m, at_sliceptr, at_rowlen, at_col, at_val, u_k, v_out, 0.0, beta_prev_dev, v_prev,
                               alpha_pair, ws, stream);
  mark(1);
  if (rc) return rc;
  rc = tb200_vec_div(n, v_out, 0.0, alpha_pair + 1, v_out, stream);
  if (rc) return rc;
  mark(2);
  rc = tb200_ct_spmv_sell_f64(nx, ny, n_det, n_ang, geom, a_sliceptr, a_rowlen, a_rowskip, a_col, a_val, 0, a_cta_order,
                              xT_scratch, v_out, u_out, 0.0, alpha_pair + 1, u_k, beta_pair, ws, stream);
  mark(3);
  if (rc) return rc;
  return tb200_vec_div(m, u_out, 0.0, beta_pair + 1, u_out, stream);
}

// out[0] = round(sum of the n double-double partials (hi, lo interleaved)), out[1] = sqrt(out[0]); one CTA.
int tb200_reduce_finalize(const double* partials, int64_t n, double* out, void* stream) {
  TB200_REQUIRE(partials && out && n >= 0, "bad argument");
  finalize_dd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, n, out);
  return check_launch("reduce_finalize");
}

}  // extern "C"
