// Shared device/host helpers for libtripsb200 (sm_100a only).
//
// Conventions used by every translation unit:
//  * every extern "C" entry point returns int: 0 = ok, otherwise a cudaError_t value (1..999) or one of the
//    TB200_E* codes below; the message is kept thread-local and read with tb200_last_error().
//  * nothing here allocates device memory or keeps a pointer after returning; all launches are asynchronous
//    on the caller's stream.
//  * element-wise fp64 arithmetic that mirrors a NumPy expression of the reference uses the explicit
//    round-to-nearest intrinsics (__dmul_rn/__dadd_rn/...) so that nvcc cannot contract a*b+c into one FMA:
//    NumPy evaluates `y + a*x` with two roundings and the 1e-10 parity gate is a statement about rounding.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define TB200_EINVAL 1001    // bad argument (null pointer, negative size, misaligned buffer)
#define TB200_ENOTSM100 1002 // not running on a compute-capability 10.x device

namespace tb200 {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

constexpr int kWarp = 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Fixed-shape block reduction (blockDim.x a multiple of 32, <= 1024). Result valid in thread 0.
// The tree is the same for every launch of the same configuration => run-to-run bitwise reproducible.
__device__ __forceinline__ double block_sum(double v, double* smem /* >= 32 doubles */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect smem reuse across consecutive calls
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.0;
  if (wid == 0) v = warp_sum(v);
  return v;
}

// ---- cache-policy loads -------------------------------------------------------------------------------
// The matrix stream (values + column indices) is read exactly once per SpMV: keep it out of L1 and mark it
// evict-first in L2 so it does not displace the gathered vector, which is the only data with reuse.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// 256-bit streaming load of four doubles (sm_100+: LDG.E.NA.EFL2.256). p must be 32-byte aligned.
__device__ __forceinline__ void ld_stream_f64x4(const double* p, double (&v)[4]) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
               : "l"(p));
}
// 128-bit streaming load of four floats (fp32-storage variant).
__device__ __forceinline__ void ld_stream_f32x4(const float* p, uint64_t pol, float (&v)[4]) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
               : "l"(p), "l"(pol));
}
// 128-bit streaming load of four int32 column indices. p must be 16-byte aligned.
__device__ __forceinline__ void ld_stream_i32x4(const int32_t* p, uint64_t pol, int32_t (&c)[4]) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
               : "l"(p), "l"(pol));
}
// Gather of the dense vector: read-only path, allocate in L1, prefer to stay in L2.
__device__ __forceinline__ double ld_gather_f64(const double* p, uint64_t pol) {
  double r;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
  return r;
}


// Single-CTA fixed-order reduction of per-CTA partials: out[0] = sum, out[1] = sqrt(sum).
// (static: every translation unit gets its own copy, so the library needs no relocatable device code.)
static __global__ void __launch_bounds__(1024) finalize_sum_kernel(const double* __restrict__ partials, int64_t n,
                                                                   double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += partials[i];
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) {
    out[0] = tot;
    out[1] = sqrt(tot);
  }
}

inline int grid_for(int64_t n, int per_block, int max_blocks) {
  int64_t g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

int sm_count();

// Extra destinations of a kernel's output vector: the same element is also stored at p[i][index] for i < n - device
// pointers into the arenas of other GPUs (csrc/comm.cu), written over NVLink from the kernel epilogue.
struct PeerOut {
  int n;
  double* p[15];
};

}  // namespace tb200

#define TB200_REQUIRE(cond, msg)                               \
  do {                                                         \
    if (!(cond)) {                                             \
      tb200::set_error("%s: %s", __func__, msg);               \
      return TB200_EINVAL;                                     \
    }                                                          \
  } while (0)
