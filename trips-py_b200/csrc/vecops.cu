// BLAS-1 style vector kernels used between the operator applies of the Krylov cores.
//
// They mirror, rounding for rounding, the NumPy expressions of the reference:
//   v/alpha, u/beta                      trips/utilities/decompositions.py:238-242
//   x + beta*p, r - beta*w, t + c*p      trips/solvers/CGLS.py:65-72
//   (v**2 + eps**2)**(p/2-1)             trips/solvers/MMGKS.py:57, trips/utilities/weights.py:66-68
//   wf*(AV@y - b), wr*(LV@y)             trips/solvers/MMGKS.py:111-113
// NumPy rounds every elementary operation, so the element-wise kernels use the *_rn intrinsics (no FMA
// contraction).  Reductions (norms, dots) are accumulated in double-double and return the correctly rounded value of
// the exact sum (tb200_dd.cuh): independent of the reduction tree, reproducible bit for bit by the CPU oracle.
#include "tb200_common.cuh"
#include "tb200_dd.cuh"

namespace tb200 {

constexpr int kVecThreads = 256;
constexpr int kMaxReduceBlocks = 1184;  // 148 SMs x 8; fixed so the reduction tree depends on n only

static inline int ew_grid(int64_t n) { return grid_for(n, kVecThreads * 4, 148 * 16); }
static inline int red_grid(int64_t n) { return grid_for(n, kVecThreads * 8, kMaxReduceBlocks); }

__device__ __forceinline__ double scalar_of(double host, const double* dev) { return dev ? *dev : host; }

__global__ void __launch_bounds__(kVecThreads) vec_div_kernel(int64_t n, const double* __restrict__ x, double dh,
                                                              const double* __restrict__ dd, double* __restrict__ out) {
  const double d = scalar_of(dh, dd);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __ddiv_rn(x[i], d);
}

// out = y + sign*(a*x) (a*x rounded first, as NumPy does); optional fused ||out||^2 partials.
__global__ void __launch_bounds__(kVecThreads) vec_axpy_kernel(int64_t n, double ah, const double* __restrict__ ad, double sign,
                                                               const double* __restrict__ x, const double* __restrict__ y,
                                                               double* __restrict__ out, double* __restrict__ partials) {
  __shared__ double red[64];
  const double a = sign * scalar_of(ah, ad);  // sign = +-1: exact
  dd_t acc = dd_zero();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = __dadd_rn(y[i], __dmul_rn(a, x[i]));
    out[i] = v;
    if (partials) acc = dd_fma(acc, v, v);
  }
  if (partials) {
    const dd_t tot = dd_block_sum(acc, red);
    if (threadIdx.x == 0) {
      partials[2 * blockIdx.x] = tot.hi;
      partials[2 * blockIdx.x + 1] = tot.lo;
    }
  }
}

// mode 0: sum x^2 ; 1: sum x*y ; 2: sum (x-y)^2
template <int MODE>
__global__ void __launch_bounds__(kVecThreads) vec_reduce_kernel(int64_t n, const double* __restrict__ x,
                                                                 const double* __restrict__ y, double* __restrict__ partials) {
  __shared__ double red[64];
  dd_t acc = dd_zero();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (MODE == 0) {
      const double v = x[i];
      acc = dd_fma(acc, v, v);
    } else if (MODE == 1) {
      acc = dd_fma(acc, x[i], y[i]);
    } else {
      const double d = __dsub_rn(x[i], y[i]);
      acc = dd_fma(acc, d, d);
    }
  }
  const dd_t tot = dd_block_sum(acc, red);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = tot.hi;
    partials[2 * blockIdx.x + 1] = tot.lo;
  }
}

// mode 0: out = x*y ; 1: out = x-y ; 2: out = w*(x-y) ; 3: out = x + y
template <int MODE>
__global__ void __launch_bounds__(kVecThreads) vec_binary_kernel(int64_t n, const double* __restrict__ x,
                                                                 const double* __restrict__ y, const double* __restrict__ w,
                                                                 double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v;
    if (MODE == 0) v = __dmul_rn(x[i], y[i]);
    else if (MODE == 1) v = __dsub_rn(x[i], y[i]);
    else if (MODE == 2) v = __dmul_rn(w[i], __dsub_rn(x[i], y[i]));
    else v = __dadd_rn(x[i], y[i]);
    out[i] = v;
  }
}

// Smoothed Holder / IRLS weights: out = (v^2 + eps^2)^expo.
__global__ void __launch_bounds__(kVecThreads) irls_weights_kernel(int64_t n, const double* __restrict__ v, double eps2,
                                                                   double expo, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double t = __dadd_rn(__dmul_rn(v[i], v[i]), eps2);
    out[i] = pow(t, expo);
  }
}

}  // namespace tb200

using namespace tb200;

extern "C" {

// Doubles of workspace any reduction in this file needs.
int64_t tb200_reduce_workspace_len(void) { return 2 * kMaxReduceBlocks; }

int tb200_vec_div(int64_t n, const double* x, double d_host, const double* d_dev, double* out, void* stream) {
  TB200_REQUIRE(n >= 0 && (n == 0 || (x && out)), "bad argument");
  if (n == 0) return 0;
  vec_div_kernel<<<ew_grid(n), kVecThreads, 0, (cudaStream_t)stream>>>(n, x, d_host, d_dev, out);
  return check_launch("vec_div");
}

// out = y + sign*a*x (sign = +1 or -1) ; if norm_out != NULL also norm_out[0] = ||out||^2, norm_out[1] = ||out|| (needs ws).
int tb200_vec_axpy(int64_t n, double a_host, const double* a_dev, double sign, const double* x, const double* y, double* out,
                   double* norm_out, double* ws, void* stream) {
  TB200_REQUIRE(n >= 0 && (n == 0 || (x && y && out)), "bad argument");
  TB200_REQUIRE(norm_out == nullptr || ws != nullptr, "norm_out requires a workspace");
  TB200_REQUIRE(sign == 1.0 || sign == -1.0, "sign must be +1 or -1");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = norm_out ? red_grid(n) : ew_grid(n);
  vec_axpy_kernel<<<g, kVecThreads, 0, st>>>(n, a_host, a_dev, sign, x, y, out, norm_out ? ws : nullptr);
  int rc = check_launch("vec_axpy");
  if (rc || !norm_out) return rc;
  finalize_dd_kernel<<<1, 1024, 0, st>>>(ws, g, norm_out);
  return check_launch("vec_axpy finalize");
}

static int reduce_common(int mode, int64_t n, const double* x, const double* y, double* out, double* ws, void* stream) {
  TB200_REQUIRE(n >= 0 && out && ws && (n == 0 || x) && (mode == 0 || n == 0 || y), "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = red_grid(n);
  if (mode == 0) vec_reduce_kernel<0><<<g, kVecThreads, 0, st>>>(n, x, y, ws);
  else if (mode == 1) vec_reduce_kernel<1><<<g, kVecThreads, 0, st>>>(n, x, y, ws);
  else vec_reduce_kernel<2><<<g, kVecThreads, 0, st>>>(n, x, y, ws);
  int rc = check_launch("vec_reduce");
  if (rc) return rc;
  finalize_dd_kernel<<<1, 1024, 0, st>>>(ws, g, out);
  return check_launch("vec_reduce finalize");
}

// The double-double partials of sum x*y (y == NULL: sum x^2) WITHOUT the final reduction: ws receives *n_partials_out
// (hi, lo) pairs, to be summed over the ranks by tb200_comm_allreduce_dd (a vector split over several GPUs).
int tb200_vec_dot_partials(int64_t n, const double* x, const double* y, double* ws, int64_t* n_partials_out, void* stream) {
  TB200_REQUIRE(n >= 0 && ws && n_partials_out && (n == 0 || x), "bad argument");
  const int g = red_grid(n);
  if (y == nullptr) vec_reduce_kernel<0><<<g, kVecThreads, 0, (cudaStream_t)stream>>>(n, x, y, ws);
  else vec_reduce_kernel<1><<<g, kVecThreads, 0, (cudaStream_t)stream>>>(n, x, y, ws);
  *n_partials_out = g;
  return check_launch("vec_dot_partials");
}

// out[0] = sum x^2, out[1] = ||x||
int tb200_vec_norm2(int64_t n, const double* x, double* out, double* ws, void* stream) {
  return reduce_common(0, n, x, nullptr, out, ws, stream);
}
// out[0] = x.y  (out[1] = sqrt of it, NaN when negative: ignore)
int tb200_vec_dot(int64_t n, const double* x, const double* y, double* out, double* ws, void* stream) {
  return reduce_common(1, n, x, y, out, ws, stream);
}
// out[0] = ||x-y||^2, out[1] = ||x-y||
int tb200_vec_diffnorm2(int64_t n, const double* x, const double* y, double* out, double* ws, void* stream) {
  return reduce_common(2, n, x, y, out, ws, stream);
}

// mode 0: out = x*y ; 1: out = x-y ; 2: out = w*(x-y) ; 3: out = x+y
int tb200_vec_binary(int mode, int64_t n, const double* x, const double* y, const double* w, double* out, void* stream) {
  TB200_REQUIRE(n >= 0 && (n == 0 || (x && y && out)) && mode >= 0 && mode <= 3, "bad argument");
  TB200_REQUIRE(mode != 2 || n == 0 || w, "mode 2 needs w");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = ew_grid(n);
  switch (mode) {
    case 0: vec_binary_kernel<0><<<g, kVecThreads, 0, st>>>(n, x, y, w, out); break;
    case 1: vec_binary_kernel<1><<<g, kVecThreads, 0, st>>>(n, x, y, w, out); break;
    case 2: vec_binary_kernel<2><<<g, kVecThreads, 0, st>>>(n, x, y, w, out); break;
    default: vec_binary_kernel<3><<<g, kVecThreads, 0, st>>>(n, x, y, w, out); break;
  }
  return check_launch("vec_binary");
}

// out = (v^2 + eps^2)^expo   (reference: weights.py:66-68 with expo = p/2-1; isoTV uses (q-2)/4)
int tb200_irls_weights(int64_t n, const double* v, double eps, double expo, double* out, void* stream) {
  TB200_REQUIRE(n >= 0 && (n == 0 || (v && out)), "bad argument");
  if (n == 0) return 0;
  irls_weights_kernel<<<ew_grid(n), kVecThreads, 0, (cudaStream_t)stream>>>(n, v, eps * eps, expo, out);
  return check_launch("irls_weights");
}

}  // extern "C"
