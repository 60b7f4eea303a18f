// Stencil operators: 2-D PSF blur (and the reference's "adjoint") and forward-difference regularisation
// operators with the IRLS re-weighting fused in.
//
// PSF blur.  Reference: trips/test_problems/Deblurring2D.py:66-73
//   A x  = scipy.ndimage.convolve(X, PSF, mode='reflect')
//   A^T b = scipy.ndimage.convolve(B, flipud(fliplr(PSF)), mode='reflect')
// ndimage.convolve is a correlation with the flipped kernel (origin moved by one for even sizes) and sums the
// taps in C order with separately rounded multiply and add; the kernel below is that correlation, tap order and
// rounding included, so the result is bit-identical to scipy's.  The host wrapper passes the flipped PSF for A
// and the PSF itself for A^T (for centro-symmetric PSFs - every Gaussian PSF of Deblurring2D.Gauss - this is the
// exact adjoint).  mode 0 = 'reflect' (half-sample symmetric), 1 = 'constant' (zeros; used by gen_data :121-133).
//
// Finite differences.  Reference: trips/utilities/operators.py:24-45 (sparse matrices there)
//   1-D      (L x)_i = x_i - x_{i+1}, i = 0..n-2
//   2-D      L = [ I (x) D ; D (x) I ]  on the row-major image: within-row differences first, then between rows
//   space-time  L = [ I_t (x) L_2D ; D_t (x) I ]  on frame-major x
// The kernels are matrix-free; L @ x can emit the IRLS weights (u^2+eps^2)^expo in the same pass
// (MMGKS.py:60,93) and L.T @ r can apply the weights on the fly, L^T (w . r) (MMGKS.py:113-117).
// The adjoint adds its (up to four / six) terms in the order scipy's CSC scatter does, so it is bit-identical.
#include "tb200_common.cuh"

namespace tb200 {

__device__ __forceinline__ int reflect_index(int i, int n) {
  // half-sample symmetric extension: ... b a | a b c ... y z | z y ...
  while (i < 0 || i >= n) {
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
  }
  return i;
}

constexpr int kTileW = 32, kTileH = 32, kConvTy = 8;

// out[i][j] = sum_{a,b in C order} W[a][b] * x[ext(i + a - ch)][ext(j + b - cw)]
__global__ void __launch_bounds__(kTileW * kConvTy)
correlate2d_kernel(int nrow, int ncol, const double* __restrict__ x, const double* __restrict__ W, int ph, int pw, int ch,
                   int cw, int mode, double* __restrict__ out) {
  extern __shared__ double sm[];
  const int tw = kTileW + pw - 1, th = kTileH + ph - 1;
  double* tile = sm;            // th x tw
  double* wsm = sm + th * tw;   // ph x pw
  const int j0 = blockIdx.x * kTileW, i0 = blockIdx.y * kTileH;
  const int tid = threadIdx.y * kTileW + threadIdx.x, nthr = kTileW * kConvTy;
  for (int t = tid; t < ph * pw; t += nthr) wsm[t] = W[t];
  for (int t = tid; t < th * tw; t += nthr) {
    const int ti = t / tw, tj = t % tw;
    int gi = i0 + ti - ch, gj = j0 + tj - cw;
    double v = 0.0;
    if (mode == 0) {
      gi = reflect_index(gi, nrow);
      gj = reflect_index(gj, ncol);
      v = x[(int64_t)gi * ncol + gj];
    } else if (gi >= 0 && gi < nrow && gj >= 0 && gj < ncol) {
      v = x[(int64_t)gi * ncol + gj];
    }
    tile[t] = v;
  }
  __syncthreads();
  const int j = j0 + threadIdx.x;
  if (j >= ncol) return;
#pragma unroll
  for (int r = 0; r < kTileH / kConvTy; ++r) {
    const int li = threadIdx.y + r * kConvTy;
    const int i = i0 + li;
    if (i >= nrow) continue;
    double acc = 0.0;
    for (int a = 0; a < ph; ++a) {
      const double* trow = tile + (li + a) * tw + threadIdx.x;
      const double* wrow = wsm + a * pw;
      for (int b = 0; b < pw; ++b) acc = __dadd_rn(acc, __dmul_rn(trow[b], wrow[b]));
    }
    out[(int64_t)i * ncol + j] = acc;
  }
}

// ---- finite differences ------------------------------------------------------------------------------
// Space-time operator on nt frames of nrow x ncol (nt = 1 and no temporal part => plain 2-D operator).
// Output layout: [ nt * p2 spatial rows | ntr * N temporal rows ],  p2 = nrow*(ncol-1) + (nrow-1)*ncol, N = nrow*ncol,
// ntr = nt - 1, or nt when x_next (the first frame owned by the next rank) is given.
// (row, col) of a flat index that advances by a fixed stride: one 32-bit division per thread, none per element.
struct RowCol {
  unsigned r, c, step_r, step_c, width;
  __device__ __forceinline__ RowCol(unsigned idx, unsigned stride, unsigned w) : width(w) {
    r = idx / w, c = idx - r * w;
    step_r = stride / w, step_c = stride - step_r * w;
  }
  __device__ __forceinline__ void next() {
    c += step_c, r += step_r;
    if (c >= width) c -= width, ++r;
  }
};

// blockIdx.y = section: frames 0 .. nt-1 are the spatial rows of that frame, nt .. nt+ntr-1 the temporal rows between
// frame (y - nt) and its successor.  All index arithmetic is 32-bit inside a section (nrow*ncol < 2^31).
__global__ void __launch_bounds__(256)
fd_apply_kernel(int nt, int nrow, int ncol, int ntr, const double* __restrict__ x, const double* __restrict__ x_next,
                double* __restrict__ u, double* __restrict__ wout, double eps2, double expo) {
  const unsigned N = (unsigned)nrow * ncol;
  const unsigned p1 = (unsigned)nrow * (ncol - 1), p2 = p1 + (unsigned)(nrow - 1) * ncol;
  const unsigned stride = gridDim.x * blockDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned sec = blockIdx.y;
  auto emit = [&](int64_t q, double a, double b) {
    const double d = __dsub_rn(a, b);
    u[q] = d;
    if (wout != nullptr) wout[q] = pow(__dadd_rn(__dmul_rn(d, d), eps2), expo);
  };
  if (sec < (unsigned)nt) {
    const double* xf = x + (int64_t)sec * N;
    const int64_t qb = (int64_t)sec * p2;
    if (ncol > 1) {  // differences along the row: entry (r, c) of nrow x (ncol-1) = x[r, c] - x[r, c+1]
      RowCol rc(first, stride, (unsigned)(ncol - 1));
      for (unsigned rq = first; rq < p1; rq += stride, rc.next()) {
        const unsigned e = rc.r * (unsigned)ncol + rc.c;
        emit(qb + rq, xf[e], xf[e + 1]);
      }
    }
    for (unsigned e = first; e < p2 - p1; e += stride) emit(qb + p1 + e, xf[e], xf[e + ncol]);  // between rows
  } else {
    const unsigned t = sec - nt;
    const int64_t qb = (int64_t)nt * p2 + (int64_t)t * N;
    const double* xa = x + (int64_t)t * N;
    const double* xb = (t + 1 < (unsigned)nt) ? x + (int64_t)(t + 1) * N : x_next;
    for (unsigned p = first; p < N; p += stride) emit(qb + p, xa[p], xb[p]);
  }
}

__device__ __forceinline__ double wr_at(const double* __restrict__ r, const double* __restrict__ w, int64_t i) {
  return w ? __dmul_rn(w[i], r[i]) : r[i];
}

// out = L^T (w . r).  rt_prev: the temporal rows of the last frame owned by the previous rank (or NULL).
// blockIdx.y = frame.
__global__ void __launch_bounds__(256)
fd_adjoint_kernel(int nt, int nrow, int ncol, int ntr, const double* __restrict__ r, const double* __restrict__ w,
                  const double* __restrict__ rt_prev, const double* __restrict__ wt_prev, double* __restrict__ out) {
  const unsigned N = (unsigned)nrow * ncol;
  const unsigned p1 = (unsigned)nrow * (ncol - 1), p2 = p1 + (unsigned)(nrow - 1) * ncol;
  const int64_t tbase = (int64_t)nt * p2;
  const unsigned stride = gridDim.x * blockDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  const int64_t sb = (int64_t)t * p2;
  RowCol rc(first, stride, (unsigned)ncol);
  for (unsigned p = first; p < N; p += stride, rc.next()) {
    const unsigned row = rc.r, c = rc.c;
    double acc = 0.0;
    // same order as scipy's csc scatter: ascending row index of L
    if (c > 0) acc = __dsub_rn(acc, wr_at(r, w, sb + row * (unsigned)(ncol - 1) + c - 1));
    if (c < (unsigned)ncol - 1) acc = __dadd_rn(acc, wr_at(r, w, sb + row * (unsigned)(ncol - 1) + c));
    if (row > 0) acc = __dsub_rn(acc, wr_at(r, w, sb + p1 + (row - 1) * (unsigned)ncol + c));
    if (row < (unsigned)nrow - 1) acc = __dadd_rn(acc, wr_at(r, w, sb + p1 + row * (unsigned)ncol + c));
    if (t > 0) acc = __dsub_rn(acc, wr_at(r, w, tbase + (int64_t)(t - 1) * N + p));
    else if (rt_prev != nullptr) acc = __dsub_rn(acc, wt_prev ? __dmul_rn(wt_prev[p], rt_prev[p]) : rt_prev[p]);
    if (t < ntr) acc = __dadd_rn(acc, wr_at(r, w, tbase + (int64_t)t * N + p));
    out[(int64_t)t * N + p] = acc;
  }
}

// ---- centred differences (isotropic TV) ------------------------------------------------------------------
// fp64 statement of the reference's first_derivative_operator_2d (trips/utilities/operators_old.py:35-45):
// pylops' 3-point centred FirstDerivative (y[1:-1] = (x[2:] - x[:-2]) / 2, zero at both ends) in
// VStack(Kronecker(I, D), Kronecker(D, I)) form: 2*nrow*ncol rows, [within-row part | between-row part].
// The reference builds it in float32; the isoTV weights of MMGKS.py:64-78,
//   w = (u1^2 + u2^2 + eps^2)^((q-2)/4) for both halves, are fused into the apply pass.
__global__ void __launch_bounds__(256)
cd2d_apply_kernel(int nrow, int ncol, const double* __restrict__ x, double* __restrict__ u, double* __restrict__ wout,
                  double eps2, double expo) {
  const int64_t N = (int64_t)nrow * ncol;
  const unsigned stride = gridDim.x * blockDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;
  RowCol rc(first, stride, (unsigned)ncol);
  for (unsigned q = first; q < (unsigned)N; q += stride, rc.next()) {
    const int r = (int)rc.r, c = (int)rc.c;
    double u1 = 0.0, u2 = 0.0;
    if (c >= 1 && c <= ncol - 2) u1 = __dsub_rn(__dmul_rn(0.5, x[q + 1]), __dmul_rn(0.5, x[q - 1]));
    if (r >= 1 && r <= nrow - 2) u2 = __dsub_rn(__dmul_rn(0.5, x[q + ncol]), __dmul_rn(0.5, x[q - ncol]));
    if (u != nullptr) {
      u[q] = u1;
      u[N + q] = u2;
    }
    if (wout != nullptr) {
      const double t = __dadd_rn(__dadd_rn(__dmul_rn(u1, u1), __dmul_rn(u2, u2)), eps2);
      const double w = pow(t, expo);
      wout[q] = w;
      wout[N + q] = w;
    }
  }
}

// out = L^T (w . r), terms added in the order of scipy's CSC scatter (ascending row index of L)
__global__ void __launch_bounds__(256)
cd2d_adjoint_kernel(int nrow, int ncol, const double* __restrict__ r, const double* __restrict__ w,
                    double* __restrict__ out) {
  const int64_t N = (int64_t)nrow * ncol;
  const unsigned stride = gridDim.x * blockDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;
  RowCol rc(first, stride, (unsigned)ncol);
  for (unsigned q = first; q < (unsigned)N; q += stride, rc.next()) {
    const int row = (int)rc.r, c = (int)rc.c;
    double acc = 0.0;
    // within-row part: L-row (row, c') is non-zero for 1 <= c' <= ncol-2 with -0.5 at c'-1 and +0.5 at c'+1
    if (c - 1 >= 1 && c - 1 <= ncol - 2) acc = __dadd_rn(acc, __dmul_rn(0.5, wr_at(r, w, q - 1)));
    if (c + 1 >= 1 && c + 1 <= ncol - 2) acc = __dsub_rn(acc, __dmul_rn(0.5, wr_at(r, w, q + 1)));
    if (row - 1 >= 1 && row - 1 <= nrow - 2) acc = __dadd_rn(acc, __dmul_rn(0.5, wr_at(r, w, N + q - ncol)));
    if (row + 1 >= 1 && row + 1 <= nrow - 2) acc = __dsub_rn(acc, __dmul_rn(0.5, wr_at(r, w, N + q + ncol)));
    out[q] = acc;
  }
}

// 1-D operator (n-1) x n and its adjoint
__global__ void __launch_bounds__(256) fd1d_apply_kernel(int64_t n, const double* __restrict__ x, double* __restrict__ u) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += (int64_t)gridDim.x * blockDim.x)
    u[i] = __dsub_rn(x[i], x[i + 1]);
}
__global__ void __launch_bounds__(256) fd1d_adjoint_kernel(int64_t n, const double* __restrict__ r, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    if (i > 0) acc = __dsub_rn(acc, r[i - 1]);
    if (i < n - 1) acc = __dadd_rn(acc, r[i]);
    out[i] = acc;
  }
}

}  // namespace tb200

using namespace tb200;

extern "C" {

// out = correlate(x, W) with ndimage tap order; ch/cw = index of the tap that sits on the output pixel.
// mode 0 reflect, 1 constant-zero.
int tb200_correlate2d_f64(int nrow, int ncol, const double* x, const double* W, int ph, int pw, int ch, int cw, int mode,
                          double* out, void* stream) {
  TB200_REQUIRE(nrow > 0 && ncol > 0 && ph > 0 && pw > 0 && x && W && out, "bad argument");
  TB200_REQUIRE(ch >= 0 && ch < ph && cw >= 0 && cw < pw && (mode == 0 || mode == 1), "bad centre or mode");
  TB200_REQUIRE(x != out, "in-place convolution is not supported");
  const size_t smem = ((size_t)(kTileW + pw - 1) * (kTileH + ph - 1) + (size_t)ph * pw) * sizeof(double);
  TB200_REQUIRE(smem <= 200 * 1024, "PSF too large for the shared-memory tile");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(correlate2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("correlate2d: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  dim3 grid((ncol + kTileW - 1) / kTileW, (nrow + kTileH - 1) / kTileH), block(kTileW, kConvTy);
  correlate2d_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(nrow, ncol, x, W, ph, pw, ch, cw, mode, out);
  return check_launch("correlate2d");
}

// Number of rows of the space-time operator for nt local frames (temporal rows: nt-1, or nt with a halo frame).
int64_t tb200_fd_rows(int nt, int nrow, int ncol, int has_next) {
  const int64_t p2 = (int64_t)nrow * (ncol - 1) + (int64_t)(nrow - 1) * ncol;
  return (int64_t)nt * p2 + (int64_t)(has_next ? nt : nt - 1) * nrow * ncol;
}

// u = L x (space-time forward differences; nt = 1 => the 2-D operator). Optional fused weights
// wout = (u^2 + eps^2)^expo.  x_next: first frame of the next rank (device pointer) or NULL.
int tb200_fd_apply(int nt, int nrow, int ncol, const double* x, const double* x_next, double* u, double* wout, double eps,
                   double expo, void* stream) {
  TB200_REQUIRE(nt >= 1 && nrow >= 1 && ncol >= 1 && x && u, "bad argument");
  TB200_REQUIRE((int64_t)nrow * ncol < ((int64_t)1 << 31), "frame too large for 32-bit indexing");
  const int ntr = x_next ? nt : nt - 1;
  const int64_t total = tb200_fd_rows(nt, nrow, ncol, x_next != nullptr);
  if (total == 0) return 0;
  TB200_REQUIRE(nt + ntr <= 65535, "too many frames for this launch shape");
  const int secs = nt + ntr;
  const dim3 grid((unsigned)grid_for((int64_t)nrow * ncol, 256 * 4, secs >= 8 ? 148 * 2 : 148 * 16), (unsigned)secs);
  fd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(nt, nrow, ncol, ntr, x, x_next, u, wout, eps * eps, expo);
  return check_launch("fd_apply");
}

// out = L^T (w . r)  (w may be NULL).  has_next selects the row layout used by tb200_fd_apply;
// rt_prev / wt_prev: temporal rows (and their weights) of the previous rank's last frame, or NULL.
int tb200_fd_adjoint(int nt, int nrow, int ncol, int has_next, const double* r, const double* w, const double* rt_prev,
                     const double* wt_prev, double* out, void* stream) {
  TB200_REQUIRE(nt >= 1 && nrow >= 1 && ncol >= 1 && r && out, "bad argument");
  TB200_REQUIRE((int64_t)nrow * ncol < ((int64_t)1 << 31) && nt <= 65535, "frame too large / too many frames");
  const int ntr = has_next ? nt : nt - 1;
  const dim3 grid((unsigned)grid_for((int64_t)nrow * ncol, 256 * 4, nt >= 8 ? 148 * 2 : 148 * 16), (unsigned)nt);
  fd_adjoint_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(nt, nrow, ncol, ntr, r, w, rt_prev, wt_prev, out);
  return check_launch("fd_adjoint");
}

// Centred-difference gradient (2*nrow*ncol rows): u = L x (u may be NULL) and, if wout != NULL, the isotropic-TV
// weights wout = (u1^2 + u2^2 + eps^2)^expo written to both halves (MMGKS.py:64-78 with nt = 1).
int tb200_cd2d_apply(int nrow, int ncol, const double* x, double* u, double* wout, double eps, double expo, void* stream) {
  TB200_REQUIRE(nrow >= 1 && ncol >= 1 && x && (u || wout), "bad argument");
  TB200_REQUIRE((int64_t)nrow * ncol < ((int64_t)1 << 31), "image too large for 32-bit indexing");
  const int64_t N = (int64_t)nrow * ncol;
  cd2d_apply_kernel<<<grid_for(N, 256 * 4, 148 * 16), 256, 0, (cudaStream_t)stream>>>(nrow, ncol, x, u, wout, eps * eps, expo);
  return check_launch("cd2d_apply");
}
// out = L^T (w . r) for the centred-difference gradient (w may be NULL).
int tb200_cd2d_adjoint(int nrow, int ncol, const double* r, const double* w, double* out, void* stream) {
  TB200_REQUIRE(nrow >= 1 && ncol >= 1 && r && out, "bad argument");
  const int64_t N = (int64_t)nrow * ncol;
  cd2d_adjoint_kernel<<<grid_for(N, 256 * 4, 148 * 16), 256, 0, (cudaStream_t)stream>>>(nrow, ncol, r, w, out);
  return check_launch("cd2d_adjoint");
}

int tb200_fd1d_apply(int64_t n, const double* x, double* u, void* stream) {
  TB200_REQUIRE(n >= 1 && x && u, "bad argument");
  if (n == 1) return 0;
  fd1d_apply_kernel<<<grid_for(n, 256 * 4, 148 * 16), 256, 0, (cudaStream_t)stream>>>(n, x, u);
  return check_launch("fd1d_apply");
}
int tb200_fd1d_adjoint(int64_t n, const double* r, double* out, void* stream) {
  TB200_REQUIRE(n >= 1 && r && out, "bad argument");
  fd1d_adjoint_kernel<<<grid_for(n, 256 * 4, 148 * 16), 256, 0, (cudaStream_t)stream>>>(n, r, out);
  return check_launch("fd1d_adjoint");
}

}  // extern "C"
