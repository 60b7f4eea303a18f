// Measured fp64 instruction-issue peak of the device (the roofline denominator of the matrix-free CT projectors).
//
// The projectors of ct_project.cu / ct_forward.cu read (almost) no DRAM: they re-evaluate every matrix entry with
// separately rounded DADD / DMUL / DSETP instructions, so what bounds them is the rate at which an SM retires fp64
// instructions, not HBM.  MEASURED_PEAKS.json carries no such number, so it is measured here, live, by bench.py:
// every thread runs 8 independent DFMA chains (enough ILP to cover the pipe latency at 32 warps/SM), no memory
// traffic; rate = thread-level fp64 instructions / elapsed time.  DADD, DMUL, DFMA and DSETP occupy the same pipe
// slot, so "instructions" is the unit that transfers to the projector kernels (a DFMA counted as ONE instruction).
#include "tb200_common.cuh"

namespace tb200 {

constexpr int PEAK_ILP = 8;
constexpr int PEAK_THREADS = 256;
constexpr int PEAK_CTAS_PER_SM = 8;

__global__ void __launch_bounds__(PEAK_THREADS)
fp64_peak_kernel(int iters, double seed, double* __restrict__ sink) {
  double a[PEAK_ILP];
#pragma unroll
  for (int k = 0; k < PEAK_ILP; ++k) a[k] = seed + (double)(threadIdx.x + k);
  const double m = 1.0000000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < PEAK_ILP; ++k) a[k] = __fma_rn(a[k], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < PEAK_ILP; ++k) s += a[k];
  if (s == 12345.678) sink[0] = s;  // never true: keeps the chains alive
}

}  // namespace tb200

extern "C" {

// Thread-level fp64 instructions one tb200_fp64_peak_run(iters) launch executes (for rate = this / elapsed time).
int64_t tb200_fp64_peak_instructions(int iters) {
  return (int64_t)tb200::sm_count() * tb200::PEAK_CTAS_PER_SM * tb200::PEAK_THREADS * tb200::PEAK_ILP * (int64_t)iters;
}

// One launch of the DFMA-chain microbenchmark on `stream` (sink: one device double, never written in practice).
int tb200_fp64_peak_run(int iters, double* sink, void* stream) {
  TB200_REQUIRE(iters > 0 && sink != nullptr, "bad arguments");
  tb200::fp64_peak_kernel<<<tb200::sm_count() * tb200::PEAK_CTAS_PER_SM, tb200::PEAK_THREADS, 0, (cudaStream_t)stream>>>(
      iters, 1.0, sink);
  return tb200::check_launch("fp64_peak");
}

}  // extern "C"
