// Intra-node communication over NVLink / NVSwitch peer memory, without NCCL on the data path (SURVEY.md 8b:
// tb200_comm_init/_destroy, tb200_allreduce_*, tb200_halo_exchange; 8e/8f-3: the exchange fused into the kernels).
// The reference has no parallelism of any kind: this layer is new.
//
// One process per GPU.  Every rank cudaMalloc's one ARENA of the same size and publishes its CUDA IPC handle; after
// tb200_comm_connect every rank holds a device pointer to every peer's arena, so kernels can store straight into a
// peer's memory (the projector epilogues of ct_project.cu / ct_forward.cu push their results to all ranks while the
// rest of the grid is still computing).  The caller (trips-py_b200/dist.py) lays the arena out; this file only knows
// three things about it: where the MAILBOXES are (the first TB200_COMM_MAILBOX_BYTES), how to sum a double-double over
// the ranks through them, and how to copy a block into the same offset of some / all peers.
//
// Deterministic all-reduce of double-double scalars (tb200_comm_allreduce_dd): rank r reduces its per-CTA partials,
// stores (hi, lo) into slot [box][r] of EVERY rank's mailbox, fences (system scope) and stores the epoch next to it;
// then it polls its own mailbox until all ranks' epochs have arrived and adds the totals IN RANK ORDER in
// double-double.  Every rank therefore computes bit-identical sums, and the same sum as a single GPU would (the
// correctly rounded exact total).  The exchange doubles as the inter-GPU barrier that orders the pushed vectors: a
// rank's pushes are issued (and fenced, system scope) before its epoch store, so whoever has seen the epoch sees them.
// Epochs grow monotonically (the caller passes them), nothing is ever reset.
#include "tb200_common.cuh"
#include "tb200_dd.cuh"

#include <cstring>

#define TB200_COMM_MAX_RANKS 16
#define TB200_COMM_BOXES 8          // independent mailboxes (alpha, beta, halo flags, ...)
#define TB200_COMM_VALUES 64        // double-double values per all-reduce
#define TB200_ECOMM 1003

namespace tb200 {

// mailbox[box][rank]: TB200_COMM_VALUES (hi, lo) pairs + the epoch; 16-byte aligned entries
struct MailSlot {
  double hilo[2 * TB200_COMM_VALUES];
  unsigned long long epoch;
  unsigned long long pad;
};
constexpr size_t kMailboxBytes = sizeof(MailSlot) * TB200_COMM_BOXES * TB200_COMM_MAX_RANKS;

struct Comm {
  int rank, nranks, device;
  size_t bytes;
  unsigned char* base[TB200_COMM_MAX_RANKS];  // base[rank] = this rank's arena; others = IPC mappings
  bool opened[TB200_COMM_MAX_RANKS];
};

struct PeerTable {
  unsigned char* base[TB200_COMM_MAX_RANKS];
};

__device__ __forceinline__ MailSlot* mail(unsigned char* arena, int box, int rank) {
  return reinterpret_cast<MailSlot*>(arena) + (box * TB200_COMM_MAX_RANKS + rank);
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// One CTA.  partials: npart x nval (hi, lo) pairs, value-major inside a partial: partials[(p*nval + j)*2 + {0,1}]
// (npart == 0: the value is taken as zero - a pure barrier).  out[j*2] = total, out[j*2+1] = sqrt(total).
__global__ void __launch_bounds__(1024)
comm_allreduce_dd_kernel(PeerTable pt, int rank, int nranks, int box, unsigned long long epoch, const double* __restrict__ partials,
                         int64_t npart, int nval, double* __restrict__ out) {
  __shared__ double red[64];
  __shared__ double mine[2 * TB200_COMM_VALUES];
  __shared__ int timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  for (int j = 0; j < nval; ++j) {
    dd_t acc = dd_zero();
    for (int64_t i = threadIdx.x; i < npart; i += blockDim.x)
      acc = dd_add(acc, dd_t{partials[(i * nval + j) * 2], partials[(i * nval + j) * 2 + 1]});
    const dd_t tot = dd_block_sum(acc, red);
    if (threadIdx.x == 0) mine[2 * j] = tot.hi, mine[2 * j + 1] = tot.lo;
    __syncthreads();
  }
  // thread r: deliver my totals to rank r's mailbox, then wait for rank r's totals in mine
  if ((int)threadIdx.x < nranks) {
    const int r = threadIdx.x;
    MailSlot* dst = mail(pt.base[r], box, rank);
    for (int j = 0; j < 2 * nval; ++j) dst->hilo[j] = mine[j];
    __threadfence_system();  // my values (and everything this GPU stored before: kernel order + this fence) before the epoch
    st_release_sys(&dst->epoch, epoch);
    const MailSlot* src = mail(pt.base[rank], box, r);
    // a peer that never arrives (crashed process) must not hang the GPU: give up after ~4 s and poison the result
    const long long t0 = clock64();
    while (ld_acquire_sys(&src->epoch) < epoch) {
      if (clock64() - t0 > 8000000000LL) {
        timed_out = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < (unsigned)nval) {
    const int j = threadIdx.x;
    dd_t tot = dd_zero();
    if (timed_out) tot.hi = __longlong_as_double(0x7ff8000000000000LL);
    for (int r = 0; r < nranks; ++r) {  // rank order: the same sum on every rank
      const MailSlot* src = mail(pt.base[rank], box, r);
      const volatile double* v = src->hilo;
      tot = dd_add(tot, dd_t{v[2 * j], v[2 * j + 1]});
    }
    out[2 * j] = tot.hi;
    out[2 * j + 1] = sqrt(tot.hi);
  }
}

// dst_r[offset + i] = src[i] for every rank r in the mask (the same arena offset everywhere), 128-bit stores
__global__ void __launch_bounds__(256)
comm_push_kernel(PeerTable pt, int nranks, unsigned mask, int64_t offset_doubles, const double* __restrict__ src, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n2 = n >> 1;
  const bool vec = ((offset_doubles & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (int r = 0; r < nranks; ++r) {
    if (!((mask >> r) & 1u)) continue;
    double* dst = reinterpret_cast<double*>(pt.base[r]) + offset_doubles;
    if (vec) {
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride)
        reinterpret_cast<double2*>(dst)[i] = reinterpret_cast<const double2*>(src)[i];
      if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[n - 1] = src[n - 1];
    } else {
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
    }
  }
  __threadfence_system();
}

// out[i] = x[i] / d (d from the device), the normalisation pass after an exchange; optionally also copies the slice
// [keep_begin, keep_begin + keep_n) of the result into `keep` (the rank's own part, which goes into its basis)
__global__ void __launch_bounds__(256)
comm_scale_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ d_dev, double* __restrict__ out,
                  int64_t keep_begin, int64_t keep_n, double* __restrict__ keep) {
  const double d = *d_dev;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = __ddiv_rn(x[i], d);
    out[i] = v;
    const int64_t k = i - keep_begin;
    if (keep != nullptr && k >= 0 && k < keep_n) keep[k] = v;
  }
}

// tb200_comm_allreduce_dd and tb200_comm_scale in ONE launch (nval = 1): CTA 0 reduces the launch partials and posts the
// rank's total to every mailbox; every CTA waits (one polling thread per rank) until all ranks' totals have arrived,
// adds them in rank order - the same double-double sum in every CTA of every rank - and divides its share of x by the
// square root.  Saves a launch and the serial 1-CTA kernel between a projector and the normalisation of its result.
__global__ void __launch_bounds__(256)
comm_allreduce_scale_kernel(PeerTable pt, int rank, int nranks, int box, unsigned long long epoch,
                            const double* __restrict__ partials, int64_t npart, int64_t n, const double* __restrict__ x,
                            double* __restrict__ out, int64_t keep_begin, int64_t keep_n, double* __restrict__ keep,
                            double* __restrict__ pair_out) {
  __shared__ double red[64];
  __shared__ double tot_s[2];
  __shared__ int timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  if (blockIdx.x == 0) {
    dd_t acc = dd_zero();
    for (int64_t i = threadIdx.x; i < npart; i += blockDim.x) acc = dd_add(acc, dd_t{partials[2 * i], partials[2 * i + 1]});
    const dd_t tot = dd_block_sum(acc, red);
    if (threadIdx.x == 0) red[0] = tot.hi, red[1] = tot.lo;
    __syncthreads();
    if ((int)threadIdx.x < nranks) {
      MailSlot* dst = mail(pt.base[threadIdx.x], box, rank);
      dst->hilo[0] = red[0];
      dst->hilo[1] = red[1];
      __threadfence_system();
      st_release_sys(&dst->epoch, epoch);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < nranks) {
    const MailSlot* src = mail(pt.base[rank], box, threadIdx.x);
    const long long t0 = clock64();
    while (ld_acquire_sys(&src->epoch) < epoch) {
      if (clock64() - t0 > 8000000000LL) {
        timed_out = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    dd_t tot = dd_zero();
    for (int r = 0; r < nranks; ++r) {
      const volatile double* v = mail(pt.base[rank], box, r)->hilo;
      tot = dd_add(tot, dd_t{v[0], v[1]});
    }
    if (timed_out) tot.hi = __longlong_as_double(0x7ff8000000000000LL);
    tot_s[0] = tot.hi;
    tot_s[1] = sqrt(tot.hi);
    if (blockIdx.x == 0) pair_out[0] = tot_s[0], pair_out[1] = tot_s[1];
  }
  __syncthreads();
  const double d = tot_s[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = __ddiv_rn(x[i], d);
    out[i] = v;
    const int64_t k = i - keep_begin;
    if (keep != nullptr && k >= 0 && k < keep_n) keep[k] = v;
  }
}

static PeerTable table_of(const Comm* c) {
  PeerTable pt;
  for (int r = 0; r < TB200_COMM_MAX_RANKS; ++r) pt.base[r] = (r < c->nranks) ? c->base[r] : nullptr;
  return pt;
}

}  // namespace tb200

using namespace tb200;

#define TB200_CUDA(call, what)                                                  \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) {                                                   \
      tb200::set_error("%s: %s", what, cudaGetErrorString(e__));                \
      return (int)e__;                                                          \
    }                                                                           \
  } while (0)

extern "C" {

int64_t tb200_comm_mailbox_bytes(void) { return (int64_t)((kMailboxBytes + 255) / 256 * 256); }
int tb200_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }
int tb200_comm_max_ranks(void) { return TB200_COMM_MAX_RANKS; }

// Allocates this rank's arena (arena_bytes >= tb200_comm_mailbox_bytes(), zero-filled: all epochs start at 0) on the
// current device and writes its CUDA IPC handle (tb200_comm_handle_bytes() bytes) to handle_out.  The caller exchanges
// the handles (any out-of-band channel: torch.distributed all_gather, MPI, files) and calls tb200_comm_connect.
int tb200_comm_init(int rank, int nranks, int64_t arena_bytes, void** comm_out, unsigned char* handle_out) {
  TB200_REQUIRE(comm_out && handle_out, "null pointer");
  TB200_REQUIRE(nranks >= 1 && nranks <= TB200_COMM_MAX_RANKS && rank >= 0 && rank < nranks, "bad rank / nranks");
  TB200_REQUIRE(arena_bytes >= (int64_t)kMailboxBytes, "arena smaller than the mailboxes");
  Comm* c = new Comm();
  std::memset(c, 0, sizeof(Comm));
  c->rank = rank, c->nranks = nranks, c->bytes = (size_t)arena_bytes;
  TB200_CUDA(cudaGetDevice(&c->device), "cudaGetDevice");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, c->bytes);
  if (e != cudaSuccess) {
    delete c;
    tb200::set_error("cudaMalloc(arena): %s", cudaGetErrorString(e));
    return (int)e;
  }
  cudaMemset(p, 0, c->bytes);
  cudaDeviceSynchronize();
  c->base[rank] = (unsigned char*)p;
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    delete c;
    tb200::set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    return (int)e;
  }
  std::memcpy(handle_out, &h, sizeof(h));
  *comm_out = c;
  return 0;
}

// all_handles: nranks handles in rank order (the own one is ignored).  Maps every peer's arena into this process.
int tb200_comm_connect(void* comm, const unsigned char* all_handles) {
  Comm* c = (Comm*)comm;
  TB200_REQUIRE(c && all_handles, "null pointer");
  for (int r = 0; r < c->nranks; ++r) {
    if (r == c->rank || c->opened[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, all_handles + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      tb200::set_error("cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
      return (int)e;
    }
    c->base[r] = (unsigned char*)p;
    c->opened[r] = true;
  }
  return 0;
}

// Device address (valid in THIS process) of rank `peer`'s arena.
void* tb200_comm_arena(void* comm, int peer) {
  Comm* c = (Comm*)comm;
  if (!c || peer < 0 || peer >= c->nranks) return nullptr;
  return c->base[peer];
}

int tb200_comm_destroy(void* comm) {
  Comm* c = (Comm*)comm;
  if (!c) return 0;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->nranks; ++r)
    if (r != c->rank && c->opened[r]) cudaIpcCloseMemHandle(c->base[r]);
  if (c->base[c->rank]) cudaFree(c->base[c->rank]);
  delete c;
  return 0;
}

// Sum over the ranks of `nval` double-double values (<= 64), each given as `npart` local partials (hi, lo), through
// mailbox `box` with a caller-chosen, strictly increasing `epoch` (> 0).  out[2j] = total_j, out[2j+1] = sqrt(total_j)
// on every rank, bit-identical everywhere; also an inter-GPU barrier (see the header).  One kernel, no NCCL.
// Replaces the norm of decompositions.py:238,241 over a vector that is split over the GPUs.
int tb200_comm_allreduce_dd(void* comm, int box, int64_t epoch, const double* partials, int64_t npart, int nval, double* out,
                            void* stream) {
  Comm* c = (Comm*)comm;
  TB200_REQUIRE(c && out, "null pointer");
  TB200_REQUIRE(box >= 0 && box < TB200_COMM_BOXES && epoch > 0, "bad mailbox / epoch");
  TB200_REQUIRE(nval >= 1 && nval <= TB200_COMM_VALUES && npart >= 0 && (npart == 0 || partials), "bad partials");
  comm_allreduce_dd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(table_of(c), c->rank, c->nranks, box, (unsigned long long)epoch,
                                                                 partials, npart, nval, out);
  return check_launch("comm_allreduce_dd");
}

// tb200_comm_allreduce_dd (one value) followed by tb200_comm_scale with its square root, in one launch:
// pair_out = (total, sqrt(total)), out = x / sqrt(total), keep[0:keep_n) = out[keep_begin : keep_begin + keep_n).
// Every CTA waits for the mailboxes itself, so the grid is limited to what is resident at once.
int tb200_comm_allreduce_scale(void* comm, int box, int64_t epoch, const double* partials, int64_t npart, int64_t n,
                               const double* x, double* out, int64_t keep_begin, int64_t keep_n, double* keep, double* pair_out,
                               void* stream) {
  Comm* c = (Comm*)comm;
  TB200_REQUIRE(c && pair_out && n >= 0 && (n == 0 || (x && out)), "null pointer");
  TB200_REQUIRE(box >= 0 && box < TB200_COMM_BOXES && epoch > 0, "bad mailbox / epoch");
  TB200_REQUIRE(npart >= 0 && (npart == 0 || partials), "bad partials");
  TB200_REQUIRE(keep == nullptr || (keep_begin >= 0 && keep_n >= 0 && keep_begin + keep_n <= n), "bad keep slice");
  const int grid = grid_for(n > 0 ? n : 1, 256 * 8, sm_count() * 4);  // <= 4 CTAs of 256 threads per SM: all resident
  comm_allreduce_scale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table_of(c), c->rank, c->nranks, box,
                                                                      (unsigned long long)epoch, partials, npart, n, x, out,
                                                                      keep_begin, keep_n, keep, pair_out);
  return check_launch("comm_allreduce_scale");
}

// Copies src[0:n) to offset_bytes of the arena of every rank in rank_mask (bit r = rank r; the own arena included if
// its bit is set).  Visibility at the receivers is established by the next tb200_comm_allreduce_dd on this stream.
int tb200_comm_push(void* comm, unsigned rank_mask, int64_t offset_bytes, const double* src, int64_t n, void* stream) {
  Comm* c = (Comm*)comm;
  TB200_REQUIRE(c && (src || n == 0), "null pointer");
  TB200_REQUIRE(offset_bytes >= (int64_t)kMailboxBytes && offset_bytes % 8 == 0 && offset_bytes + 8 * n <= (int64_t)c->bytes,
                "push outside the arena");
  if (n == 0) return 0;
  comm_push_kernel<<<grid_for(n, 256 * 8, 148 * 4), 256, 0, (cudaStream_t)stream>>>(table_of(c), c->nranks, rank_mask,
                                                                                  offset_bytes / 8, src, n);
  return check_launch("comm_push");
}

// One-frame halo exchange for a frame-sharded operator (dynamic CT): send_next / send_prev (nullable,
// n doubles each) are stored into the halo area (4n doubles at offset_bytes of every arena: two alternating buffers) of
// rank+1 / rank-1; after the call's epoch has been exchanged through mailbox `box`, recv_prev / recv_next (nullable
// outputs) hold what the neighbours sent.  Replaces the torch.distributed isend/irecv pair of dist.FrameComm.
int tb200_halo_exchange(void* comm, int box, int64_t epoch, int64_t offset_bytes, const double* send_prev, const double* send_next,
                        int64_t n, double* recv_prev, double* recv_next, double* scratch_pair, void* stream) {
  Comm* c = (Comm*)comm;
  TB200_REQUIRE(c && scratch_pair, "null pointer");
  int rc;
  // area layout at offset_bytes: two buffers of 2n doubles, used alternately (epoch parity) so that a neighbour's stores
  // of the NEXT exchange never land on data this rank has not copied out yet; in a buffer [0, n) = block received from
  // the previous rank, [n, 2n) = block received from the next
  offset_bytes += (epoch & 1) ? 16 * n : 0;
  if (send_next && c->rank + 1 < c->nranks) {
    rc = tb200_comm_push(comm, 1u << (c->rank + 1), offset_bytes, send_next, n, stream);
    if (rc) return rc;
  }
  if (send_prev && c->rank > 0) {
    rc = tb200_comm_push(comm, 1u << (c->rank - 1), offset_bytes + 8 * n, send_prev, n, stream);
    if (rc) return rc;
  }
  rc = tb200_comm_allreduce_dd(comm, box, epoch, nullptr, 0, 1, scratch_pair, stream);
  if (rc) return rc;
  const double* area = reinterpret_cast<const double*>(c->base[c->rank] + offset_bytes);
  if (recv_prev && c->rank > 0)
    TB200_CUDA(cudaMemcpyAsync(recv_prev, area, 8 * n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream), "halo copy");
  if (recv_next && c->rank + 1 < c->nranks)
    TB200_CUDA(cudaMemcpyAsync(recv_next, area + n, 8 * n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream), "halo copy");
  return 0;
}

// out = x / d (d on the device) over n elements and, optionally, keep[0:keep_n) = out[keep_begin : keep_begin + keep_n):
// the replicated normalisation after an exchanged vector's norm is known (v / alpha, u / beta of
// decompositions.py:239,242), writing the rank's own slice into its basis column in the same pass.
int tb200_comm_scale(int64_t n, const double* x, const double* d_dev, double* out, int64_t keep_begin, int64_t keep_n,
                     double* keep, void* stream) {
  TB200_REQUIRE(n >= 0 && (n == 0 || (x && d_dev && out)), "null pointer");
  TB200_REQUIRE(keep == nullptr || (keep_begin >= 0 && keep_n >= 0 && keep_begin + keep_n <= n), "bad keep slice");
  if (n == 0) return 0;
  comm_scale_kernel<<<grid_for(n, 256 * 4, 148 * 8), 256, 0, (cudaStream_t)stream>>>(n, x, d_dev, out, keep_begin, keep_n, keep);
  return check_launch("comm_scale");
}

}  // extern "C"
