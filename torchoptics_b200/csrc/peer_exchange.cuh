// Peer-memory exchange: the one data-path collective of the sharded spot pass (a SUM
// all-reduce of the [B,F,W,n] fp64 moment sums, 27 KB for the config-2 lens) done by ONE small
// kernel over NVLink / NVSwitch peer memory instead of an NCCL ring/tree.
//
//   every rank owns a window  { slots[2][world][capacity] doubles, flags[2][world], epoch, ... }
//   allocated with cudaMalloc and opened by every peer through CUDA IPC.
//
//   step e (parity p = e & 1), rank r, one CTA per destination q:
//     push   : store the rank's vector into   window[q].slots[p][r][:]      (remote NVLink stores)
//     signal : fence, then                     window[q].flags[p][r] = e+1   (release, system scope)
//     wait   : spin until every                window[r].flags[p][s] >= e+1  (acquire, system scope)
//     sum    : CTA q adds its 1/world chunk of slots[p][0..world-1] in RANK ORDER (every rank gets
//              the same bits) and writes it to the caller's output vector (out of place: other
//              CTAs may still be pushing the input)
//   The epoch lives in device memory (one copy per CTA, bumped by that CTA at its end), so the kernel takes
//   no per-step argument and replays unchanged inside a CUDA graph.  Two slot parities suffice: a rank can
//   only start step e+2 after every peer signalled step e+1, i.e. after they finished reading e.
//
// This replaces nothing of the reference (it is single-GPU); it is the
// multi-GPU exchange of DESIGN.md section 6.  Included at the end of trace_kernels.cu.
#pragma once

namespace tlpeer {

constexpr int kMaxWorld = 16;
constexpr int kThreads = 512;
constexpr unsigned long long kSpinLimitNs = 4000000000ull;   // 4 s: a dead peer must not hang the box

struct Window {                  // header of a rank's window (device memory)
  unsigned int flags[2][kMaxWorld];
  unsigned int epoch;            // steps completed by this rank (CTA 0's copy: what tl_peer_status reports)
  unsigned int done;             // (unused since the per-CTA epochs)
  unsigned int status;           // 0 = ok, 1 = a wait timed out
  unsigned int cta_epoch[kMaxWorld];   // steps completed, one copy per CTA of the exchange kernel
  unsigned int pad[13];          // header = 256 bytes, slots 8-byte aligned
};
static_assert(sizeof(Window) == 256, "window header layout");

struct Peers {
  Window *win[kMaxWorld];        // peer windows as mapped into THIS process
};

__device__ __forceinline__ double *slots_of(Window *w) { return reinterpret_cast<double *>(w + 1); }

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void store_release_sys(unsigned int *p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned int load_acquire_sys(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// grid = world CTAs (CTA q serves destination q and output chunk q)
__global__ void __launch_bounds__(kThreads)
k_peer_allreduce(Peers peers, int rank, int world, long long capacity, const double *data, double *out,
                 long long n) {
  Window *mine = peers.win[rank];
  const int q = blockIdx.x;
  // every CTA keeps its own copy of the epoch: it is read at the start and bumped at the end by the same CTA of
  // consecutive launches (ordered by the stream), so the tail needs no cross-CTA counter, atomic or fence -- those
  // were 2-3 us on the critical path of the step's next kernel
  const unsigned int e = *reinterpret_cast<volatile unsigned int *>(&mine->cta_epoch[q]);
  const unsigned int p = e & 1u;
  const unsigned int want = e + 1u;

  // push this rank's vector into slot [p][rank] of destination q
  {
    double *dst = slots_of(peers.win[q]) + ((long long)p * world + rank) * capacity;
    // eight loads in flight per thread, then the eight remote stores (one load-to-store round trip per element
    // made this loop 5 us of a 27 KB push)
    long long i = threadIdx.x;
    for (; i + 7 * kThreads < n; i += 8 * kThreads) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = data[i + u * kThreads];
#pragma unroll
      for (int u = 0; u < 8; ++u) dst[i + u * kThreads] = v[u];
    }
    {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (i + u * kThreads < n) ? data[i + u * kThreads] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i + u * kThreads < n) dst[i + u * kThreads] = v[u];
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) store_release_sys(&peers.win[q]->flags[p][rank], want);

  // wait for every source to have pushed into this rank's window.  A wait that times out (a late
  // or dead peer) makes the failure LOUD and STICKY: status = 1, this and every later step write
  // NaN instead of a sum of stale slots (the NaN reaches the host with the loss, also through a
  // replayed CUDA graph), and later steps do not wait again (epochs may be out of step for good).
  __shared__ int timed_out;
  if (threadIdx.x == 0) timed_out = *reinterpret_cast<volatile unsigned int *>(&mine->status) != 0u;
  __syncthreads();
  const bool poisoned_before = timed_out != 0;
  if (!poisoned_before && threadIdx.x < world) {
    const unsigned int *flag = &mine->flags[p][threadIdx.x];
    const unsigned long long t0 = now_ns();
    while ((int)(load_acquire_sys(flag) - want) < 0) {
      if (now_ns() - t0 > kSpinLimitNs) { timed_out = 1; break; }
    }
  }
  __syncthreads();
  const bool poisoned = timed_out != 0;
  if (poisoned && !poisoned_before && threadIdx.x == 0) atomicExch(&mine->status, 1u);

  // sum chunk q over the ranks, in rank order
  {
    const long long per = (n + world - 1) / world;
    const long long lo = per * q, hi = (lo + per < n) ? lo + per : n;
    const double *src = slots_of(mine) + (long long)p * world * capacity;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (long long i = lo + threadIdx.x; i < hi; i += kThreads) {
      // all the ranks' values in flight together (kMaxWorld loads, predicated), then the sum in RANK ORDER
      double v[kMaxWorld];
#pragma unroll
      for (int s = 0; s < kMaxWorld; ++s) v[s] = s < world ? __ldcv(src + (long long)s * capacity + i) : 0.0;
      double acc = v[0];
#pragma unroll
      for (int s = 1; s < kMaxWorld; ++s)
        if (s < world) acc += v[s];
      out[i] = poisoned ? nan : acc;
    }
  }

  // this CTA's next epoch (and, from CTA 0, the copy the host reads)
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned int *>(&mine->cta_epoch[q]) = want;
    if (q == 0) *reinterpret_cast<volatile unsigned int *>(&mine->epoch) = want;
  }
}

struct Comm {
  int rank = 0, world = 1, device = 0;
  long long capacity = 0;
  Window *local = nullptr;
  Peers peers{};
  bool opened[kMaxWorld] = {};
  bool connected = false;
};

}  // namespace tlpeer

extern "C" {

size_t tl_peer_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int tl_peer_create(int32_t rank, int32_t world, int64_t capacity, TlPeerComm **comm_out, void *handle_out) {
  using namespace tlpeer;
  if (!comm_out || !handle_out) return fail(TL_ERR_INVALID, "tl_peer_create: null argument%s");
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || capacity < 1)
    return fail(TL_ERR_INVALID, "tl_peer_create: need 0 <= rank < world <= 16 and capacity >= 1%s");
  Comm *comm = new Comm();
  comm->rank = rank;
  comm->world = world;
  comm->capacity = capacity;
  TL_CHECK_CUDA(cudaGetDevice(&comm->device));
  const size_t bytes = sizeof(Window) + sizeof(double) * 2 * (size_t)world * (size_t)capacity;
  void *ptr = nullptr;
  cudaError_t err = cudaMalloc(&ptr, bytes);
  if (err != cudaSuccess) {
    delete comm;
    return fail(TL_ERR_CUDA, "tl_peer_create: cudaMalloc: %s", cudaGetErrorString(err));
  }
  comm->local = static_cast<Window *>(ptr);
  err = cudaMemset(ptr, 0, bytes);
  if (err == cudaSuccess) err = cudaDeviceSynchronize();
  if (err == cudaSuccess && world > 1)
    err = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle_out), ptr);
  if (err != cudaSuccess) {
    cudaFree(ptr);
    cudaGetLastError();
    delete comm;
    return fail(TL_ERR_CUDA, "tl_peer_create: %s", cudaGetErrorString(err));
  }
  if (world == 1) {
    memset(handle_out, 0, sizeof(cudaIpcMemHandle_t));
    comm->peers.win[0] = comm->local;
    comm->connected = true;
  }
  *comm_out = reinterpret_cast<TlPeerComm *>(comm);
  return TL_OK;
}

int tl_peer_connect(TlPeerComm *comm_, const void *all_handles) {
  using namespace tlpeer;
  Comm *comm = reinterpret_cast<Comm *>(comm_);
  if (!comm || !all_handles) return fail(TL_ERR_INVALID, "tl_peer_connect: null argument%s");
  if (comm->connected) return TL_OK;
  TL_CHECK_CUDA(cudaSetDevice(comm->device));
  const cudaIpcMemHandle_t *handles = static_cast<const cudaIpcMemHandle_t *>(all_handles);
  for (int q = 0; q < comm->world; ++q) {
    if (q == comm->rank) {
      comm->peers.win[q] = comm->local;
      continue;
    }
    void *ptr = nullptr;
    const cudaError_t err = cudaIpcOpenMemHandle(&ptr, handles[q], cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
      cudaGetLastError();      // not sticky: later launches must not inherit it
      return fail(TL_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(err));
    }
    comm->peers.win[q] = static_cast<Window *>(ptr);
    comm->opened[q] = true;
  }
  comm->connected = true;
  return TL_OK;
}

int tl_peer_allreduce_f64(TlPeerComm *comm_, const double *data, double *out, int64_t n, void *stream_) {
  using namespace tlpeer;
  Comm *comm = reinterpret_cast<Comm *>(comm_);
  if (!comm || !data || !out) return fail(TL_ERR_INVALID, "tl_peer_allreduce_f64: null argument%s");
  if (data == out && comm->world > 1)
    return fail(TL_ERR_INVALID, "tl_peer_allreduce_f64: in-place reduction is not supported%s");
  if (!comm->connected) return fail(TL_ERR_INVALID, "tl_peer_allreduce_f64: tl_peer_connect has not run%s");
  if (n < 1 || n > comm->capacity)
    return fail(TL_ERR_INVALID, "tl_peer_allreduce_f64: n outside [1, capacity]%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  k_peer_allreduce<<<comm->world, kThreads, 0, stream>>>(comm->peers, comm->rank, comm->world,
                                                          (long long)comm->capacity, data, out, (long long)n);
  g_launches++;
  TL_CHECK_CUDA(cudaGetLastError());
  return TL_OK;
}

int tl_peer_status(TlPeerComm *comm_, int32_t *status_out, uint32_t *epoch_out) {
  using namespace tlpeer;
  Comm *comm = reinterpret_cast<Comm *>(comm_);
  if (!comm || !status_out) return fail(TL_ERR_INVALID, "tl_peer_status: null argument%s");
  Window head;
  TL_CHECK_CUDA(cudaMemcpy(&head, comm->local, sizeof(Window), cudaMemcpyDeviceToHost));
  *status_out = (int32_t)head.status;
  if (epoch_out) *epoch_out = head.epoch;
  return TL_OK;
}

int tl_peer_destroy(TlPeerComm *comm_) {
  using namespace tlpeer;
  Comm *comm = reinterpret_cast<Comm *>(comm_);
  if (!comm) return TL_OK;
  cudaSetDevice(comm->device);
  cudaDeviceSynchronize();
  for (int q = 0; q < comm->world; ++q)
    if (comm->opened[q]) cudaIpcCloseMemHandle(comm->peers.win[q]);
  if (comm->local) cudaFree(comm->local);
  delete comm;
  return TL_OK;
}

}  // extern "C"
