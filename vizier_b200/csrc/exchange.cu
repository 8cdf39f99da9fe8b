// Global top-k over candidate-pool shards: one fused "push + publish + wait + merge" kernel over NVLink
// peer memory, with an in-library NCCL all-gather as the checked fallback.
//
// The reference is a single process: its counterpart is the arg-partition over ONE pool at
// vizier/_src/algorithms/optimizers/vectorized_base.py:575-587.  Here the pool is sharded over the GPUs
// of one box and every rank needs the global top-`count`.  The payload is count*(Dc+2)
// doubles per rank, so the exchange is pure latency: an NCCL all-gather costs a launch + protocol
// round (~15-25 us) and couples the ranks' host threads; this kernel costs one launch of one CTA:
//   1. every thread copies this rank's rows into slot [rank] of EVERY rank's buffer (plain stores through
//      the peer mappings: NVLink writes; the local copy included),
//   2. __threadfence_system + CTA barrier, then one st.release.sys per peer raises flag [slot][rank] there
//      to the step's sequence number,
//   3. `world` threads spin with ld.acquire.sys on the local flags until every rank has published,
//   4. the CTA merges the world*count rows (merge_topk_block) - identical code and inputs on every rank,
//      so every rank holds the identical result without a broadcast.
// Two slots alternate by sequence parity: a rank can be at most one step ahead of the slowest one (it
// needs everybody's flag of step t before it can publish step t+1), so slot t%2 is never overwritten
// while a peer still merges step t.  A rank that never shows up is a deadlock by nature; the wait is
// bounded by a timeout (globaltimer) that raises a status flag instead of hanging the GPU.
#include <dlfcn.h>

#include <cstring>

#include "launchers.h"
#include "topk_merge.cuh"

namespace vzgp {

constexpr int kMaxWorld = 32;
constexpr int kSlots = 2;

struct ExchangeArgs {
  double* peer[kMaxWorld];          // base of every rank's exchange buffer as mapped in this process
  unsigned long long seq;           // this step's sequence number (1, 2, ...)
  unsigned long long timeout_ns;
  const double* payload;            // [count x width] this rank's rows
  double* out;                      // [count x width] merged rows
  int* status;                      // device int: set to 1 if a peer did not publish in time
  int rank, world, count, width;
};

__host__ __device__ inline size_t slot_doubles(int world, int count, int width) { return (size_t)world * count * width; }
__host__ __device__ inline size_t flags_offset_doubles(int world, int count, int width) {
  return kSlots * slot_doubles(world, count, width);
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(256) k_exchange_merge(ExchangeArgs a) {
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  const int n = a.count * a.width;
  const int slot = (int)(a.seq % kSlots);
  const size_t sd = slot_doubles(a.world, a.count, a.width);
  const size_t fo = flags_offset_doubles(a.world, a.count, a.width);
  if (tid == 0) s_fail = 0;
  // 1. push
  for (int p = 0; p < a.world; ++p) {
    double* dst = a.peer[p] + slot * sd + (size_t)a.rank * n;
    for (int e = tid; e < n; e += 256) dst[e] = a.payload[e];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish
  if (tid < a.world) {
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(a.peer[tid] + fo) + slot * a.world + a.rank;
    st_release_sys(flag, a.seq);
  }
  // 3. wait for every rank's rows of this step
  if (tid < a.world) {
    const unsigned long long* flag =
        reinterpret_cast<const unsigned long long*>(a.peer[a.rank] + fo) + slot * a.world + tid;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flag) < a.seq) {
      if (global_timer_ns() - t0 > a.timeout_ns) { s_fail = 1; break; }
      __nanosleep(40);
    }
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) *a.status = 1;
    for (int e = tid; e < n; e += 256) a.out[e] = (e % a.width == 0) ? -INFINITY : (e % a.width == 1 ? -1.0 : 0.0);
    return;
  }
  // 4. merge (rows were written by peers: read them past L1)
  merge_topk_block<true>(a.peer[a.rank] + slot * sd, a.world * a.count, a.width, a.count, a.out);
}

// ---- NCCL through dlopen (no link-time dependency; torch ships libnccl.so.2) ----------------------
struct NcclId { char internal[128]; };
struct NcclApi {
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
static const NcclApi& nccl_api() {
  static const NcclApi api = [] {
    NcclApi a;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy torch already loaded, if any
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return a;
    a.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(lib, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclId, int)>(dlsym(lib, "ncclCommInitRank"));
    a.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(dlsym(lib, "ncclAllGather"));
    a.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(lib, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(lib, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllGather && a.CommDestroy;
    return a;
  }();
  return api;
}
constexpr int kNcclFloat64 = 8;

}  // namespace vzgp

using namespace vzgp;

struct vzgp_exchange {
  int device = 0, rank = 0, world = 1, count = 1, width = 2;
  double* local = nullptr;               // cudaMalloc: slots | flags | status
  size_t bytes = 0;
  double* peer[kMaxWorld] = {};
  bool ipc_opened[kMaxWorld] = {};
  bool peers_set = false;
  unsigned long long seq = 0;
  unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;
  void* nccl_comm = nullptr;
  double* gathered = nullptr;            // NCCL fallback: [world*count x width]
};

static int* exchange_status_ptr(const vzgp_exchange* x) {
  return reinterpret_cast<int*>(x->local + flags_offset_doubles(x->world, x->count, x->width) + kSlots * x->world);
}

extern "C" {

int vzgp_exchange_create(vzgp_handle* h, int rank, int world, int count, int width, vzgp_exchange** out) {
  VZ_ARG(h && out, "handle / out");
  *out = nullptr;
  VZ_ARG(world >= 1 && world <= kMaxWorld, "1 <= world <= 32");
  VZ_ARG(rank >= 0 && rank < world, "rank");
  VZ_ARG(count >= 1 && count <= 256 && world * count <= 2048, "count (<= 256, world*count <= 2048)");
  VZ_ARG(width >= 2, "width >= 2");
  cudaSetDevice(h->device);
  vzgp_exchange* x = new vzgp_exchange();
  x->device = h->device; x->rank = rank; x->world = world; x->count = count; x->width = width;
  x->bytes = sizeof(double) * (flags_offset_doubles(world, count, width) + kSlots * world + 8);
  cudaError_t e = cudaMalloc(&x->local, x->bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(exchange): %s", cudaGetErrorString(e)); delete x; return VZGP_ERR_CUDA; }
  // zeroed and complete before any peer can learn the address (callers exchange handles afterwards)
  e = cudaMemset(x->local, 0, x->bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { set_error("exchange init: %s", cudaGetErrorString(e)); cudaFree(x->local); delete x; return VZGP_ERR_CUDA; }
  x->peer[rank] = x->local;
  x->peers_set = (world == 1);
  if (const char* t = getenv("VZGP_EXCHANGE_TIMEOUT_MS")) x->timeout_ns = (unsigned long long)atoll(t) * 1000000ull;
  *out = x;
  return 0;
}

int vzgp_exchange_destroy(vzgp_exchange* x) {
  if (!x) return 0;
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < x->world; ++p)
    if (x->ipc_opened[p]) cudaIpcCloseMemHandle(x->peer[p]);
  if (x->nccl_comm && nccl_api().ok) nccl_api().CommDestroy(x->nccl_comm);
  if (x->gathered) cudaFree(x->gathered);
  if (x->local) cudaFree(x->local);
  delete x;
  return 0;
}

/* cudaIpcMemHandle_t of this rank's buffer (64 bytes) for the other processes. */
int vzgp_exchange_ipc_handle(vzgp_exchange* x, void* handle_out) {
  VZ_ARG(x && handle_out, "exchange / out");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaSetDevice(x->device);
  cudaIpcMemHandle_t hd;
  VZ_CUDA(cudaIpcGetMemHandle(&hd, x->local));
  std::memcpy(handle_out, &hd, sizeof(hd));
  return 0;
}

/* handles: [world][64] bytes in rank order (this rank's own entry is ignored). */
int vzgp_exchange_open(vzgp_exchange* x, const void* handles) {
  VZ_ARG(x && handles, "exchange / handles");
  cudaSetDevice(x->device);
  for (int p = 0; p < x->world; ++p) {
    if (p == x->rank) continue;
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, static_cast<const char*>(handles) + 64 * (size_t)p, sizeof(hd));
    void* ptr = nullptr;
    VZ_CUDA(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    x->peer[p] = static_cast<double*>(ptr);
    x->ipc_opened[p] = true;
  }
  x->peers_set = true;
  return 0;
}

/* Same-process peers (several handles driven by one process, or the single-GPU loop-back test): base
 * pointers returned by vzgp_exchange_base of every rank, in rank order. */
int vzgp_exchange_set_peers(vzgp_exchange* x, void* const* bases) {
  VZ_ARG(x && bases, "exchange / bases");
  cudaSetDevice(x->device);
  for (int p = 0; p < x->world; ++p) {
    if (p == x->rank) continue;
    VZ_ARG(bases[p] != nullptr, "peer base");
    cudaPointerAttributes at;
    VZ_CUDA(cudaPointerGetAttributes(&at, bases[p]));
    if (at.device != x->device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        set_error("cudaDeviceEnablePeerAccess(%d): %s", at.device, cudaGetErrorString(e));
        return VZGP_ERR_CUDA;
      }
      cudaGetLastError();
    }
    x->peer[p] = static_cast<double*>(bases[p]);
  }
  x->peers_set = true;
  return 0;
}

void* vzgp_exchange_base(vzgp_exchange* x) { return x ? x->local : nullptr; }

int vzgp_nccl_unique_id(void* id_out128) {
  VZ_ARG(id_out128 != nullptr, "out");
  if (!nccl_api().ok) { set_error("libnccl.so.2 could not be loaded"); return VZGP_ERR_UNSUPPORTED; }
  NcclId id;
  int r = nccl_api().GetUniqueId(&id);
  if (r != 0) { set_error("ncclGetUniqueId: %s", nccl_api().GetErrorString ? nccl_api().GetErrorString(r) : "?"); return VZGP_ERR_CUDA; }
  std::memcpy(id_out128, &id, sizeof(id));
  return 0;
}

/* Collective: every rank calls it with rank 0's id. */
int vzgp_exchange_nccl_init(vzgp_exchange* x, const void* id128) {
  VZ_ARG(x && id128, "exchange / id");
  if (!nccl_api().ok) { set_error("libnccl.so.2 could not be loaded"); return VZGP_ERR_UNSUPPORTED; }
  cudaSetDevice(x->device);
  NcclId id;
  std::memcpy(&id, id128, sizeof(id));
  int r = nccl_api().CommInitRank(&x->nccl_comm, x->world, id, x->rank);
  if (r != 0) { set_error("ncclCommInitRank: %s", nccl_api().GetErrorString ? nccl_api().GetErrorString(r) : "?"); x->nccl_comm = nullptr; return VZGP_ERR_CUDA; }
  VZ_CUDA(cudaMalloc(&x->gathered, sizeof(double) * (size_t)x->world * x->count * x->width));
  return 0;
}

int vzgp_allgather_topk(vzgp_handle* h, vzgp_exchange* x, const double* payload_dev, double* out_dev,
                        double* host_out, int use_nccl) {
  VZ_ARG(h && x && payload_dev && out_dev, "handle / exchange / pointers");
  VZ_ARG(h->device == x->device, "handle and exchange must live on the same device");
  cudaSetDevice(h->device);
  const size_t n = (size_t)x->count * x->width;
  if (use_nccl) {
    if (!x->nccl_comm) { set_error("vzgp_allgather_topk(use_nccl): call vzgp_exchange_nccl_init first"); return VZGP_ERR_STATE; }
    int r = nccl_api().AllGather(payload_dev, x->gathered, n, kNcclFloat64, x->nccl_comm, h->stream);
    if (r != 0) { set_error("ncclAllGather: %s", nccl_api().GetErrorString ? nccl_api().GetErrorString(r) : "?"); return VZGP_ERR_CUDA; }
    VZ_TRY(launch_merge_topk(h, x->gathered, x->world * x->count, x->width, x->count, out_dev));
  } else {
    if (!x->peers_set) { set_error("vzgp_allgather_topk: peers not mapped (vzgp_exchange_open / _set_peers)"); return VZGP_ERR_STATE; }
    ExchangeArgs a;
    for (int p = 0; p < kMaxWorld; ++p) a.peer[p] = p < x->world ? x->peer[p] : nullptr;
    a.seq = ++x->seq;
    a.timeout_ns = x->timeout_ns;
    a.payload = payload_dev; a.out = out_dev; a.status = exchange_status_ptr(x);
    a.rank = x->rank; a.world = x->world; a.count = x->count; a.width = x->width;
    k_exchange_merge<<<1, 256, 0, h->stream>>>(a);
    VZ_CHECK_LAUNCH();
    h->launches++;
  }
  if (host_out) VZ_CUDA(cudaMemcpyAsync(host_out, out_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

/* Synchronises the handle's stream; *status_out = 1 if any fused exchange so far timed out on a peer. */
int vzgp_exchange_status(vzgp_handle* h, vzgp_exchange* x, int* status_out) {
  VZ_ARG(h && x && status_out, "pointers");
  cudaSetDevice(h->device);
  VZ_CUDA(cudaStreamSynchronize(h->stream));
  VZ_CUDA(cudaMemcpy(status_out, exchange_status_ptr(x), sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
