// dist.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink 5.
//
// The reference has no multi-process mode at all; this
// is the sharding the north star asks for: rank r owns the amplitudes whose top
// log2(P) physical index bits equal r.  Diagonal gates and controls on a global
// position need no communication.  A pairing gate whose target sits on a global
// position g is made local by SWAPPING physical positions g and a local
// position l: each rank keeps the half of its shard whose bit l equals its own
// g-bit and trades the other half with partner rank ^ (1 << (g - nl)).  Half a
// shard crosses NVLink in each direction (8 * 2^nl bytes per rank per
// direction), after which the engine's logical->physical permutation records
// the new layout and every later gate on that qubit is local.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library itself is dlopen'ed

#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../../include/qcs_cuda.h"
#include "engine.h"

namespace qcs {

DistContext &dist() {
  static DistContext d;
  return d;
}

// NCCL is bound at run time, not at link time: a process that also hosts
// PyTorch must end up with ONE libnccl.so.2 (torch's bundled copy), and a
// single-GPU user should not need NCCL installed at all.  dlopen by SONAME
// returns the copy already in the process if there is one.
namespace nccl_api {
#define QCS_NCCL_FUNCS(X)                                                              \
  X(ncclGetUniqueId) X(ncclCommInitRank) X(ncclCommDestroy) X(ncclGetErrorString)      \
  X(ncclGroupStart) X(ncclGroupEnd) X(ncclSend) X(ncclRecv) X(ncclAllReduce)           \
  X(ncclAllGather)
#define X(name) static decltype(&::name) qn_##name = nullptr;
QCS_NCCL_FUNCS(X)
#undef X
static bool loaded = false;
static int load() {
  if (loaded) return QCS_CUDA_OK;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return set_error(QCS_CUDA_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define X(name)                                                                        \
  qn_##name = (decltype(&::name))dlsym(h, #name);                                      \
  if (!qn_##name) return set_error(QCS_CUDA_ERR_NCCL, "libnccl lacks %s", #name);
  QCS_NCCL_FUNCS(X)
#undef X
  loaded = true;
  return QCS_CUDA_OK;
}
}  // namespace nccl_api
// From here on the NCCL names resolve to the run-time bound pointers.
#define ncclGetUniqueId nccl_api::qn_ncclGetUniqueId
#define ncclCommInitRank nccl_api::qn_ncclCommInitRank
#define ncclCommDestroy nccl_api::qn_ncclCommDestroy
#define ncclGetErrorString nccl_api::qn_ncclGetErrorString
#define ncclGroupStart nccl_api::qn_ncclGroupStart
#define ncclGroupEnd nccl_api::qn_ncclGroupEnd
#define ncclSend nccl_api::qn_ncclSend
#define ncclRecv nccl_api::qn_ncclRecv
#define ncclAllReduce nccl_api::qn_ncclAllReduce
#define ncclAllGather nccl_api::qn_ncclAllGather

static void *g_small_dev = nullptr;  // scratch for host-visible small collectives
static const size_t kSmallBytes = 1 << 20;

static int check_nccl(ncclResult_t r, const char *what) {
  if (r == ncclSuccess) return QCS_CUDA_OK;
  return set_error(QCS_CUDA_ERR_NCCL, "%s: %s", what, ncclGetErrorString(r));
}
#define NK(call)                              \
  do {                                        \
    int rc_ = check_nccl((call), #call);      \
    if (rc_ != QCS_CUDA_OK) return rc_;       \
  } while (0)
#define CK(call)                              \
  do {                                        \
    int rc_ = check_cuda((call), #call);      \
    if (rc_ != QCS_CUDA_OK) return rc_;       \
  } while (0)

namespace {

// staging[j] = live[expand(j)] (pack) or the reverse (unpack), where expand
// inserts bit value `bitval` at position `pos`.
__global__ void __launch_bounds__(256)
pack_half_kernel(const double2 *__restrict__ live, double2 *__restrict__ staging, uint64_t n_half,
                 int pos, uint64_t bitval) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_half; j += stride) {
    const uint64_t low = j & ((1ull << pos) - 1ull);
    const uint64_t i = (((j >> pos) << (pos + 1)) | low) | (bitval << pos);
    __stcs(staging + j, __ldcs(live + i));
  }
}

__global__ void __launch_bounds__(256)
unpack_half_kernel(double2 *__restrict__ live, const double2 *__restrict__ staging,
                   uint64_t n_half, int pos, uint64_t bitval) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_half; j += stride) {
    const uint64_t low = j & ((1ull << pos) - 1ull);
    const uint64_t i = (((j >> pos) << (pos + 1)) | low) | (bitval << pos);
    __stcs(live + i, __ldcs(staging + j));
  }
}

// In-place position swap over NVLink peer memory: pair j = (my element with bit lpos ==
// `leaving`, the partner's element with the opposite bit).  Each rank handles half of the pairs
// (`share`), reading one side remotely and writing one side remotely, so both directions of the
// link carry 8 * 2^nl bytes -- the minimum -- and no staging buffer or pack/unpack pass exists.
__global__ void __launch_bounds__(256)
p2p_swap_kernel(double2 *__restrict__ mine, double2 *__restrict__ peer, uint64_t pairs_begin,
                uint64_t pairs_end, int pos, uint64_t leaving) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t bit = 1ull << pos;
  for (uint64_t j = pairs_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < pairs_end;
       j += stride) {
    const uint64_t low = j & (bit - 1ull);
    const uint64_t i = (((j >> pos) << (pos + 1)) | low) | (leaving << pos);
    const uint64_t ip = i ^ bit;
    const double2 a = mine[i];
    const double2 b = peer[ip];
    mine[i] = b;
    peer[ip] = a;
  }
}

}  // namespace

// ---- peer mappings and flag arrays outlive the engines that first needed them ----------------------
// Programs in the reference's style create and destroy circuits freely; on a sharded engine every
// qc_create used to open 2 x (world - 1) CUDA IPC handles (milliseconds each for a 16 GiB buffer),
// allocate and clear a flag array and close everything again in qc_destroy -- a quarter of the
// end-to-end time of a 33-qubit QFT on 8 GPUs (BENCH_r01 / SCALE_r01).  State buffers come back from
// the buffer pool with the same address and the same IPC handle, so the mappings are cached by
// (rank, handle) until qcs_cuda_dist_finalize; a buffer a peer may still have mapped is pinned in the
// pool (pool_pin, engine.cu) so that it is never freed underneath that mapping.
namespace {
struct PeerMapping {
  int rank;
  cudaIpcMemHandle_t handle;
  void *ptr;
};
std::vector<PeerMapping> g_peer_maps;
struct FlagArray {
  size_t n_tiles;
  uint32_t *ptr;
  bool in_use;  // engines alive at the same time never share one (their passes could interleave)
};
std::vector<FlagArray> g_flag_arrays;
uint32_t g_swap_epoch = 0;  // one sequence for all engines of the process: flag arrays are reused, epochs are not

void *open_peer_cached(int rank, const cudaIpcMemHandle_t &h) {
  for (const PeerMapping &m : g_peer_maps)
    if (m.rank == rank && std::memcmp(&m.handle, &h, sizeof(h)) == 0) return m.ptr;
  void *ptr = nullptr;
  if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  g_peer_maps.push_back(PeerMapping{rank, h, ptr});
  return ptr;
}
}  // namespace

// Exchanges CUDA IPC handles of every rank's state buffer and tile-flag array so that partners
// can address them.
int dist_open_peers(Engine &e) {
  DistContext &d = dist();
  if (!d.active || e.opt.dryrun) return QCS_CUDA_OK;
  const size_t n_tiles = e.nl >= QCS_MIN_TILE_BITS ? (size_t)1 << (e.nl - QCS_MIN_TILE_BITS) : 0;
  if (n_tiles) {
    // per sending rank of a remap group (<= 2^QCS_MAX_REMAP) one word per tile + (behind them) the
    // abort word of the bounded handshake wait
    const size_t words = (n_tiles << QCS_MAX_REMAP) + 1;
    for (FlagArray &f : g_flag_arrays)
      if (f.n_tiles == n_tiles && !f.in_use) {
        f.in_use = true;
        e.tile_flags = f.ptr;
        break;
      }
    if (!e.tile_flags) {
      CK(cudaMalloc(&e.tile_flags, words * sizeof(uint32_t)));
      CK(cudaMemset(e.tile_flags, 0, words * sizeof(uint32_t)));
      g_flag_arrays.push_back(FlagArray{n_tiles, e.tile_flags, true});
    }
    e.tile_flag_stride = (uint32_t)n_tiles;
    e.swap_abort_flag = e.tile_flags + (n_tiles << QCS_MAX_REMAP);
    CK(cudaMemsetAsync(e.swap_abort_flag, 0, sizeof(uint32_t), e.stream));
    if (!e.swap_status_host) CK(cudaMallocHost(&e.swap_status_host, sizeof(int)));
    *e.swap_status_host = 0;
  }
  struct Handles { cudaIpcMemHandle_t live, alt, flags; int has_alt; } mine;
  std::memset(&mine, 0, sizeof(mine));
  CK(cudaIpcGetMemHandle(&mine.live, e.live));
  pool_pin(e.live);  // peers keep it mapped after this engine is gone
  if (e.alt && e.tile_flags && cudaIpcGetMemHandle(&mine.alt, e.alt) == cudaSuccess) mine.has_alt = 1;
  else cudaGetLastError();
  if (e.tile_flags) CK(cudaIpcGetMemHandle(&mine.flags, e.tile_flags));
  std::vector<Handles> all(d.world);
  int rc = dist_allgather_host(&mine, all.data(), sizeof(mine));
  if (rc) return rc;
  // the second buffer is used by all ranks or by none
  bool all_alt = true;
  for (int r = 0; r < d.world; r++) all_alt = all_alt && all[r].has_alt;
  auto drop_alt = [&]() {
    pool_release(e.alt, e.local_size * sizeof(double2));
    e.alt = nullptr;
    e.peer_alt.clear();
  };
  if (!all_alt) drop_alt();
  else pool_pin(e.alt);
  e.peer_live.assign(d.world, nullptr);
  e.peer_flags.assign(d.world, nullptr);
  if (e.alt) e.peer_alt.assign(d.world, nullptr);
  for (int r = 0; r < d.world; r++) {
    if (r == d.rank) {
      e.peer_live[r] = e.live;
      e.peer_flags[r] = e.tile_flags;
      if (e.alt) e.peer_alt[r] = e.alt;
      continue;
    }
    void *ptr = open_peer_cached(r, all[r].live);
    void *fptr = (ptr && e.tile_flags) ? open_peer_cached(r, all[r].flags) : nullptr;
    void *aptr = (ptr && e.alt) ? open_peer_cached(r, all[r].alt) : nullptr;
    if (!ptr || (e.tile_flags && !fptr) || (e.alt && !aptr)) {
      e.peer_live.clear();  // no peer access: the NCCL path stays in charge
      e.peer_flags.clear();
      drop_alt();  // (every rank fails the same way: peer access is symmetric)
      return QCS_CUDA_OK;
    }
    e.peer_live[r] = (double2 *)ptr;
    e.peer_flags[r] = (uint32_t *)fptr;
    if (e.alt) e.peer_alt[r] = (double2 *)aptr;
  }
  return QCS_CUDA_OK;
}

// qc_destroy: the mappings stay (cache above); the flag array goes back on the shelf.
void dist_close_peers(Engine &e) {
  e.peer_alt.clear();
  e.peer_live.clear();
  e.peer_flags.clear();
  for (FlagArray &f : g_flag_arrays)
    if (f.ptr == e.tile_flags) f.in_use = false;
  e.tile_flags = nullptr;
  e.swap_abort_flag = nullptr;
  if (e.swap_status_host) cudaFreeHost(e.swap_status_host);
  e.swap_status_host = nullptr;
}

// qcs_cuda_dist_finalize: every rank closes what it mapped, then (behind a barrier) frees what it exported.
static void release_peer_cache() {
  for (const PeerMapping &m : g_peer_maps) cudaIpcCloseMemHandle(m.ptr);
  g_peer_maps.clear();
  double dummy = 0.0;
  std::vector<double> all(dist().world > 0 ? dist().world : 1);
  dist_allgather_host(&dummy, all.data(), sizeof(double));  // nobody maps my buffers any more
  for (const FlagArray &f : g_flag_arrays) cudaFree(f.ptr);
  g_flag_arrays.clear();
  pool_unpin_all();
}

static int stream_barrier(Engine &e, cudaStream_t stream) {
  // a 4-byte all-reduce on `stream`: completes only once every rank's stream got here
  DistContext &d = dist();
  int *flag = (int *)((char *)g_small_dev + kSmallBytes - 64);  // clear of dist_allgather_host's area
  NK(ncclAllReduce(flag, flag, 1, ncclInt, ncclSum, (ncclComm_t)d.comm, stream));
  (void)e;
  return QCS_CUDA_OK;
}

// How long a CTA of a swap-carrying pass waits for its partner CTA (QCS_CUDA_SWAP_TIMEOUT_MS, default
// 20 s: the ranks enter a pass up to a whole pass apart, but never seconds).
static unsigned long long swap_spin_limit() {
  static unsigned long long ticks = 0;
  if (!ticks) {
    const char *v = getenv("QCS_CUDA_SWAP_TIMEOUT_MS");
    double ms = v ? atof(v) : 20000.0;
    if (!(ms > 0)) ms = 20000.0;
    int dev = 0, khz = 1965000;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    ticks = (unsigned long long)(ms * (double)khz);
  }
  return ticks;
}

bool dist_p2p_available(const Engine &e) {
  return dist().active && e.opt.exchange == 1 && e.opt.sem == SEM_CORRECTED && !e.peer_live.empty();
}

// A remap of k position pairs folded into the stores of a fused pass (kernels.h SwapStore): fills in
// the addresses of the 2^k ranks of the exchange.  False when the peer-memory path is not available.
bool dist_fused_swap_args(Engine &e, int k, const int *lpos, const int *gpos, SwapStore &sw) {
  if (!dist_p2p_available(e) || !e.tile_flags || k < 1 || k > QCS_MAX_REMAP) return false;
  DistContext &d = dist();
  std::memset(&sw, 0, sizeof(sw));
  sw.k = (uint32_t)k;
  for (int i = 0; i < k; i++) {
    sw.lpos[i] = (uint32_t)lpos[i];
    sw.my_gbits |= (uint32_t)((d.rank >> (gpos[i] - e.nl)) & 1) << i;
  }
  for (int b = 0; b < (1 << k); b++) {
    int r = d.rank;
    for (int i = 0; i < k; i++) {
      const int gbit = gpos[i] - e.nl;
      r = (r & ~(1 << gbit)) | (((b >> i) & 1) << gbit);
    }
    if (!e.peer_live[r] || !e.peer_flags[r]) return false;
    sw.peer[b] = e.alt ? e.peer_alt[r] : e.peer_live[r];
    sw.peer_flags[b] = e.peer_flags[r] + (size_t)sw.my_gbits * e.tile_flag_stride;
  }
  sw.my_flags = e.tile_flags;
  sw.flag_stride = e.tile_flag_stride;
  sw.epoch = ++g_swap_epoch;
  e.swap_epoch = sw.epoch;
  sw.abort_flag = e.swap_abort_flag;
  sw.spin_limit = swap_spin_limit();
  sw.out_of_place = e.alt ? 1u : 0u;
  sw.own_out = e.alt;
  e.swap_out_of_place = e.alt != nullptr;
  return true;  // in_tile, n_out, out_pair, out_tile_bit, bulk, row_bits: the caller knows the pass's tile
}

// After a pass that stored into the partner's shard: nobody reads swapped data before both
// ranks' kernels (hence their remote stores) have completed.  The same 4-byte all-reduce carries the
// abort words of the bounded handshake wait: its sum lands in pinned host memory and is looked at
// at the next host synchronisation (dist_check_fused_swaps) -- non-zero on EVERY rank when any CTA
// on any rank gave up waiting for its partner.
int dist_after_fused_swap(Engine &e) {
  DistContext &d = dist();
  if (e.swap_out_of_place) {
    // the pass read `live` and wrote every rank's `alt`: the buffers trade roles -- on every rank, the pass
    // is collective.  The all-reduce below is also what makes the old `live` safe to overwrite in the
    // NEXT carrying pass: no rank gets past it before every rank's kernel has finished reading.
    std::swap(e.live, e.alt);
    std::swap(e.peer_live, e.peer_alt);
    e.swap_out_of_place = false;
    e.out_of_place_remaps++;
  }
  int *flag = (int *)((char *)g_small_dev + kSmallBytes - 64);  // clear of dist_allgather_host's area
  CK(cudaMemcpyAsync(flag, e.swap_abort_flag, sizeof(int), cudaMemcpyDeviceToDevice, e.stream));
  NK(ncclAllReduce(flag, flag, 1, ncclInt, ncclSum, (ncclComm_t)d.comm, e.stream));
  CK(cudaMemcpyAsync(e.swap_status_host, flag, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  e.swap_status_pending = true;
  return QCS_CUDA_OK;
}

// Call after the stream has been synchronised.
int dist_check_fused_swaps(Engine &e) {
  if (!e.swap_status_pending) return QCS_CUDA_OK;
  e.swap_status_pending = false;
  if (e.swap_status_host && *e.swap_status_host != 0) {
    e.poisoned = true;
    return set_error(QCS_CUDA_ERR_CUDA,
                     "position swap: %d rank(s) gave up waiting for the partner's tiles (partner kernel not "
                     "running: failed launch, diverged plan or a co-tenant on the GPU); the state is invalid",
                     *e.swap_status_host);
  }
  return QCS_CUDA_OK;
}

// In-place position swap on `stream`.
int dist_p2p_swap(Engine &e, cudaStream_t stream, int lpos, int gpos) {
  DistContext &d = dist();
  const int gbit = gpos - e.nl;
  const int partner = d.rank ^ (1 << gbit);
  const uint64_t mybit = (uint64_t)((d.rank >> gbit) & 1);
  const uint64_t n_pairs = e.local_size >> 1;
  const uint64_t begin = mybit ? n_pairs / 2 : 0, end = mybit ? n_pairs : n_pairs / 2;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (e.timing) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, stream);
  }
  int rc = stream_barrier(e, stream);  // the partner finished its previous work on these elements
  if (rc) return rc;
  p2p_swap_kernel<<<(unsigned)device_sm_count() * 16, 256, 0, stream>>>(e.live, e.peer_live[partner], begin, end, lpos,
                                                1ull - mybit);
  CK(cudaGetLastError());
  e.kernel_launches++;
  rc = stream_barrier(e, stream);      // nobody reads swapped data before both halves of the pairs are done
  if (rc) return rc;
  if (ev0) {
    cudaEventRecord(ev1, stream);
    e.pending_xchg_events.emplace_back(ev0, ev1);
  }
  return QCS_CUDA_OK;
}

int dist_swap_positions(Engine &e, int lpos, int gpos) {
  DistContext &d = dist();
  if (!d.active) return set_error(QCS_CUDA_ERR_INVALID, "global position without a communicator");
  if (dist_p2p_available(e)) return dist_p2p_swap(e, e.stream, lpos, gpos);
  ncclComm_t comm = (ncclComm_t)d.comm;
  const int gbit = gpos - e.nl;
  const int partner = d.rank ^ (1 << gbit);
  const uint64_t mybit = (uint64_t)((d.rank >> gbit) & 1);
  const uint64_t leaving = 1ull - mybit;  // the half whose bit lpos differs from my g-bit leaves
  const uint64_t n_half = e.local_size >> 1;
  const size_t half_bytes = n_half * sizeof(double2);
  if (!e.staging) CK(cudaMalloc(&e.staging, 2 * half_bytes));  // returned to the engine's pool on destroy
  double2 *send_buf = e.staging;
  double2 *recv_buf = e.staging + n_half;

  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (e.timing) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, e.stream);
  }
  const bool contiguous = (lpos == e.nl - 1);
  const double2 *src = send_buf;
  if (contiguous) {
    src = e.live + leaving * n_half;  // the leaving half is one contiguous range
  } else {
    pack_half_kernel<<<(unsigned)device_sm_count() * 8, 256, 0, e.stream>>>(e.live, send_buf, n_half, lpos, leaving);
    CK(cudaGetLastError());
    e.kernel_launches++;
  }
  NK(ncclGroupStart());
  NK(ncclSend(src, half_bytes, ncclChar, partner, comm, e.stream));
  NK(ncclRecv(recv_buf, half_bytes, ncclChar, partner, comm, e.stream));
  NK(ncclGroupEnd());
  if (contiguous) {
    CK(cudaMemcpyAsync(e.live + leaving * n_half, recv_buf, half_bytes, cudaMemcpyDeviceToDevice,
                       e.stream));
  } else {
    unpack_half_kernel<<<(unsigned)device_sm_count() * 8, 256, 0, e.stream>>>(e.live, recv_buf, n_half, lpos, leaving);
    CK(cudaGetLastError());
    e.kernel_launches++;
  }
  if (ev0) {
    cudaEventRecord(ev1, e.stream);
    e.pending_xchg_events.emplace_back(ev0, ev1);
  }
  return QCS_CUDA_OK;
}

int dist_allreduce_sum(Engine &e, double *dev_values, int count) {
  DistContext &d = dist();
  if (!d.active) return QCS_CUDA_OK;
  NK(ncclAllReduce(dev_values, dev_values, count, ncclDouble, ncclSum, (ncclComm_t)d.comm,
                   e.stream));
  return QCS_CUDA_OK;
}

int dist_allreduce_max_i64(Engine &e, long long *dev_values, size_t count) {
  DistContext &d = dist();
  if (!d.active) return QCS_CUDA_OK;
  NK(ncclAllReduce(dev_values, dev_values, count, ncclInt64, ncclMax, (ncclComm_t)d.comm,
                   e.stream));
  return QCS_CUDA_OK;
}

int dist_allgather_host(const void *mine, void *all, size_t bytes_each) {
  DistContext &d = dist();
  if (!d.active) {
    std::memcpy(all, mine, bytes_each);
    return QCS_CUDA_OK;
  }
  if (bytes_each * (size_t)(d.world + 1) > kSmallBytes - 64)
    return set_error(QCS_CUDA_ERR_INVALID, "dist_allgather_host: payload too large");
  char *dev = (char *)g_small_dev;
  char *dev_all = dev + bytes_each;
  CK(cudaMemcpy(dev, mine, bytes_each, cudaMemcpyHostToDevice));
  NK(ncclAllGather(dev, dev_all, bytes_each, ncclChar, (ncclComm_t)d.comm, 0));
  CK(cudaStreamSynchronize(0));
  CK(cudaMemcpy(all, dev_all, bytes_each * d.world, cudaMemcpyDeviceToHost));
  return QCS_CUDA_OK;
}

int dist_barrier(Engine &e) {
  double dummy = 0.0;
  std::vector<double> all(dist().world > 0 ? dist().world : 1);
  (void)e;
  return dist_allgather_host(&dummy, all.data(), sizeof(double));
}

}  // namespace qcs

using namespace qcs;

extern "C" {

int qcs_cuda_dist_unique_id(char id[128]) {
  static_assert(NCCL_UNIQUE_ID_BYTES == 128, "unique id size");
  ncclUniqueId uid;
  int rc = nccl_api::load();
  if (rc) return rc;
  rc = check_nccl(ncclGetUniqueId(&uid), "ncclGetUniqueId");
  if (rc) return rc;
  std::memcpy(id, &uid, 128);
  return QCS_CUDA_OK;
}

int qcs_cuda_dist_init(int rank, int world, const char id[128], int device) {
  DistContext &d = dist();
  if (d.active) return set_error(QCS_CUDA_ERR_INVALID, "communicator already initialised");
  if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world)
    return set_error(QCS_CUDA_ERR_INVALID, "world size must be a power of two (got %d)", world);
  int rc = nccl_api::load();
  if (rc) return rc;
  rc = check_cuda(cudaSetDevice(device), "cudaSetDevice");
  if (rc) return rc;
  ncclUniqueId uid;
  std::memcpy(&uid, id, 128);
  ncclComm_t comm = nullptr;
  rc = check_nccl(ncclCommInitRank(&comm, world, uid, rank), "ncclCommInitRank");
  if (rc) return rc;
  rc = check_cuda(cudaMalloc(&g_small_dev, kSmallBytes), "cudaMalloc(small)");
  if (rc) return rc;
  d.comm = comm;
  d.rank = rank;
  d.world = world;
  d.device = device;
  d.rank_bits = 0;
  while ((1 << d.rank_bits) < world) d.rank_bits++;
  d.active = world > 1;
  if (!d.active) {  // single rank: behave exactly like the non-distributed engine
    ncclCommDestroy(comm);
    d.comm = nullptr;
  }
  return QCS_CUDA_OK;
}

int qcs_cuda_dist_init_plan_only(int rank, int world) {
  DistContext &d = dist();
  if (d.active) return set_error(QCS_CUDA_ERR_INVALID, "communicator already initialised");
  if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world)
    return set_error(QCS_CUDA_ERR_INVALID, "world size must be a power of two (got %d)", world);
  d = DistContext();
  d.rank = rank;
  d.world = world;
  while ((1 << d.rank_bits) < world) d.rank_bits++;
  d.active = world > 1;
  return QCS_CUDA_OK;
}

int qcs_cuda_dist_finalize(void) {
  DistContext &d = dist();
  if (d.comm) release_peer_cache();
  if (d.comm) ncclCommDestroy((ncclComm_t)d.comm);
  if (g_small_dev) cudaFree(g_small_dev);
  g_small_dev = nullptr;
  d = DistContext();
  return QCS_CUDA_OK;
}

int qcs_cuda_dist_rank(void) { return dist().active ? dist().rank : 0; }
int qcs_cuda_dist_world(void) { return dist().active ? dist().world : 1; }

}  // extern "C"
