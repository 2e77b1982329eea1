// engine.h -- internal definition of qcs_cuda_engine (the handle behind the ABI).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "common.h"
#include "kernels.h"
#include "planner.h"

struct qcs_cuda_stats;

namespace qcs {

struct Options {
  Semantics sem = SEM_REFERENCE;
  bool fusion = true;
  bool dryrun = false;
  double pass_flops = 1000.0;  // FP64 work per amplitude a pass may fuse; measured flat above ~200 (scripts/budget_sweep.py): a pass costs max(memory, FP64), splitting it never helps
  int exchange = 0;  // 0: NCCL send/recv, 1: peer-memory swap kernel
  int tile_bits = 11;      // largest tile (10..12 bits) a pass may use; it runs on the smallest that fits.
                           // 11: measured best on the random circuits (12: 13 % fewer passes, each 25 % slower)
  double compute_bound_flops = 90.0;  // a pass with this much FP64 work per amplitude stops growing its tile at 10 bits
  int fixed_low = 5;       // positions every tile contains (contiguous global rows of 16 << fixed_low bytes)
  int min_fused_victim = 5;  // lowest local position a pass-carried swap may trade away (2..5; measured: a
                             // qubit arriving at an always-in-tile position saves a pass in a QFT but its
                             // 64..256-byte remote store runs cost 1.5-3 ms per swap: no net gain)
  bool fast_math = false;  // math=fast (opt-in, corrected semantics, ldg8 kernels): fused multiply-adds and
                           // controlled-phase fans collapsed into one factor per thread.  Amplitudes then
                           // agree with the reference to ~1e-15 relative (tested at 1e-12, the north star's
                           // bar) instead of bit for bit; the default (math=exact) stays bit-exact.
  bool reorder = true;     // math=fast, single GPU: commuting gates may trade places so that a pass follows its
                           // tile through several circuit layers (planner.h PlannerConfig::reorder)
  int reorder_segments = 8;
  bool peephole = true;    // drop exactly self-cancelling gate pairs from the queue (corrected semantics)
  bool fuse_swaps = true;  // fold position swaps into the stores of a pass (peer-memory path only)
  bool tile_search = true;  // plan again with smaller tile caps and keep the smallest that needs no more passes
  bool tile_search_heavy = true;  // ... and with FP64-heavy passes allowed to grow, kept when it saves a pass
  int fan_tables = 2;  // math=fast: runs of controlled phases with at least this many in-tile controls become
                       // per-thread tables (0 = never)
  int victim_policy = 1;   // which local position leaves when the queue pairs none of the candidates again:
                           // 1 the most recently paired (cyclic sweeps), 0 the least recently paired
  int remap_max = 3;       // position pairs one carrying pass may trade (1..3): an all-to-all among 2^k ranks
  bool fuse_argmax = true; // qc_find_most_likely_state folds its first reduction level into the pass it flushes
  bool lazy_init = true;   // qc_create writes nothing; see Engine::zero_ket_pending
  int remap_buffer = 1;    // 1 inplace (default): carrying passes overwrite the ranks' shards behind a per-tile
                           // handshake; 0 auto: into a second shard buffer when it fits (kernels.h
                           // SwapStore::out_of_place: no flags, no waits), 2 double: required.  Measured equal
                           // in speed on 2 and 8 GPUs (profiles/r2q_bench_8gpu.json vs r2p, r2r_carry_probe.log),
                           // so the default does not spend the memory.
  bool swap_bulk = true;   // such a pass hands the amplitudes that leave to TMA bulk stores (row-sized NVLink
                           // writes out of shared memory) instead of 16-byte st.global from the compute threads
  int tile_kernel = 3;  // 0: ldg (256 thr x 16 amps, plain loads), 1: tma16 (TMA, 256 x 16), 2: tma (TMA, 512 x 8), 3: ldg8 (512 thr x 8 amps, plain loads; default)
};

struct DistContext {
  bool active = false;
  int rank = 0;
  int world = 1;
  int rank_bits = 0;
  int device = 0;
  void *comm = nullptr;  // ncclComm_t
};

DistContext &dist();

struct Engine {
  int n = 0;        // logical qubits
  int nl = 0;       // local (per-shard) qubits
  uint64_t local_size = 0;
  uint64_t shard_base = 0;  // rank << nl
  Options opt;

  double2 *live = nullptr;     // reference: state->vector
  double2 *scratch = nullptr;  // reference: state->scratch_vector (lazily allocated)
  double2 *staging = nullptr;  // half-shard exchange buffer (multi-GPU, NCCL path)
  std::vector<double2 *> peer_live;  // every rank's state buffer mapped through CUDA IPC (P2P path)
  double2 *alt = nullptr;            // second shard buffer: carrying passes read `live`, write every rank's `alt`,
  std::vector<double2 *> peer_alt;   // and the two trade places (multi-GPU P2P path, when memory allows)
  cudaStream_t stream = nullptr;
  uint32_t *tile_flags = nullptr;         // one word per tile: handshake of passes that swap on the way out
  std::vector<uint32_t *> peer_flags;     // every rank's tile_flags, peer-mapped
  uint32_t swap_epoch = 0;
  uint32_t tile_flag_stride = 0;          // words per sending rank in tile_flags
  uint32_t *swap_abort_flag = nullptr;    // device word behind tile_flags: a CTA gave up waiting for its partner
  int *swap_status_host = nullptr;        // pinned: all-reduced abort words of the last swap-carrying pass
  bool swap_out_of_place = false;         // the carrying pass being launched writes `alt` (dist_fused_swap_args)
  bool swap_status_pending = false;       // a status copy is in flight on the stream
  bool poisoned = false;                  // a failed flush / swap left the amplitudes undefined: every later
                                          // call that needs them returns the error
  std::string poison_reason;
  bool zero_ket_pending = false;   // `live` has not been written since qc_create: the state is |0...0> by
                                   // definition.  The first fused pass synthesises its input (kernels.h
                                   // QCS_PASS_SYNTH_ZERO_KET); anything else that looks at `live` first
                                   // writes the state out (materialize, engine.cu)
  // qc_find_most_likely_state right behind a run of gates: the last pass of the flush it triggers also
  // leaves one (max |a|^2, index) candidate per tile (kernels.h QCS_PASS_ARGMAX)
  bool argmax_request = false;     // set by qcs_cuda_argmax around its flush
  bool argmax_done = false;        // the flush's last pass wrote argmax_tiles candidates
  uint64_t argmax_tiles = 0;
  double *argmax_tile_p = nullptr;       // device, one per tile of the smallest tile size
  long long *argmax_tile_idx = nullptr;
  bool carried_sum_valid = false;  // ws.result[RES_LOCAL_SUM_*] holds the sum of this shard's amplitudes
  ReduceWorkspace ws{};
  void *ws_slab = nullptr;       // one allocation backing every array of ws
  size_t ws_slab_bytes = 0;
  double2 *table_buf = nullptr;   // math=fast: thread-table fans of the pass about to run (device)
  double *u_dev = nullptr;
  long long *idx_dev = nullptr;
  int shots_cap = 0;

  // logical qubit -> physical position and back
  int perm[64];
  int inv_perm[64];

  std::vector<HostGate> queue;
  std::vector<PassPlan> last_plan;
  struct TraceEntry { double v[12]; };
  std::vector<TraceEntry> trace;  // dry-run only

  // statistics
  long long gates_submitted = 0, gates_executed = 0, passes = 0, kernel_launches = 0,
            segments = 0, remaps = 0;
  long long gates_cancelled = 0;  // dropped by the queue peephole
  long long multi_remaps = 0;  // carrying passes that traded more than one position pair
  long long out_of_place_remaps = 0;  // carrying passes that wrote the second buffer (no handshake)
  long long fused_swaps = 0;   // remaps that rode on a pass instead of getting a kernel of their own
  double algorithmic_bytes = 0, pass_bytes = 0, pass_ms = 0, exchange_bytes = 0, exchange_ms = 0;
  double pass_flops_per_amp = 0;  // planner's FP64 operation count per amplitude, summed over executed passes
  double fused_swap_bytes = 0;    // bytes per direction moved inside carrying passes
  double fused_swap_pass_ms = 0;  // device time of the passes that carried a swap (also in pass_ms)
  bool timing = false;
  // one event pair per launched pass (timing on): folded into pass_ms / fused_swap_pass_ms and, for the
  // passes of the most recent flush, into last_plan_ms (per-pass roofline figures of bench.py)
  struct PendingPass { cudaEvent_t begin, end; bool carries_swap; long long generation; int index; };
  std::vector<PendingPass> pending_pass_events;
  long long plan_generation = 0;        // bumped whenever last_plan is cleared
  std::vector<double> last_plan_ms;     // device time of last_plan[k], -1 where not measured
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_xchg_events;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t markers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// engine.cu
// a pooled buffer whose CUDA IPC handle other ranks may hold mapped: never cudaFree'd (it stays in the
// pool, whatever the pool's size limits say) until pool_unpin_all
void pool_pin(void *ptr);
void pool_release(void *ptr, size_t bytes);  // = pool_free, for dist.cu
void pool_unpin_all();
int set_error(int code, const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);

// dist.cu
// Swaps physical positions `lpos` (local) and `gpos` (global) of the shard
// layout with the partner rank (half of the shard crosses NVLink).
int dist_swap_positions(Engine &e, int lpos, int gpos);
int dist_allreduce_sum(Engine &e, double *dev_values, int count);
int dist_allreduce_max_i64(Engine &e, long long *dev_values, size_t count);
int dist_allgather_host(const void *mine, void *all, size_t bytes_each);
int dist_barrier(Engine &e);
int dist_open_peers(Engine &e);
bool dist_p2p_available(const Engine &e);
int dist_p2p_swap(Engine &e, cudaStream_t stream, int lpos, int gpos);
bool dist_fused_swap_args(Engine &e, int k, const int *lpos, const int *gpos, SwapStore &sw);
int dist_after_fused_swap(Engine &e);
int dist_check_fused_swaps(Engine &e);  // after a stream synchronisation: error if a handshake wait timed out
void dist_close_peers(Engine &e);

}  // namespace qcs

struct qcs_cuda_engine : public qcs::Engine {};
