// engine.cu -- the extern "C" ABI of include/qcs_cuda.h: device-resident state,
// deferred gate queue, flush through the fusion planner, measurement and
// sampling.  Everything the reference does on the host with state->vector
// (SURVEY.md section 8b lists the sites) funnels through here.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "../../../include/qcs_cuda.h"

namespace qcs {

// ------------------------------------------------------------------ errors
static thread_local char g_error[512] = "";

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return QCS_CUDA_OK;
  return set_error(e == cudaErrorMemoryAllocation ? QCS_CUDA_ERR_NOMEM : QCS_CUDA_ERR_CUDA,
                   "%s: %s", what, cudaGetErrorString(e));
}

#define CK(call)                                   \
  do {                                             \
    int rc_ = check_cuda((call), #call);           \
    if (rc_ != QCS_CUDA_OK) return rc_;            \
  } while (0)
#define RC(call)                                   \
  do {                                             \
    int rc_ = (call);                              \
    if (rc_ != QCS_CUDA_OK) return rc_;            \
  } while (0)

// ------------------------------------------------------------------ options
static std::map<std::string, std::string> &defaults() {
  static std::map<std::string, std::string> d;
  return d;
}
// engines may be created from several host threads (one circuit per thread): the option table and
// the buffer pool are the process-wide state they share
static std::mutex &globals_mutex() {
  static std::mutex m;
  return m;
}

static std::string option_value(const char *key) {
  {
    std::lock_guard<std::mutex> lock(globals_mutex());
    auto it = defaults().find(key);
    if (it != defaults().end()) return it->second;
  }
  std::string env = "QCS_CUDA_";
  for (const char *p = key; *p; p++) env += (char)toupper(*p);
  const char *v = std::getenv(env.c_str());
  return v ? std::string(v) : std::string();
}

static int load_options(Options &o) {
  std::string v = option_value("semantics");
  if (v == "corrected") o.sem = SEM_CORRECTED;
  else if (v.empty() || v == "reference") o.sem = SEM_REFERENCE;
  else return set_error(QCS_CUDA_ERR_INVALID, "semantics must be reference|corrected, got '%s'", v.c_str());
  v = option_value("fusion");
  if (v == "off" || v == "0") o.fusion = false;
  else if (v.empty() || v == "on" || v == "1") o.fusion = true;
  else return set_error(QCS_CUDA_ERR_INVALID, "fusion must be on|off, got '%s'", v.c_str());
  v = option_value("dryrun");
  o.dryrun = (v == "1" || v == "on");
  v = option_value("pass_flops");
  if (!v.empty()) o.pass_flops = std::atof(v.c_str());
  if (!(o.pass_flops > 0)) o.pass_flops = Options().pass_flops;
  v = option_value("tile_kernel");
  if (v == "ldg") o.tile_kernel = 0;
  else if (v == "tma16") o.tile_kernel = 1;
  else if (v == "tma" || v == "tma8") o.tile_kernel = 2;
  else if (v == "ldg8") o.tile_kernel = 3;
  else if (v.empty()) o.tile_kernel = Options().tile_kernel;
  else return set_error(QCS_CUDA_ERR_INVALID, "tile_kernel must be ldg|ldg8|tma|tma16, got '%s'", v.c_str());
  v = option_value("exchange");
  if (v == "nccl") o.exchange = 0;
  else if (v.empty() || v == "p2p") o.exchange = 1;
  else return set_error(QCS_CUDA_ERR_INVALID, "exchange must be p2p|nccl, got '%s'", v.c_str());
  v = option_value("fuse_swaps");
  o.fuse_swaps = !(v == "off" || v == "0");
  v = option_value("tile_search");
  o.tile_search = !(v == "off" || v == "0");
  o.tile_search_heavy = o.tile_search && v != "caps";  // "caps": smaller caps only, FP64-heavy passes never grow
  v = option_value("fan_tables");
  // "off" | "on" | the fewest in-tile controls a run needs to become a table (on = 2)
  o.fan_tables = (v == "off" || v == "0") ? 0 : (v.empty() || v == "on") ? 2 : std::max(2, std::atoi(v.c_str()));
  v = option_value("remap_buffer");
  if (v == "auto") o.remap_buffer = 0;
  else if (v.empty() || v == "inplace") o.remap_buffer = 1;
  else if (v == "double") o.remap_buffer = 2;
  else return set_error(QCS_CUDA_ERR_INVALID, "remap_buffer must be auto|inplace|double, got '%s'", v.c_str());
  v = option_value("victim_policy");
  if (v == "lru") o.victim_policy = 0;
  else if (v.empty() || v == "mru") o.victim_policy = 1;
  else if (v == "mru2") o.victim_policy = 2;  // experimental: mru, but not from the last pass in front of the remap
  else return set_error(QCS_CUDA_ERR_INVALID, "victim_policy must be mru|lru|mru2, got '%s'", v.c_str());
  v = option_value("remap_max");
  if (!v.empty()) o.remap_max = std::atoi(v.c_str());
  if (o.remap_max < 1 || o.remap_max > QCS_MAX_REMAP) o.remap_max = Options().remap_max;
  v = option_value("fuse_argmax");
  o.fuse_argmax = !(v == "off" || v == "0");
  v = option_value("lazy_init");
  o.lazy_init = !(v == "off" || v == "0");
  v = option_value("swap_store");
  if (v == "thread" || v == "st") o.swap_bulk = false;
  else if (v.empty() || v == "bulk") o.swap_bulk = true;
  else return set_error(QCS_CUDA_ERR_INVALID, "swap_store must be bulk|thread, got '%s'", v.c_str());
  v = option_value("tile_bits");
  if (!v.empty()) o.tile_bits = std::atoi(v.c_str());
  if (o.tile_bits < QCS_MIN_TILE_BITS || o.tile_bits > QCS_TILE_BITS) o.tile_bits = Options().tile_bits;
  v = option_value("compute_bound_flops");
  if (!v.empty()) o.compute_bound_flops = std::atof(v.c_str());
  if (!(o.compute_bound_flops > 0)) o.compute_bound_flops = Options().compute_bound_flops;
  v = option_value("fixed_low");
  if (!v.empty()) o.fixed_low = std::atoi(v.c_str());
  if (o.fixed_low < 1 || o.fixed_low > QCS_LANE_BITS) o.fixed_low = Options().fixed_low;
  v = option_value("min_fused_victim");
  if (!v.empty()) o.min_fused_victim = std::atoi(v.c_str());
  if (o.min_fused_victim < 0 || o.min_fused_victim > QCS_LANE_BITS) o.min_fused_victim = Options().min_fused_victim;
  v = option_value("peephole");
  o.peephole = !(v == "off" || v == "0");
  v = option_value("math");
  if (v == "fast") o.fast_math = true;
  else if (v.empty() || v == "exact") o.fast_math = false;
  else return set_error(QCS_CUDA_ERR_INVALID, "math must be exact|fast, got '%s'", v.c_str());
  v = option_value("reorder");
  o.reorder = !(v == "off" || v == "0");
  v = option_value("reorder_segments");
  if (!v.empty()) o.reorder_segments = std::atoi(v.c_str());
  if (o.reorder_segments < 1 || o.reorder_segments > QCS_MAX_PASS_SEGMENTS) o.reorder_segments = Options().reorder_segments;
  if (o.fast_math && (o.sem != SEM_CORRECTED || (o.tile_kernel != 3 && o.tile_kernel != 0)))
    return set_error(QCS_CUDA_ERR_INVALID,
                     "math=fast needs semantics=corrected and tile_kernel=ldg8|ldg (bug-compatible `reference` "
                     "semantics are bit-exact by definition)");
  return QCS_CUDA_OK;
}

// ------------------------------------------------------------------ device-buffer pool
// cudaMalloc / cudaFree of a 16 GiB state cost 10-350 ms each (measured); programs in the
// reference's style create and destroy circuits freely (every test does), so freed buffers are
// kept (at most kPoolSlots of them) and handed back to the next qc_create of the same size.
// QCS_CUDA_POOL=0 disables; an allocation failure trims the pool and retries.
namespace {
struct PoolEntry { void *ptr; size_t bytes; int device; };
std::vector<PoolEntry> g_pool;
std::vector<void *> g_pinned;  // see pool_pin (engine.h)
bool is_pinned(void *p) { return std::find(g_pinned.begin(), g_pinned.end(), p) != g_pinned.end(); }
constexpr size_t kPoolSlots = 6;
constexpr size_t kPoolBytes = (size_t)40 << 30;  // never sit on more than 40 GiB of freed buffers
bool pool_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *v = std::getenv("QCS_CUDA_POOL");
    on = (v && (v[0] == '0')) ? 0 : 1;
  }
  return on == 1;
}
void pool_trim_locked() {
  std::vector<PoolEntry> keep;
  for (auto &p : g_pool) {
    if (is_pinned(p.ptr)) keep.push_back(p);
    else cudaFree(p.ptr);
  }
  g_pool.swap(keep);
}
cudaError_t pool_alloc(void **out, size_t bytes) {
  std::lock_guard<std::mutex> lock(globals_mutex());
  int dev = 0;
  cudaGetDevice(&dev);
  for (size_t i = 0; i < g_pool.size(); i++) {
    if (g_pool[i].bytes == bytes && g_pool[i].device == dev) {
      *out = g_pool[i].ptr;
      g_pool.erase(g_pool.begin() + (long)i);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaErrorMemoryAllocation && !g_pool.empty()) {
    cudaGetLastError();
    pool_trim_locked();
    e = cudaMalloc(out, bytes);
  }
  return e;
}
// a buffer of exactly this size if the pool holds one (no cudaMalloc)
bool pool_take_cached(void **out, size_t bytes) {
  std::lock_guard<std::mutex> lock(globals_mutex());
  int dev = 0;
  cudaGetDevice(&dev);
  for (size_t i = 0; i < g_pool.size(); i++)
    if (g_pool[i].bytes == bytes && g_pool[i].device == dev) {
      *out = g_pool[i].ptr;
      g_pool.erase(g_pool.begin() + (long)i);
      return true;
    }
  return false;
}
void pool_free(void *ptr, size_t bytes) {
  if (!ptr) return;
  std::lock_guard<std::mutex> lock(globals_mutex());
  const bool pinned = is_pinned(ptr);
  if (!pinned && (!pool_enabled() || bytes < (1u << 20) || bytes > kPoolBytes)) {
    cudaFree(ptr);
    return;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  g_pool.push_back(PoolEntry{ptr, bytes, dev});
  size_t total = 0, unpinned = 0;
  for (auto &p : g_pool) {
    if (is_pinned(p.ptr)) continue;
    total += p.bytes;
    unpinned++;
  }
  for (size_t i = 0; i < g_pool.size() && (unpinned > kPoolSlots || total > kPoolBytes);) {  // evict oldest first
    if (is_pinned(g_pool[i].ptr)) {
      i++;
      continue;
    }
    total -= g_pool[i].bytes;
    unpinned--;
    cudaFree(g_pool[i].ptr);
    g_pool.erase(g_pool.begin() + (long)i);
  }
}
}  // namespace

void pool_release(void *ptr, size_t bytes) { pool_free(ptr, bytes); }

void pool_pin(void *ptr) {
  std::lock_guard<std::mutex> lock(globals_mutex());
  if (ptr && !is_pinned(ptr)) g_pinned.push_back(ptr);
}

void pool_unpin_all() {
  std::lock_guard<std::mutex> lock(globals_mutex());
  g_pinned.clear();
  pool_trim_locked();
}

// ------------------------------------------------------------------ events / stats
static cudaEvent_t get_event(Engine &e) {
  if (!e.event_pool.empty()) {
    cudaEvent_t ev = e.event_pool.back();
    e.event_pool.pop_back();
    return ev;
  }
  cudaEvent_t ev = nullptr;
  cudaEventCreate(&ev);
  return ev;
}

static void fold_events(Engine &e) {
  if (e.pending_pass_events.empty() && e.pending_xchg_events.empty()) return;
  cudaStreamSynchronize(e.stream);
  for (auto &pp : e.pending_pass_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pp.begin, pp.end) == cudaSuccess) {
      e.pass_ms += ms;
      if (pp.carries_swap) e.fused_swap_pass_ms += ms;
      if (pp.generation == e.plan_generation && pp.index >= 0 && (size_t)pp.index < e.last_plan.size()) {
        if (e.last_plan_ms.size() < e.last_plan.size()) e.last_plan_ms.resize(e.last_plan.size(), -1.0);
        e.last_plan_ms[(size_t)pp.index] = ms;
      }
    }
    e.event_pool.push_back(pp.begin);
    e.event_pool.push_back(pp.end);
  }
  e.pending_pass_events.clear();
  for (auto &pr : e.pending_xchg_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) e.exchange_ms += ms;
    e.event_pool.push_back(pr.first);
    e.event_pool.push_back(pr.second);
  }
  e.pending_xchg_events.clear();
}

// ------------------------------------------------------------------ lazy |0...0>
// q_state_init zeroes the vector and sets amplitude 0 (reference src/q_state.c:93-99).  That is 16
// bytes written per amplitude which the first pass would read straight back, so qc_create only
// NOTES that the state is |0...0>; whoever needs the bytes in memory calls this first.
static int materialize(Engine &e) {
  if (!e.zero_ket_pending) return QCS_CUDA_OK;
  e.zero_ket_pending = false;
  if (e.opt.dryrun) return QCS_CUDA_OK;
  const bool owns_zero = !dist().active || dist().rank == 0;
  CK(launch_init_state(e.live, e.local_size, owns_zero, e.stream));
  e.kernel_launches++;
  return QCS_CUDA_OK;
}

// ------------------------------------------------------------------ scratch buffer
static int ensure_scratch(Engine &e) {
  if (e.scratch || e.opt.dryrun) return QCS_CUDA_OK;
  CK(pool_alloc((void **)&e.scratch, e.local_size * sizeof(double2)));
  // q_state_init zeroes both buffers (reference src/q_state.c:93-96)
  CK(launch_init_state(e.scratch, e.local_size, false, e.stream));
  e.kernel_launches++;
  return QCS_CUDA_OK;
}

// ------------------------------------------------------------------ executing gates
static PhysGate to_phys(const Engine &e, const HostGate &g) {
  PhysGate p;
  p.c = classify_gate(g.m, g.control >= 0, e.opt.sem);
  p.tpos = e.perm[g.target];
  p.cpos = g.control >= 0 ? e.perm[g.control] : -1;
  return p;
}

static bool is_pairing_kind(uint8_t k) {
  return k == GK_PAIR_GENERIC || k == GK_PAIR_REAL || k == GK_PAIR_HSYM || k == GK_PAIR_SWAP;
}

// SURVEY.md section 8(d) byte accounting for one gate executed on its own.
static double gate_bytes(const Engine &e, const PhysGate &g) {
  const double amps = (double)e.local_size;
  if (g.c.kind == GK_NOP) return 0.0;
  double frac = 1.0;
  if (g.cpos >= 0) frac *= 0.5;
  if (g.c.kind == GK_DIAG) {
    int touched = ((g.c.flags & GF_D0_IDENT) ? 0 : 1) + ((g.c.flags & GF_D1_IDENT) ? 0 : 1);
    frac *= 0.5 * touched;
  } else if (g.c.flags & GF_ROW0_ONLY) {
    frac *= 0.75;  // read both members, write one
  }
  return 32.0 * amps * frac;
}

static void trace_gates(Engine &e, const std::vector<PhysGate> &gates) {
  for (const PhysGate &g : gates) {
    Engine::TraceEntry t{};
    t.v[0] = 1.0;
    t.v[1] = (double)(g.c.kind | (g.c.flags << 8));
    t.v[2] = (double)g.tpos;
    t.v[3] = (double)g.cpos;
    for (int k = 0; k < 8; k++) t.v[4 + k] = g.c.m[k];
    e.trace.push_back(t);
  }
}

// smallest tile the selected kernel variant runs on (only ldg8 is instantiated below 12 bits)
static int min_tile_bits(const Engine &e) {
  if (e.opt.tile_kernel == 3 || e.opt.tile_kernel == 0) return std::min(QCS_MIN_TILE_BITS, e.opt.tile_bits);
  return QCS_TILE_BITS;  // the TMA-staged kernels exist for 12-bit tiles only
}

// whole_batch: every planned pass is launched before anything else happens to the state (run_local), so
// the passes may take the batch's gates out of order where that is exact (PlannerConfig::reorder_exact);
// the sharded in-order path launches a prefix of the plan and counts gates by position: in order only.
static std::vector<PassPlan> plan_batch(const Engine &e, const std::vector<PhysGate> &gates, bool whole_batch = false) {
  PlannerConfig cfg;
  cfg.n_local = e.nl;
  cfg.rank_bits = e.n - e.nl;
  cfg.shard_base = e.shard_base;
  cfg.sem = e.opt.sem;
  cfg.pass_flops_budget = e.opt.pass_flops;
  cfg.direct_io = true;
  cfg.reg_bits = (e.opt.tile_kernel >= 2) ? 3 : 4;
  cfg.fixed_low = e.opt.fixed_low;
  cfg.compute_bound_flops = e.opt.compute_bound_flops;
  cfg.fast_math = e.opt.fast_math;
  // (sharded engines: run_range_reordered keeps track of which gates a pass took)
  cfg.reorder = e.opt.fast_math && e.opt.reorder;
  cfg.reorder_exact = whole_batch && !e.opt.fast_math && e.opt.reorder && e.opt.sem == SEM_CORRECTED;
  cfg.reorder_segments = e.opt.reorder_segments;
  cfg.thread_tables = e.opt.fan_tables;
  cfg.tile_bits_min = min_tile_bits(e);
  cfg.tile_bits_max = std::max(cfg.tile_bits_min, std::min(e.opt.tile_bits, e.nl));
  std::vector<PassPlan> best = plan_passes(gates, cfg);
  // Larger tiles exist to save passes.  When a smaller cap needs no more of them, its plan wins: the
  // same gates spread evenly over more CTAs per SM (30-qubit QFT, 5 passes either way: targets
  // 10+5+5+5+5 on 10-bit tiles against 11+6+6+6+1 with 11 allowed -- 62.3 vs 63.6 ms bit-exact, 34.4 vs
  // 36.4 ms fast; the random circuits keep their 11-bit plans, which are passes shorter).  Planning is
  // repeated once per smaller cap, so only for queues where that is cheap.
  if (e.opt.tile_search && gates.size() <= 1024) {
    for (int cap = cfg.tile_bits_max - 1; cap >= cfg.tile_bits_min; cap--) {
      PlannerConfig smaller = cfg;
      smaller.tile_bits_max = cap;
      std::vector<PassPlan> alt = plan_passes(gates, smaller);
      if (alt.size() <= best.size()) best.swap(alt);
    }
    // ... and FP64-heavy passes stay on the smallest tiles (compute_bound_flops) unless letting them
    // grow saves a whole pass: a 31..33-qubit QFT shard needs 10 + 5 + 5 + 5 + 5 + 1 targets' worth of
    // passes on 10-bit tiles and 10 + 6 + 6 + 6 + 3 with 11 allowed
    if (!cfg.fast_math && e.opt.tile_search_heavy)
      for (int cap = cfg.tile_bits_max; cap > cfg.tile_bits_min; cap--) {
        PlannerConfig heavy = cfg;
        heavy.tile_bits_max = cap;
        heavy.compute_bound_flops = 1e30;
        std::vector<PassPlan> alt = plan_passes(gates, heavy);
        if (alt.size() < best.size()) best.swap(alt);
      }
  }
  return best;
}

// Launches planned passes [first, last); `swap` (may be null) rides on the stores of the last one.
// (n_swaps position pairs swap_lpos[i] <-> swap_gpos[i]: dist_fused_swap_args filled in the ranks' addresses)
static int launch_passes(Engine &e, const std::vector<PassPlan> &plan, size_t first, size_t last,
                         const SwapStore *swap, int n_swaps = 0, const int *swap_lpos = nullptr,
                         const int *swap_gpos = nullptr) {
  const bool timed = e.timing && !e.opt.dryrun;
  for (size_t k = first; k < last; k++) {
    const PassPlan &p = plan[k];
    const bool carries = swap && k + 1 == last;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (timed) {
      ev0 = get_event(e);
      ev1 = get_event(e);
      cudaEventRecord(ev0, e.stream);
    }
    uint32_t pass_flags = 0;
    if (e.zero_ket_pending) {
      if (e.opt.tile_kernel == 0 || e.opt.tile_kernel == 3) {
        pass_flags = QCS_PASS_SYNTH_ZERO_KET;  // this pass makes its own input
        e.zero_ket_pending = false;
      } else {
        RC(materialize(e));  // the TMA-staged kernels always read
      }
    }
    PassExtras extras{};
    if (e.argmax_request && k + 1 == last && !swap && !e.opt.dryrun) {
      pass_flags |= QCS_PASS_ARGMAX;
      extras.argmax_p = e.argmax_tile_p;
      extras.argmax_idx = e.argmax_tile_idx;
      e.argmax_done = true;
      e.argmax_tiles = e.local_size >> p.params.tile_bits;
    }
    if (!e.opt.dryrun && !p.thread_tables.empty()) {
      // math=fast: the pass's thread-table fans.  One device buffer serves every pass: the copy is ordered
      // on the stream behind the previous pass's kernel (the source is pageable host memory, staged by the
      // driver before cudaMemcpyAsync returns, so `plan` may die before the copy runs)
      const size_t cap = (size_t)QCS_MAX_PASS_FANS * sizeof(double2) << (QCS_TILE_BITS - 3);
      if (!e.table_buf) CK(cudaMalloc((void **)&e.table_buf, cap));
      const size_t bytes = p.thread_tables.size() * sizeof(double);
      if (bytes > cap) return set_error(QCS_CUDA_ERR_CUDA, "internal: thread tables exceed their buffer");
      CK(cudaMemcpyAsync(e.table_buf, p.thread_tables.data(), bytes, cudaMemcpyHostToDevice, e.stream));
      extras.thread_tables = e.table_buf;
    }
    if (!e.opt.dryrun) {
      if (carries) {
        SwapStore sw = *swap;
        sw.in_tile = 0;
        sw.n_out = 0;
        for (uint32_t i = 0; i < sw.k; i++) {
          bool in_tile = false;
          for (int pos : p.tile_positions)
            if ((uint32_t)pos == sw.lpos[i]) in_tile = true;
          if (in_tile) {
            sw.in_tile |= 1u << i;
            continue;
          }
          // which bit of the tile NUMBER this position is; kept sorted ascending (kernels.h)
          const uint32_t bit = (uint32_t)__builtin_popcountll(p.params.nontile_mask & ((1ull << sw.lpos[i]) - 1ull));
          uint32_t j = sw.n_out++;
          while (j > 0 && sw.out_tile_bit[j - 1] > bit) {
            sw.out_tile_bit[j] = sw.out_tile_bit[j - 1];
            sw.out_pair[j] = sw.out_pair[j - 1];
            j--;
          }
          sw.out_tile_bit[j] = bit;
          sw.out_pair[j] = i;
        }
        sw.bulk = e.opt.swap_bulk ? 1u : 0u;
        sw.row_bits = 0;
        while (sw.row_bits < (uint32_t)p.params.tile_bits && p.params.tile_pos[sw.row_bits] == sw.row_bits) sw.row_bits++;
        CK(launch_fused_pass(e.live, p.params, e.nl, e.stream, e.opt.tile_kernel, &sw, e.opt.fast_math, pass_flags, &extras));
        RC(dist_after_fused_swap(e));
      } else {
        CK(launch_fused_pass(e.live, p.params, e.nl, e.stream, e.opt.tile_kernel, nullptr, e.opt.fast_math, pass_flags, &extras));
      }
    }
    e.passes++;
    e.kernel_launches++;
    e.segments += p.params.n_segments;
    e.gates_executed += p.params.n_gates - p.n_fan_headers;
    const double bytes = (pass_flags & QCS_PASS_SYNTH_ZERO_KET ? 16.0 : 32.0) * (double)e.local_size;
    e.algorithmic_bytes += bytes;
    e.pass_bytes += bytes;
    e.pass_flops_per_amp += p.flops_per_amp;
    e.last_plan.push_back(p);
    if (carries) {
      e.last_plan.back().n_swaps = n_swaps;
      for (int i = 0; i < n_swaps && i < 3; i++) {
        e.last_plan.back().swap_lpos[i] = swap_lpos[i];
        e.last_plan.back().swap_gpos[i] = swap_gpos[i];
      }
    }
    if (timed) {
      cudaEventRecord(ev1, e.stream);
      e.pending_pass_events.push_back(Engine::PendingPass{ev0, ev1, carries, e.plan_generation,
                                                          (int)e.last_plan.size() - 1});
      if (e.pending_pass_events.size() > 512) fold_events(e);
    }
  }
  return QCS_CUDA_OK;
}

// Runs a list of physical gates whose pairing targets are all local.
static int run_local(Engine &e, const std::vector<PhysGate> &gates) {
  if (gates.empty()) return QCS_CUDA_OK;
  if (e.opt.dryrun) trace_gates(e, gates);
  const bool fused = e.opt.fusion && e.nl >= min_tile_bits(e);
  if (fused) {
    std::vector<PassPlan> plan = plan_batch(e, gates, true);
    RC(launch_passes(e, plan, 0, plan.size(), nullptr));
  } else {
    RC(materialize(e));
    for (const PhysGate &g : gates) {
      if (g.c.kind == GK_NOP) continue;
      DGate dg;
      std::memset(&dg, 0, sizeof(dg));
      std::memcpy(dg.m, g.c.m, sizeof(dg.m));
      dg.kind = g.c.kind;
      dg.flags = g.c.flags;
      dg.tpos = (int8_t)g.tpos;
      dg.cpos = (int8_t)g.cpos;
      dg.op = 0;  // the per-gate kernels read kind / flags, not the fused interpreter's case label
      dg.csel = dg.tsel = 0xFF;
      if (!e.opt.dryrun) CK(launch_simple_gate(e.live, dg, e.nl, e.shard_base, e.stream));
      e.kernel_launches++;
      e.gates_executed++;
      e.algorithmic_bytes += gate_bytes(e, g);
    }
  }
  return QCS_CUDA_OK;
}

// Picks the local position to trade for a global one at queue index `from`: the position whose
// next use as a pairing target lies farthest ahead.  Among equally idle ones, the one whose last
// pairing use since `since` lies farthest BACK (the swap may then ride on an earlier pass), then
// the highest (long contiguous rows stay intact).  Positions below `lowest` are never traded: 5 for
// stand-alone swaps (positions 0..4 are the 512-byte rows the swap kernels move); option
// min_fused_victim (2..5) when the swap rides on a pass's stores -- a qubit that arrives at one of
// the positions every tile contains can be paired in ANY later pass (QFT: the three incoming
// qubits join the last pass instead of needing one of their own), at the price of 64..256-byte
// instead of 512-byte remote store runs.
struct VictimRanking {
  std::vector<int> order;       // local positions, best victim first
  std::vector<long> next_use;   // per position (local AND global): queue index of the next pairing use at or
                                // after `from` of the qubit sitting there, q.size() + 1 if none
};
// mru_before: among positions the queue never pairs again, one last paired at or after this queue index
// ranks behind the others (run_range passes the first gate of the LAST pass in front of the remap: a
// victim paired there pins the remap to that pass, and the gate that needed the remap -- often the
// last of a QFT -- then gets a pass of its own; a victim from the pass before lets the last pass be
// planned again behind the remap, with that gate in it).
static VictimRanking rank_victims(const Engine &e, const std::vector<HostGate> &q, size_t from, size_t since,
                                  int lowest, long mru_before = -1) {
  VictimRanking vr;
  vr.next_use.assign(e.n, (long)q.size() + 1);
  std::vector<long> last_use(e.n, -1);
  for (size_t i = q.size(); i-- > since;) {
    Classified c = classify_gate(q[i].m, q[i].control >= 0, e.opt.sem);
    if (!is_pairing_kind(c.kind)) continue;
    int pos = e.perm[q[i].target];
    if (i >= from) vr.next_use[pos] = (long)i;
    else if (last_use[pos] < 0) last_use[pos] = (long)i;
  }
  for (int pos = e.nl - 1; pos >= lowest && pos >= 0; pos--) vr.order.push_back(pos);
  // Farthest next use first (Belady).  Ties only arise between positions the queue never pairs again;
  // among those: one it has not paired at all in this window, then -- the queue says nothing about
  // the future, but circuits sweep their qubits layer after layer -- the MOST recently paired one
  // (a cyclic sweep comes back to it last; evicting the least recently paired one instead costs a
  // repeated QFT a second remap per round, stand-alone at that: nothing runs before the first gate of
  // the next round for it to ride on); then positions every tile contains (only offered when
  // lowest < 5: whatever arrives there is pairable in every later pass); then the highest.
  const bool mru = e.opt.victim_policy >= 1;
  std::stable_sort(vr.order.begin(), vr.order.end(), [&](int a, int b) {
    if (vr.next_use[a] != vr.next_use[b]) return vr.next_use[a] > vr.next_use[b];
    if (last_use[a] != last_use[b]) {
      if (!mru) return last_use[a] < last_use[b];
      if ((last_use[a] < 0) != (last_use[b] < 0)) return last_use[a] < 0;
      if (mru_before >= 0) {
        const bool late_a = last_use[a] >= mru_before, late_b = last_use[b] >= mru_before;
        if (late_a != late_b) return late_b;
      }
      return last_use[a] > last_use[b];
    }
    return (a < QCS_LANE_BITS) && !(b < QCS_LANE_BITS);
  });
  if (vr.order.empty()) vr.order.push_back(e.nl - 1);
  return vr;
}

static int pick_victim(const Engine &e, const std::vector<HostGate> &q, size_t from, size_t since,
                       int lowest) {
  return rank_victims(e, q, from, since, lowest).order[0];
}

static void note_swap(Engine &e, int lpos, int gpos) {
  if (e.opt.dryrun) {
    Engine::TraceEntry t{};
    t.v[0] = 2.0;
    t.v[1] = (double)lpos;
    t.v[2] = (double)gpos;
    e.trace.push_back(t);
  }
  const int ql = e.inv_perm[lpos], qg = e.inv_perm[gpos];
  e.perm[ql] = gpos;
  e.perm[qg] = lpos;
  e.inv_perm[lpos] = qg;
  e.inv_perm[gpos] = ql;
  e.remaps++;
  e.exchange_bytes += 8.0 * (double)e.local_size;
}

static int swap_positions(Engine &e, int lpos, int gpos) {
  RC(materialize(e));
  if (!e.opt.dryrun) RC(dist_swap_positions(e, lpos, gpos));
  note_swap(e, lpos, gpos);
  if (e.opt.dryrun) {  // plan-only engines list stand-alone swaps between the passes (qcs_cuda_last_plan_swap)
    PassPlan marker;
    std::memset(&marker.params, 0, sizeof(marker.params));
    marker.n_gates_api = 0;
    marker.flops_per_amp = 0.0;
    marker.n_swaps = 1;
    marker.swap_lpos[0] = lpos;
    marker.swap_gpos[0] = gpos;
    e.last_plan.push_back(marker);
  }
  return QCS_CUDA_OK;
}

// A swap can ride on the stores of a fused pass (SwapStore) when passes run through the plain-load
// tile kernels and the partner's memory is mapped; plan-only engines follow the same schedule.
static bool can_fuse_swaps(const Engine &e) {
  if (!e.opt.fusion || e.nl < min_tile_bits(e) || e.opt.sem != SEM_CORRECTED || e.opt.exchange != 1 ||
      !e.opt.fuse_swaps)
    return false;
  if (e.opt.tile_kernel != 3) return false;  // the remap machinery is compiled into the ldg8 kernels only
  return e.opt.dryrun ? dist().active : (dist_p2p_available(e) && e.tile_flags != nullptr);
}

// Executes queue[begin, end).  A pairing gate whose target sits on a global position needs a
// position swap first.  The swap does not get a kernel of its own when it can be avoided: it is
// folded into the stores of a pass that runs anyway -- the EARLIEST pass after the victim
// position's last pairing use, so that the NVLink transfer hides behind as much arithmetic as
// possible and later swaps find later passes free.
static bool reorder_enabled(const Engine &e) {
  return e.opt.fast_math && e.opt.reorder && e.opt.fusion && e.nl >= min_tile_bits(e);
}

// math=fast with reordering on a sharded engine.  The planner gets the WHOLE pending queue: it runs
// everything the current layout allows, several layers deep where qubits do not depend on a global
// one, and leaves out the pairing gates on global positions and what hangs on them (planner.h).
// Then one global qubit is swapped in.  Same policy as run_range below -- the swap rides on the
// stores of the earliest pass after which nothing pairs on the victim position -- but "the gates that
// have run" is a set, not a prefix: they are taken out of the pending list by index, and what is left
// (in queue order) is planned again under the new layout.
static int run_range_reordered(Engine &e, const std::vector<HostGate> &q, size_t begin, size_t end) {
  const bool fuse = can_fuse_swaps(e);
  std::vector<HostGate> pend(q.begin() + (long)begin, q.begin() + (long)end);
  while (!pend.empty()) {
    std::vector<PhysGate> batch;
    for (const HostGate &g : pend) batch.push_back(to_phys(e, g));
    std::vector<PassPlan> plan = plan_batch(e, batch);
    std::vector<char> planned(pend.size(), 0);
    for (const PassPlan &p : plan)
      for (int id : p.api_ids) planned[(size_t)id] = 1;
    std::vector<HostGate> rest;  // what cannot run under this layout, in queue order
    for (size_t i = 0; i < pend.size(); i++)
      if (!planned[i]) rest.push_back(pend[i]);
    if (rest.empty()) {
      if (e.opt.dryrun) trace_gates(e, batch);
      return launch_passes(e, plan, 0, plan.size(), nullptr);
    }
    // the first gate left behind pairs on a global position (anything else would have been taken)
    const PhysGate blocker = to_phys(e, rest[0]);
    if (!(is_pairing_kind(blocker.c.kind) && blocker.tpos >= e.nl))
      return set_error(QCS_CUDA_ERR_CUDA, "internal: scheduler left a runnable gate behind");
    const int gpos = blocker.tpos;
    const int victim = pick_victim(e, rest, 0, 0, fuse ? e.opt.min_fused_victim : QCS_LANE_BITS);
    size_t chosen = 0;  // last pass that pairs on the victim position (the first pass if none does)
    for (size_t k = 0; k < plan.size(); k++) {
      const PassParams &pp = plan[k].params;
      for (int gi = 0; gi < pp.n_gates; gi++)
        if (!(pp.gate[gi].flags & GF_FAN_HEADER) && is_pairing_kind(pp.gate[gi].kind) &&
            pp.gate[gi].tpos == victim)
          chosen = k;
    }
    SwapStore sw{};
    const bool ride = fuse && !plan.empty() && (e.opt.dryrun || dist_fused_swap_args(e, 1, &victim, &gpos, sw));
    const size_t n_run = plan.empty() ? 0 : (ride ? chosen + 1 : plan.size());
    std::vector<char> ran(pend.size(), 0);
    std::vector<PhysGate> ran_gates;
    for (size_t k = 0; k < n_run; k++)
      for (int id : plan[k].api_ids) {
        ran[(size_t)id] = 1;
        ran_gates.push_back(batch[(size_t)id]);
      }
    if (e.opt.dryrun) trace_gates(e, ran_gates);
    if (ride) {
      RC(launch_passes(e, plan, 0, n_run, &sw, 1, &victim, &gpos));
      note_swap(e, victim, gpos);
      e.fused_swap_bytes += 8.0 * (double)e.local_size;
      e.fused_swaps++;
    } else {
      RC(launch_passes(e, plan, 0, n_run, nullptr));  // nothing to ride on, or no peer mapping
      RC(swap_positions(e, victim, gpos));
    }
    size_t w = 0;
    for (size_t i = 0; i < pend.size(); i++)
      if (!ran[i]) pend[w++] = pend[i];
    pend.resize(w);
  }
  return QCS_CUDA_OK;
}

static int run_range(Engine &e, const std::vector<HostGate> &q, size_t begin, size_t end) {
  if (dist().active && reorder_enabled(e)) return run_range_reordered(e, q, begin, end);
  const bool fuse = can_fuse_swaps(e);
  size_t i = begin;
  while (i < end) {
    size_t x = i;  // first pairing gate on a global position under the current layout
    std::vector<PhysGate> batch;
    for (; x < end; x++) {
      PhysGate p = to_phys(e, q[x]);
      if (is_pairing_kind(p.c.kind) && p.tpos >= e.nl) break;
      batch.push_back(p);
    }
    if (x == end) return run_local(e, batch);
    const int gpos = e.perm[q[x].target];
    std::vector<PassPlan> plan;
    long last_pass_from = -1;  // queue index of the first gate of the batch plan's last pass
    if (fuse) {
      plan = plan_batch(e, batch);
      if (plan.size() > 1) {
        last_pass_from = (long)i;
        for (size_t k = 0; k + 1 < plan.size(); k++) last_pass_from += plan[k].n_gates_api;
      }
    }
    // (victim_policy=mru2: see rank_victims; measured in the planner only -- on 8 ranks the layouts it leaves
    // behind cost a second carrying pass in two QFT steps out of five)
    const VictimRanking vr = rank_victims(e, q, x, i, fuse ? e.opt.min_fused_victim : QCS_LANE_BITS,
                                          e.opt.victim_policy == 2 ? last_pass_from : -1);
    const int victim = vr.order[0];
    if (!fuse) {
      RC(run_local(e, batch));
      RC(swap_positions(e, victim, gpos));
      i = x;
      continue;
    }
    // Remap of k positions at once: the pass that carries this swap can bring in
    // OTHER global-resident qubits too -- an all-to-all among 2^k ranks moves (1 - 2^-k) of the shard
    // where k separate swaps move k halves, and costs one carrying pass instead of k.  A qubit joins
    // when it will be paired before the victim it displaces would be (evicting that victim now
    // instead of then loses nothing); candidates in order of urgency, victims in order of idleness.
    int lpos_k[QCS_MAX_REMAP] = {victim, -1, -1}, gpos_k[QCS_MAX_REMAP] = {gpos, -1, -1};
    int k_pairs = 1;
    {
      std::vector<int> others;
      for (int g = e.nl; g < e.n; g++)
        if (g != gpos) others.push_back(g);
      std::sort(others.begin(), others.end(), [&](int a, int b) { return vr.next_use[a] < vr.next_use[b]; });
      const int kmax = std::min(std::min(e.opt.remap_max, QCS_MAX_REMAP), e.n - e.nl);
      size_t vi = 1;
      for (int g : others) {
        if (k_pairs >= kmax || vi >= vr.order.size()) break;
        const int v = vr.order[vi];
        if (vr.next_use[g] < vr.next_use[v]) {
          lpos_k[k_pairs] = v;
          gpos_k[k_pairs] = g;
          k_pairs++;
          vi++;
        }
      }
    }
    // the remap may happen anywhere after the last pairing gate on a victim position
    size_t legal_from = i;
    for (size_t k = i; k < x; k++)
      if (is_pairing_kind(batch[k - i].c.kind))
        for (int j = 0; j < k_pairs; j++)
          if (batch[k - i].tpos == lpos_k[j]) legal_from = k + 1;
    size_t pass_end = i, chosen = plan.size();
    for (size_t k = 0; k < plan.size(); k++) {
      pass_end += (size_t)plan[k].n_gates_api;
      if (pass_end >= legal_from) {
        chosen = k;
        break;
      }
    }
    if (chosen == plan.size()) {  // nothing to ride on (empty batch or only value-preserving gates)
      if (e.opt.dryrun) trace_gates(e, batch);
      RC(swap_positions(e, victim, gpos));
      i = x;
      continue;
    }
    SwapStore sw{};
    if (!e.opt.dryrun && !dist_fused_swap_args(e, k_pairs, lpos_k, gpos_k, sw)) {
      RC(run_local(e, batch));
      RC(swap_positions(e, victim, gpos));
      i = x;
      continue;
    }
    const std::vector<PhysGate> prefix(batch.begin(), batch.begin() + (long)(pass_end - i));
    if (e.opt.dryrun) trace_gates(e, prefix);
    {
      // The gates that run before the remap are now fixed (a prefix of the queue): plan just those again,
      // out of order where that is exact (PlannerConfig::reorder_exact), and let the remap ride on the
      // last pass -- legal wherever the in-order plan's carrying pass was, since every gate of the prefix
      // has been applied by then.  Kept when it needs fewer passes.
      std::vector<PassPlan> again = plan_batch(e, prefix, true);
      if (!again.empty() && again.size() < chosen + 1) {
        plan.swap(again);
        chosen = plan.size() - 1;
      }
    }
    RC(launch_passes(e, plan, 0, chosen + 1, &sw, k_pairs, lpos_k, gpos_k));
    for (int j = 0; j < k_pairs; j++) note_swap(e, lpos_k[j], gpos_k[j]);
    // bytes that crossed NVLink per direction: (1 - 2^-k) of the shard, not k halves
    e.exchange_bytes += (16.0 * (1.0 - std::ldexp(1.0, -k_pairs)) - 8.0 * k_pairs) * (double)e.local_size;
    e.fused_swap_bytes += 16.0 * (1.0 - std::ldexp(1.0, -k_pairs)) * (double)e.local_size;
    e.fused_swaps++;
    if (k_pairs > 1) e.multi_remaps++;
    i = pass_end;  // the rest of the batch is planned again under the new layout
  }
  return QCS_CUDA_OK;
}

// Queue-level peephole: drops pairs of identical gates that undo each
// other EXACTLY, before they cost a pass.  Only gates that permute or negate amplitudes qualify
// (X, Y, Z, and their controlled forms: exact X is a move, exact -1 / +-i factors round nothing),
// so the surviving gates see bit-identical inputs; H.H is NOT dropped -- it is the identity only
// up to rounding, and the reference computes it.  Two such gates cancel when every gate between
// them leaves their qubits alone (a permutation / sign flip of qubit a commutes exactly with
// arithmetic that does not involve a).  Corrected semantics only: under `reference` semantics a
// controlled-X is not an involution and the scratch buffer makes the last gate observable.
// The reference's own qc_optimize edits only its history (src/qcs.c:650-698); the host layer keeps
// that behaviour, this pass never changes qc_get_num_gates.
static bool exact_involution(const HostGate &g) {
  const double *m = g.m;
  auto is = [&](int k, double re, double im) { return m[2 * k] == re && m[2 * k + 1] == im; };
  if (is(0, 0, 0) && is(3, 0, 0)) {
    if (is(1, 1, 0) && is(2, 1, 0)) return true;    // X
    if (is(1, 0, -1) && is(2, 0, 1)) return true;   // Y
  }
  if (is(1, 0, 0) && is(2, 0, 0) && is(0, 1, 0) && is(3, -1, 0)) return true;  // Z
  return false;
}

static size_t cancel_exact_pairs(std::vector<HostGate> &q) {
  const size_t before = q.size();
  std::vector<char> dead(q.size(), 0);
  auto touches = [](const HostGate &g, int qubit) { return g.target == qubit || g.control == qubit; };
  for (bool changed = true; changed;) {
    changed = false;
    for (size_t i = 0; i < q.size(); i++) {
      if (dead[i] || !exact_involution(q[i])) continue;
      for (size_t j = i + 1; j < q.size(); j++) {
        if (dead[j]) continue;
        const bool involved = touches(q[j], q[i].target) || (q[i].control >= 0 && touches(q[j], q[i].control));
        if (!involved) continue;
        if (q[j].target == q[i].target && q[j].control == q[i].control &&
            std::memcmp(q[j].m, q[i].m, sizeof(q[i].m)) == 0) {
          dead[i] = dead[j] = 1;
          changed = true;
        }
        break;  // the first gate that involves these qubits decides
      }
    }
  }
  size_t w = 0;
  for (size_t i = 0; i < q.size(); i++)
    if (!dead[i]) q[w++] = q[i];
  q.resize(w);
  return before - w;
}

static int canonicalize(Engine &e);

static int flush_queue(Engine &e);

// Executes everything queued.  A failure part-way (allocation, launch, NCCL, a swap handshake that
// timed out) leaves the amplitudes partially updated and the rest of the queue dropped: the engine is
// then POISONED -- this and every later call that needs the amplitudes returns the error instead of
// silently computing on a state no gate sequence produces.
static int flush(Engine &e) {
  if (e.poisoned)
    return set_error(QCS_CUDA_ERR_CUDA, "engine unusable after an earlier failure: %s", e.poison_reason.c_str());
  int rc = flush_queue(e);
  // every call that reads or edits the amplitudes flushes first: from here on they are in memory
  if (rc == QCS_CUDA_OK) rc = materialize(e);
  if (rc == QCS_CUDA_OK && e.swap_status_pending) {
    rc = check_cuda(cudaStreamSynchronize(e.stream), "cudaStreamSynchronize");
    if (rc == QCS_CUDA_OK) rc = dist_check_fused_swaps(e);
  }
  if (rc != QCS_CUDA_OK) {
    e.poisoned = true;
    e.poison_reason = qcs_cuda_last_error();
  }
  return rc;
}

static int flush_queue(Engine &e) {
  if (e.queue.empty()) return QCS_CUDA_OK;
  e.carried_sum_valid = false;  // gates are about to change the amplitudes
  std::vector<HostGate> q;
  q.swap(e.queue);
  fold_events(e);  // per-pass times of the previous flush are final before its plan is dropped
  e.last_plan.clear();
  e.last_plan_ms.clear();
  e.plan_generation++;
  if (e.opt.sem == SEM_CORRECTED && e.opt.peephole) {
    e.gates_cancelled += (long long)cancel_exact_pairs(q);
    if (q.empty()) return QCS_CUDA_OK;
  }
  if (e.opt.sem == SEM_REFERENCE) {
    // Scratch-observability rule: after any gate the
    // reference's scratch buffer holds the complete pre-gate state.  Execute
    // all but the last gate fused in place, snapshot, then the last gate.
    RC(run_range(e, q, 0, q.size() - 1));
    if (dist().active) RC(canonicalize(e));  // the snapshot is taken in the identity layout
    RC(ensure_scratch(e));
    RC(materialize(e));
    if (!e.opt.dryrun) {
      CK(cudaMemcpyAsync(e.scratch, e.live, e.local_size * sizeof(double2),
                         cudaMemcpyDeviceToDevice, e.stream));
    }
    RC(run_range(e, q, q.size() - 1, q.size()));
  } else {
    RC(run_range(e, q, 0, q.size()));
  }
  return QCS_CUDA_OK;
}

// Trades two LOCAL positions (a bit permutation of the shard, no communication).
static int swap_local_positions(Engine &e, int a, int b) {
  if (a == b) return QCS_CUDA_OK;
  RC(materialize(e));
  if (!e.opt.dryrun) {
    CK(launch_swap_local_bits(e.live, e.nl, a, b, e.stream));
    e.kernel_launches++;
    e.algorithmic_bytes += 16.0 * (double)e.local_size;
  } else {
    Engine::TraceEntry t{};
    t.v[0] = 3.0;
    t.v[1] = (double)a;
    t.v[2] = (double)b;
    e.trace.push_back(t);
  }
  const int qa = e.inv_perm[a], qb = e.inv_perm[b];
  e.perm[qa] = b;
  e.perm[qb] = a;
  e.inv_perm[a] = qb;
  e.inv_perm[b] = qa;
  return QCS_CUDA_OK;
}

// Restores the identity qubit layout (needed by every order-dependent read).
static int canonicalize(Engine &e) {
  // 1. bring every global qubit home (half-shard exchanges)
  for (int g = e.nl; g < e.n; g++) {
    while (e.inv_perm[g] != g) {
      const int where = e.perm[g];  // position currently holding logical qubit g
      if (where < e.nl) {
        RC(swap_positions(e, where, g));
      } else {
        // two global positions hold each other's qubits: route through a local position
        RC(swap_positions(e, e.nl - 1, where));
      }
    }
  }
  // 2. successive swaps through one global position leave the local qubits permuted among
  //    themselves: undo with local bit swaps
  for (int p = 0; p < e.nl; p++) {
    while (e.inv_perm[p] != p) RC(swap_local_positions(e, p, e.inv_perm[p]));
  }
  for (int q = 0; q < e.n; q++)
    if (e.perm[q] != q) return set_error(QCS_CUDA_ERR_CUDA, "internal: layout not canonical");
  return QCS_CUDA_OK;
}

static bool layout_is_identity(const Engine &e) {
  for (int q = 0; q < e.n; q++)
    if (e.perm[q] != q) return false;
  return true;
}

// physical basis index (rank bits on top) of logical basis index `index` under the current layout
static uint64_t physical_index(const Engine &e, uint64_t index) {
  uint64_t phys = 0;
  for (int q = 0; q < e.n; q++)
    if ((index >> q) & 1ull) phys |= 1ull << e.perm[q];
  return phys;
}

static int require_data(const Engine &e) {
  if (e.opt.dryrun) return set_error(QCS_CUDA_ERR_DRYRUN, "plan-only engine holds no amplitudes");
  return QCS_CUDA_OK;
}

// Exact left-to-right sum of |a|^2 over the whole logical state (all ranks),
// optionally masked to indices with bit `pos` clear.  Leaves chunk_exact on
// every rank; *total is the same on every rank.
// pos = SEL_RE / SEL_IM (kernels.h): the sequential sum of the amplitudes' real / imaginary parts
// instead (q_apply_diffusion, reference src/q_gates.c:334-336); same chaining across ranks.
static int exact_total(Engine &e, int pos, double *total) {
  DistContext &d = dist();
  double *res = e.ws.result;
  int mask_local = pos <= SEL_RE ? pos : (pos >= 0 && pos < e.nl) ? pos : SEL_ALL;
  // a masked GLOBAL position: whole shards either count or not
  const bool shard_counts = !(pos >= e.nl && ((e.shard_base >> pos) & 1ull));
  if (!d.active) {
    CK(launch_chunk_sums(e.live, e.local_size, mask_local, e.ws, e.stream));
    CK(launch_chunk_deltas(e.live, e.local_size, mask_local, res + RES_ZERO, e.ws, e.stream));
    CK(launch_chunk_resolve(e.live, e.local_size, mask_local, res + RES_ZERO, e.ws, e.stream));
    e.kernel_launches += 5;
    CK(cudaMemcpyAsync(total, res + RES_EXACT_TOTAL, sizeof(double), cudaMemcpyDeviceToHost,
                       e.stream));
    CK(cudaStreamSynchronize(e.stream));
    e.algorithmic_bytes += 2 * 16.0 * (double)e.local_size * (mask_local >= 0 ? 0.5 : 1.0);
    return QCS_CUDA_OK;
  }
  // Multi-rank: approximate offsets first (parallel), then the exact walk chained rank by rank.
  std::vector<double> approx_tot(d.world, 0.0);
  double mine = 0.0;
  if (shard_counts) {
    CK(launch_chunk_sums(e.live, e.local_size, mask_local, e.ws, e.stream));
    e.kernel_launches += 2;
    if (e.local_size >= (uint64_t)SEQ_CHUNK) {
      CK(cudaMemcpyAsync(&mine, res + RES_APPROX_TOTAL, sizeof(double), cudaMemcpyDeviceToHost,
                         e.stream));
      CK(cudaStreamSynchronize(e.stream));
    }
  }
  RC(dist_allgather_host(&mine, approx_tot.data(), sizeof(double)));
  double approx_start = 0.0;
  for (int r = 0; r < d.rank; r++) approx_start += approx_tot[r];
  CK(cudaMemcpyAsync(res + RES_START, &approx_start, sizeof(double), cudaMemcpyHostToDevice,
                     e.stream));
  if (shard_counts) {
    CK(launch_chunk_deltas(e.live, e.local_size, mask_local, res + RES_START, e.ws, e.stream));
    e.kernel_launches += 2;
  }
  // exact chain: rank r starts from the exact running sum after rank r-1
  std::vector<double> exact_after(d.world, 0.0);
  double running = 0.0;
  for (int r = 0; r < d.world; r++) {
    if (r == d.rank) {
      if (shard_counts) {
        CK(cudaMemcpyAsync(res + RES_START + 1, &running, sizeof(double), cudaMemcpyHostToDevice,
                           e.stream));
        CK(launch_chunk_resolve(e.live, e.local_size, mask_local, res + RES_START + 1, e.ws,
                                e.stream));
        e.kernel_launches++;
        CK(cudaMemcpyAsync(&mine, res + RES_EXACT_TOTAL, sizeof(double), cudaMemcpyDeviceToHost,
                           e.stream));
        CK(cudaStreamSynchronize(e.stream));
      } else {
        mine = running;
      }
    }
    double tmp = (r == d.rank) ? mine : 0.0;
    std::vector<double> all(d.world, 0.0);
    RC(dist_allgather_host(&tmp, all.data(), sizeof(double)));
    running = all[r];
    exact_after[r] = running;
  }
  *total = running;
  return QCS_CUDA_OK;
}

static int normalize_now(Engine &e) {
  e.carried_sum_valid = false;
  double total = 0.0;
  RC(exact_total(e, -1, &total));
  // q_state_normalize (reference src/q_utils.c:111-117)
  if (total > 1e-12 && total != 1.0) {
    const double inv = 1.0 / std::sqrt(total);
    CK(launch_scale(e.live, e.local_size, inv, e.stream));
    e.kernel_launches++;
    e.algorithmic_bytes += 32.0 * (double)e.local_size;
  }
  return QCS_CUDA_OK;
}

}  // namespace qcs

// ====================================================================== ABI
using namespace qcs;

extern "C" {

const char *qcs_cuda_last_error(void) { return g_error; }

int qcs_cuda_set_default(const char *key, const char *value) {
  if (!key) return set_error(QCS_CUDA_ERR_INVALID, "null key");
  std::lock_guard<std::mutex> lock(globals_mutex());
  if (!value) defaults().erase(key);
  else defaults()[key] = value;
  return QCS_CUDA_OK;
}

long qcs_cuda_get_default(const char *key, char *buf, long cap) {
  if (!key) return -1;
  std::lock_guard<std::mutex> lock(globals_mutex());
  auto it = defaults().find(key);
  if (it == defaults().end()) return -1;
  if (buf && cap > 0) {
    const long n = (long)it->second.size() < cap - 1 ? (long)it->second.size() : cap - 1;
    std::memcpy(buf, it->second.data(), (size_t)n);
    buf[n] = 0;
  }
  return (long)it->second.size();
}

long qcs_cuda_trim_pool(void) {
  std::lock_guard<std::mutex> lock(globals_mutex());
  long freed = 0;
  for (auto &p : g_pool) freed += (long)p.bytes;
  pool_trim_locked();
  return freed;
}

int qcs_cuda_state_create(qcs_cuda_engine **out, int n_qubits) {
  if (!out) return set_error(QCS_CUDA_ERR_INVALID, "null out pointer");
  *out = nullptr;
  if (n_qubits <= 0)  // reference: "Number of qubits must be positive" (src/q_state.c:38-41)
    return set_error(QCS_CUDA_ERR_INVALID, "Number of qubits must be positive.");
  DistContext &d = dist();
  const int rank_bits = d.active ? d.rank_bits : 0;
  if (n_qubits > 40 || n_qubits - rank_bits < 1)
    return set_error(QCS_CUDA_ERR_INVALID, "unsupported qubit count %d for %d rank(s)", n_qubits,
                     d.active ? d.world : 1);
  qcs_cuda_engine *e = new qcs_cuda_engine();
  int rc = load_options(e->opt);
  if (rc != QCS_CUDA_OK) {
    delete e;
    return rc;
  }
  e->n = n_qubits;
  e->nl = n_qubits - rank_bits;
  e->local_size = 1ull << e->nl;
  e->shard_base = d.active ? ((uint64_t)d.rank << e->nl) : 0ull;
  for (int q = 0; q < 64; q++) e->perm[q] = e->inv_perm[q] = q;
  if (e->opt.dryrun) {
    *out = e;
    return QCS_CUDA_OK;
  }
  if (d.active && !d.comm) {
    delete e;
    return set_error(QCS_CUDA_ERR_INVALID, "plan-only sharding needs dry-run engines");
  }
  auto fail = [&](int code) {
    qcs_cuda_state_destroy(e);
    return code;
  };
  std::string dev = option_value("device");
  if (d.active) {
    rc = check_cuda(cudaSetDevice(d.device), "cudaSetDevice");
  } else if (!dev.empty()) {
    rc = check_cuda(cudaSetDevice(std::atoi(dev.c_str())), "cudaSetDevice");
  } else {
    int cur = 0;
    rc = check_cuda(cudaGetDevice(&cur), "cudaGetDevice (no CUDA device: QCS_GPU_CUDA has no CPU fallback)");
  }
  if (rc) return fail(rc);
  if ((rc = check_cuda(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking), "cudaStreamCreate")))
    return fail(rc);
  if ((rc = check_cuda(pool_alloc((void **)&e->live, e->local_size * sizeof(double2)), "cudaMalloc(state)")))
    return fail(rc);
  ReduceWorkspace &ws = e->ws;
  const size_t n_chunks = e->local_size >= SEQ_CHUNK ? e->local_size / SEQ_CHUNK : 1;
  ws.n_chunks_cap = n_chunks;
  {
    // one slab for every small work array (a dozen cudaMallocs otherwise)
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t sz[] = {up(4 * REDUCE_MAX_BLOCKS * sizeof(double)), up(REDUCE_MAX_BLOCKS * sizeof(long long)),
                         up(RES_COUNT * sizeof(double)), up(4 * sizeof(long long)),
                         up(n_chunks * sizeof(double)), up((n_chunks + 1) * sizeof(double)),
                         up(n_chunks * sizeof(double)), up(n_chunks), up((n_chunks + 1) * sizeof(double)),
                         up((n_chunks / QCS_RESOLVE_GROUP + 2) * sizeof(double)),
                         up((n_chunks / QCS_RESOLVE_GROUP + 2) * sizeof(double)), up(n_chunks / QCS_RESOLVE_GROUP + 2), up(n_chunks * sizeof(double)), up(n_chunks * sizeof(double)),
                         up(n_chunks * sizeof(double))};
    size_t total = 0;
    for (size_t b : sz) total += b;
    e->ws_slab_bytes = total;
    if ((rc = check_cuda(pool_alloc(&e->ws_slab, total), "cudaMalloc(workspace)"))) return fail(rc);
    char *p = (char *)e->ws_slab;
    ws.partials = (double *)p; p += sz[0];
    ws.ipartials = (long long *)p; p += sz[1];
    ws.result = (double *)p; p += sz[2];
    ws.iresult = (long long *)p; p += sz[3];
    ws.chunk_sum = (double *)p; p += sz[4];
    ws.chunk_approx = (double *)p; p += sz[5];
    ws.chunk_delta = (double *)p; p += sz[6];
    ws.chunk_flag = (unsigned char *)p; p += sz[7];
    ws.chunk_exact = (double *)p; p += sz[8];
    ws.group_sum = (double *)p; p += sz[9];
    ws.group_exact = (double *)p; p += sz[10];
    ws.group_flag = (unsigned char *)p; p += sz[11];
    ws.chunk_sum2 = (double *)p; p += sz[12];
    ws.chunk_abs = (double *)p; p += sz[13];
    ws.chunk_abs2 = (double *)p;
  }
  if ((rc = check_cuda(cudaMemsetAsync(ws.result, 0, RES_COUNT * sizeof(double), e->stream), "cudaMemset")))
    return fail(rc);
  // zero everything, amplitude 0 := 1 (reference src/q_state.c:93-99) -- lazily, see materialize()
  e->zero_ket_pending = true;
  if (!e->opt.lazy_init && (rc = materialize(*e))) return fail(rc);
  if ((rc = check_cuda(cudaStreamSynchronize(e->stream), "init sync"))) return fail(rc);
  if (d.active && e->opt.exchange == 1 && e->opt.fuse_swaps && e->opt.remap_buffer != 1 &&
      e->opt.sem == SEM_CORRECTED && e->nl >= QCS_MIN_TILE_BITS) {
    // Second shard buffer for the passes that carry a remap (kernels.h SwapStore::out_of_place).  auto: a
    // buffer the pool already holds, else a fresh one only while an eighth of the device stays free;
    // dist_open_peers drops it again unless every rank got one.
    const size_t bytes = e->local_size * sizeof(double2);
    void *buf = nullptr;
    if (!pool_take_cached(&buf, bytes)) {
      size_t free_b = 0, total_b = 0;
      bool fits = e->opt.remap_buffer == 2;
      if (!fits && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) fits = free_b >= bytes + total_b / 8;
      if (fits && pool_alloc(&buf, bytes) != cudaSuccess) {
        cudaGetLastError();
        buf = nullptr;
      }
    }
    e->alt = (double2 *)buf;
  }
  if (d.active && e->opt.exchange == 1 && (rc = dist_open_peers(*e))) return fail(rc);
  if (e->opt.remap_buffer == 2 && d.active && !e->alt)
    return fail(set_error(QCS_CUDA_ERR_CUDA, "remap_buffer=double: no second shard buffer on every rank (memory, or "
                                             "no peer access / exchange=nccl / semantics=reference)"));
  *out = e;
  return QCS_CUDA_OK;
}

void qcs_cuda_state_destroy(qcs_cuda_engine *e) {
  if (!e) return;
  if (e->opt.dryrun) {
    delete e;
    return;
  }
  if (e->stream) cudaStreamSynchronize(e->stream);
  // (no barrier: every pass or swap through which a peer wrote into this buffer was followed by an
  // all-reduce on the streams of both ranks, and this rank's stream has just been drained)
  if (dist().active) dist_close_peers(*e);
  for (auto &pp : e->pending_pass_events) { cudaEventDestroy(pp.begin); cudaEventDestroy(pp.end); }
  for (auto &pr : e->pending_xchg_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (auto ev : e->event_pool) cudaEventDestroy(ev);
  for (auto ev : e->markers) if (ev) cudaEventDestroy(ev);
  pool_free(e->live, e->local_size * sizeof(double2));
  pool_free(e->scratch, e->local_size * sizeof(double2));
  pool_free(e->alt, e->local_size * sizeof(double2));
  pool_free(e->staging, e->local_size * sizeof(double2));
  pool_free(e->ws_slab, e->ws_slab_bytes);
  cudaFree(e->u_dev);
  cudaFree(e->idx_dev);
  cudaFree(e->table_buf);
  pool_free(e->argmax_tile_p, (e->local_size >> QCS_MIN_TILE_BITS) * sizeof(double));
  pool_free(e->argmax_tile_idx, (e->local_size >> QCS_MIN_TILE_BITS) * sizeof(long long));
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int qcs_cuda_num_qubits(const qcs_cuda_engine *e) { return e ? e->n : 0; }

int qcs_cuda_get_layout(const qcs_cuda_engine *e, int *perm) {
  if (!e || !perm) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  for (int q = 0; q < e->n; q++) perm[q] = e->perm[q];
  return QCS_CUDA_OK;
}

int qcs_cuda_apply_1q(qcs_cuda_engine *e, const double m[8], int target) {
  if (!e || !m) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  if (target < 0 || target >= e->n)  // reference src/q_gates.c:32-36
    return set_error(QCS_CUDA_ERR_INVALID, "Invalid arguments for 1-qubit gate application.");
  HostGate g;
  std::memcpy(g.m, m, sizeof(g.m));
  g.target = target;
  g.control = -1;
  e->queue.push_back(g);
  e->gates_submitted++;
  if (e->queue.size() >= 4096) return flush(*e);
  return QCS_CUDA_OK;
}

int qcs_cuda_apply_c1q(qcs_cuda_engine *e, const double m[8], int control, int target) {
  if (!e || !m) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  if (control < 0 || target < 0 || control >= e->n || target >= e->n || control == target)
    return set_error(QCS_CUDA_ERR_INVALID, "Invalid arguments for 2-qubit gate application.");  // :165-170
  HostGate g;
  std::memcpy(g.m, m, sizeof(g.m));
  g.target = target;
  g.control = control;
  e->queue.push_back(g);
  e->gates_submitted++;
  if (e->queue.size() >= 4096) return flush(*e);
  return QCS_CUDA_OK;
}

int qcs_cuda_flush(qcs_cuda_engine *e) {
  if (!e) return set_error(QCS_CUDA_ERR_INVALID, "null engine");
  RC(flush(*e));
  if (!e->opt.dryrun) CK(cudaStreamSynchronize(e->stream));
  return QCS_CUDA_OK;
}

int qcs_cuda_phase_flip(qcs_cuda_engine *e, long index) {
  if (!e) return set_error(QCS_CUDA_ERR_INVALID, "null engine");
  if (index < 0 || (uint64_t)index >= (1ull << e->n))  // reference src/q_gates.c:306-309
    return set_error(QCS_CUDA_ERR_INVALID, "Invalid state or index for phase flip.");
  RC(flush(*e));
  // reference semantics: the buffers trade roles below, so both must be in the identity layout
  // (flush snapshots the scratch buffer in it)
  if (e->opt.sem == SEM_REFERENCE) RC(canonicalize(*e));
  if (e->opt.dryrun) return QCS_CUDA_OK;
  // one amplitude: found through the layout, wherever position swaps have left it
  const uint64_t phys = physical_index(*e, (uint64_t)index);
  const bool mine = (phys >> e->nl) == (e->shard_base >> e->nl);
  const uint64_t local = phys & (e->local_size - 1);
  if (e->opt.sem == SEM_REFERENCE) {
    // swap buffers, then live[idx] = -scratch[idx] (reference src/q_gates.c:311-316)
    RC(ensure_scratch(*e));
    std::swap(e->live, e->scratch);
    if (mine) {
      CK(launch_negate_one(e->live, e->scratch, local, nullptr, e->stream));
      e->kernel_launches++;
    }
  } else if (mine) {
    CK(launch_negate_one(e->live, e->live, local,
                         e->carried_sum_valid ? e->ws.result + RES_LOCAL_SUM_RE : nullptr, e->stream));
    e->kernel_launches++;
  }
  return QCS_CUDA_OK;
}

int qcs_cuda_diffusion(qcs_cuda_engine *e) {
  if (!e) return set_error(QCS_CUDA_ERR_INVALID, "null engine");
  RC(flush(*e));
  if (e->opt.dryrun) return QCS_CUDA_OK;
  double *res = e->ws.result;
  if (e->opt.fast_math) {
    // math=fast: tree sums (more accurate than the reference's sequential sum, hence NOT within 1e-12
    // of it on large registers -- DESIGN.md); the write pass of one diffusion carries the sum for the
    // next when nothing but phase flips touched the state since (Grover's loop)
    if (!e->carried_sum_valid) {
      CK(launch_complex_sum(e->live, e->local_size, e->ws, e->stream));
      e->kernel_launches += 2;
      e->algorithmic_bytes += 16.0 * (double)e->local_size;
    }
    CK(cudaMemcpyAsync(res + RES_SUM_RE, res + RES_LOCAL_SUM_RE, 2 * sizeof(double),
                       cudaMemcpyDeviceToDevice, e->stream));
    if (dist().active) RC(dist_allreduce_sum(*e, res + RES_SUM_RE, 2));
  } else if (!dist().active) {
    // The reference's left-to-right sums of the real and of the imaginary parts, bit for bit
    // (kernels_reduce.cu, "Signed sums"); everything stays on the stream, no host round trip.
    CK(launch_chunk_sums_complex(e->live, e->local_size, e->ws, e->stream));
    e->kernel_launches++;
    for (int comp = 0; comp < 2; comp++) {
      const int sel = comp == 0 ? SEL_RE : SEL_IM;
      // (a Grover state is real: the imaginary parts' sum is skipped on the device when all are zero)
      const long long *skip = comp == 1 ? e->ws.iresult + 2 : nullptr;
      CK(launch_chunk_deltas(e->live, e->local_size, sel, res + RES_ZERO, e->ws, e->stream, skip));
      CK(launch_chunk_resolve(e->live, e->local_size, sel, res + RES_ZERO, e->ws, e->stream, true, skip));
      CK(cudaMemcpyAsync(res + RES_SUM_RE + comp, res + RES_EXACT_TOTAL, sizeof(double),
                         cudaMemcpyDeviceToDevice, e->stream));
      e->kernel_launches += 4;  // scan, deltas, group sums, walk (no per-chunk prefixes: total only)
    }
    e->algorithmic_bytes += 3 * 16.0 * (double)e->local_size;
  } else {
    // sharded: rank r continues the sum where rank r-1 stopped, in basis-index order
    RC(canonicalize(*e));
    double sum[2] = {0.0, 0.0};
    RC(exact_total(*e, SEL_RE, &sum[0]));
    RC(exact_total(*e, SEL_IM, &sum[1]));
    CK(cudaMemcpyAsync(res + RES_SUM_RE, sum, 2 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));  // `sum` is a stack buffer
    e->algorithmic_bytes += 4 * 16.0 * (double)e->local_size;
  }
  CK(launch_diffusion_mean(e->ws, e->opt.sem == SEM_CORRECTED, (double)(1ull << e->n), e->stream));
  if (e->opt.sem == SEM_REFERENCE) {
    // new values go to the other buffer, then the buffers trade roles (:348-355)
    RC(ensure_scratch(*e));
    CK(launch_diffusion_write(e->live, e->scratch, e->local_size, e->ws, e->stream));
    std::swap(e->live, e->scratch);
    e->kernel_launches += 2;
  } else if (e->opt.fast_math) {
    CK(launch_diffusion_write_sum(e->live, e->live, e->local_size, e->ws, e->stream));
    e->carried_sum_valid = true;
    e->kernel_launches += 3;
  } else {
    CK(launch_diffusion_write(e->live, e->live, e->local_size, e->ws, e->stream));
    e->kernel_launches += 2;
  }
  e->algorithmic_bytes += 32.0 * (double)e->local_size;
  return QCS_CUDA_OK;
}

int qcs_cuda_normalize(qcs_cuda_engine *e) {
  if (!e) return set_error(QCS_CUDA_ERR_INVALID, "null engine");
  RC(flush(*e));
  RC(require_data(*e));
  RC(canonicalize(*e));
  return normalize_now(*e);
}

int qcs_cuda_prob0(qcs_cuda_engine *e, int qubit, double *p0) {
  if (!e || !p0) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  if (qubit < 0 || qubit >= e->n) return set_error(QCS_CUDA_ERR_INVALID, "qubit out of range");
  RC(flush(*e));
  RC(require_data(*e));
  RC(canonicalize(*e));
  return exact_total(*e, qubit, p0);
}

int qcs_cuda_collapse(qcs_cuda_engine *e, int qubit, int outcome) {
  if (!e) return set_error(QCS_CUDA_ERR_INVALID, "null engine");
  if (qubit < 0 || qubit >= e->n || (outcome != 0 && outcome != 1))
    return set_error(QCS_CUDA_ERR_INVALID, "bad collapse arguments");
  RC(flush(*e));
  RC(require_data(*e));
  RC(canonicalize(*e));
  if (qubit < e->nl) {
    CK(launch_zero_half(e->live, e->local_size, qubit, outcome, e->stream));
    e->algorithmic_bytes += 8.0 * (double)e->local_size;
  } else if ((int)((e->shard_base >> qubit) & 1ull) != outcome) {
    CK(launch_init_state(e->live, e->local_size, false, e->stream));
    e->algorithmic_bytes += 16.0 * (double)e->local_size;
  }
  e->kernel_launches++;
  return normalize_now(*e);
}

int qcs_cuda_read_amplitudes(qcs_cuda_engine *e, int which, long first, long count, double *out) {
  if (!e || !out) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  RC(flush(*e));
  RC(require_data(*e));
  RC(canonicalize(*e));
  const uint64_t base = e->shard_base;
  if (first < 0 || count < 0 || (uint64_t)first < base ||
      (uint64_t)first + (uint64_t)count > base + e->local_size)
    return set_error(QCS_CUDA_ERR_INVALID, "amplitude range [%ld, %ld) is not inside this rank's shard", first, first + count);
  const double2 *src = which ? e->scratch : e->live;
  if (which && !src) {  // never touched: still the zeros of q_state_init
    std::memset(out, 0, (size_t)count * 16);
    return QCS_CUDA_OK;
  }
  CK(cudaMemcpyAsync(out, src + ((uint64_t)first - base), (size_t)count * sizeof(double2),
                     cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return QCS_CUDA_OK;
}

int qcs_cuda_write_amplitudes(qcs_cuda_engine *e, int which, long first, long count, const double *in) {
  if (!e || !in) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  RC(flush(*e));
  RC(require_data(*e));
  RC(canonicalize(*e));
  const uint64_t base = e->shard_base;
  if (first < 0 || count < 0 || (uint64_t)first < base ||
      (uint64_t)first + (uint64_t)count > base + e->local_size)
    return set_error(QCS_CUDA_ERR_INVALID, "amplitude range is not inside this rank's shard");
  if (which) RC(ensure_scratch(*e));
  e->carried_sum_valid = false;
  double2 *dst = which ? e->scratch : e->live;
  CK(cudaMemcpyAsync(dst + ((uint64_t)first - base), in, (size_t)count * sizeof(double2),
                     cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return QCS_CUDA_OK;
}

int qcs_cuda_get_amplitude(qcs_cuda_engine *e, long index, double out[2]) {
  if (!e || !out) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  if (index < 0 || (uint64_t)index >= (1ull << e->n))
    return set_error(QCS_CUDA_ERR_INVALID, "state index out of range");
  RC(flush(*e));
  RC(require_data(*e));
  // one amplitude: found through the layout, wherever position swaps have left it
  const uint64_t phys = physical_index(*e, (uint64_t)index);
  double2 v = make_double2(0.0, 0.0);
  const bool mine = (phys >> e->nl) == (e->shard_base >> e->nl);
  if (mine) {
    CK(cudaMemcpyAsync(&v, e->live + (phys & (e->local_size - 1)), sizeof(v),
                       cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  if (dist().active) {
    std::vector<double2> all(dist().world);
    RC(dist_allgather_host(&v, all.data(), sizeof(v)));
    v = all[phys >> e->nl];
  }
  out[0] = v.x;
  out[1] = v.y;
  return QCS_CUDA_OK;
}

int qcs_cuda_probability(qcs_cuda_engine *e, long index, double *p) {
  if (!e || !p) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  if (index < 0 || (uint64_t)index >= (1ull << e->n)) {  // reference src/qcs.c:392-393
    *p = 0.0;
    return QCS_CUDA_OK;
  }
  double a[2];
  RC(qcs_cuda_get_amplitude(e, index, a));
  *p = (a[0] * a[0]) + (a[1] * a[1]);  // c_norm_sq; host FMA contraction is off (-ffp-contract=off)
  return QCS_CUDA_OK;
}

int qcs_cuda_argmax(qcs_cuda_engine *e, long *index) {
  if (!e || !index) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  // Gates still queued on a single-GPU engine: the flush's last pass can do the first level of the
  // reduction while the amplitudes are in registers (one candidate per tile), saving a 16-byte-per-
  // amplitude sweep.  In-order plans only (one launch group, its last pass is the last word on every
  // amplitude); same values, same tie rule as the stand-alone kernels.
  e->argmax_request = e->argmax_done = false;
  if (!e->queue.empty() && !dist().active && !e->opt.dryrun && e->opt.fuse_argmax && e->opt.fusion &&
      e->opt.sem == SEM_CORRECTED && (e->opt.tile_kernel == 0 || e->opt.tile_kernel == 3) &&
      e->nl >= min_tile_bits(*e) && !e->poisoned) {
    const size_t n = (size_t)(e->local_size >> QCS_MIN_TILE_BITS);
    if (!e->argmax_tile_p) {
      if (pool_alloc((void **)&e->argmax_tile_p, n * sizeof(double)) != cudaSuccess) e->argmax_tile_p = nullptr;
      if (e->argmax_tile_p && pool_alloc((void **)&e->argmax_tile_idx, n * sizeof(long long)) != cudaSuccess) {
        pool_free(e->argmax_tile_p, n * sizeof(double));
        e->argmax_tile_p = nullptr;
      }
      cudaGetLastError();
    }
    e->argmax_request = e->argmax_tile_p != nullptr;
  }
  const int flush_rc = flush(*e);
  const bool fused = e->argmax_request && e->argmax_done;
  e->argmax_request = e->argmax_done = false;
  RC(flush_rc);
  RC(require_data(*e));
  const bool permuted = !layout_is_identity(*e);
  if (fused) {
    CK(launch_argmax_pairs(e->argmax_tile_p, e->argmax_tile_idx, e->argmax_tiles, e->ws, e->stream));
  } else if (permuted) {
    // ties are decided by logical index inside the kernel: no need to restore the layout
    LogicalIndexLut lut;
    std::memset(&lut, 0, sizeof(lut));
    for (int p = e->nl; p < e->n; p++)
      if ((e->shard_base >> p) & 1ull) lut.base |= 1ull << e->inv_perm[p];
    for (int byte = 0; byte < 4; byte++)
      for (int v = 0; v < 256; v++) {
        unsigned long long l = 0;
        for (int b = 0; b < 8; b++) {
          const int p = 8 * byte + b;
          if (p < e->nl && ((v >> b) & 1)) l |= 1ull << e->inv_perm[p];
        }
        lut.t[byte][v] = l;
      }
    CK(launch_argmax_permuted(e->live, e->local_size, lut, e->ws, e->stream));
  } else {
    CK(launch_argmax(e->live, e->local_size, e->ws, e->stream));
  }
  e->kernel_launches += 2;
  if (!fused) e->algorithmic_bytes += 16.0 * (double)e->local_size;
  struct { double p; long long idx; } mine{0.0, 0};
  CK(cudaMemcpyAsync(&mine.p, e->ws.result, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(&mine.idx, e->ws.iresult, sizeof(long long), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (!permuted && !fused) mine.idx += (long long)e->shard_base;
  if (dist().active) {
    std::vector<decltype(mine)> all(dist().world);
    RC(dist_allgather_host(&mine, all.data(), sizeof(mine)));
    double best_p = 0.0;
    long long best_i = 0;
    for (auto &c : all)  // the first maximum in LOGICAL index order wins (strict >, qcs.c:472)
      if (c.p > best_p || (c.p == best_p && c.p > 0.0 && c.idx < best_i)) { best_p = c.p; best_i = c.idx; }
    mine.idx = best_i;
  }
  *index = (long)mine.idx;
  return QCS_CUDA_OK;
}

int qcs_cuda_sample(qcs_cuda_engine *e, const double *u, int shots, long *indices) {
  if (!e || !u || !indices) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  if (shots <= 0) return QCS_CUDA_OK;
  RC(flush(*e));
  RC(require_data(*e));
  RC(canonicalize(*e));
  double total = 0.0;
  RC(exact_total(*e, -1, &total));
  if (shots > e->shots_cap) {
    cudaFree(e->u_dev);
    cudaFree(e->idx_dev);
    e->u_dev = nullptr;
    e->idx_dev = nullptr;
    e->shots_cap = 0;
    CK(cudaMalloc(&e->u_dev, (size_t)shots * sizeof(double)));
    CK(cudaMalloc(&e->idx_dev, (size_t)shots * sizeof(long long)));
    e->shots_cap = shots;
  }
  CK(cudaMemcpyAsync(e->u_dev, u, (size_t)shots * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  CK(launch_sample(e->live, e->local_size, e->ws, e->u_dev, shots, e->idx_dev,
                   (long long)e->shard_base, e->stream));
  e->kernel_launches++;
  static_assert(sizeof(long) == sizeof(long long), "LP64 expected");
  // multi-rank: exactly one rank found each surviving shot, everyone else wrote -1
  RC(dist_allreduce_max_i64(*e, e->idx_dev, (size_t)shots));
  CK(cudaMemcpyAsync(indices, e->idx_dev, (size_t)shots * sizeof(long long), cudaMemcpyDeviceToHost,
                     e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return QCS_CUDA_OK;
}

int qcs_cuda_set_timing(qcs_cuda_engine *e, int enabled) {
  if (!e) return set_error(QCS_CUDA_ERR_INVALID, "null engine");
  e->timing = enabled != 0;
  return QCS_CUDA_OK;
}

int qcs_cuda_get_stats(qcs_cuda_engine *e, qcs_cuda_stats *out) {
  if (!e || !out) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  fold_events(*e);
  out->gates_submitted = e->gates_submitted;
  out->gates_executed = e->gates_executed;
  out->passes = e->passes;
  out->kernel_launches = e->kernel_launches;
  out->segments = e->segments;
  out->remaps = e->remaps;
  out->algorithmic_bytes = e->algorithmic_bytes;
  out->pass_bytes = e->pass_bytes;
  out->pass_ms = e->pass_ms;
  out->exchange_bytes = e->exchange_bytes;
  out->exchange_ms = e->exchange_ms;
  out->fused_remaps = e->fused_swaps;
  out->fused_remap_pass_ms = e->fused_swap_pass_ms;
  out->pass_flops_per_amp = e->pass_flops_per_amp;
  out->gates_cancelled = e->gates_cancelled;
  out->multi_remaps = e->multi_remaps;
  out->out_of_place_remaps = e->out_of_place_remaps;
  out->fused_remap_bytes = e->fused_swap_bytes;
  return QCS_CUDA_OK;
}

int qcs_cuda_reset_stats(qcs_cuda_engine *e) {
  if (!e) return set_error(QCS_CUDA_ERR_INVALID, "null engine");
  fold_events(*e);
  e->gates_submitted = e->gates_executed = e->passes = e->kernel_launches = e->segments = e->remaps = 0;
  e->algorithmic_bytes = e->pass_bytes = e->pass_ms = e->exchange_bytes = e->exchange_ms = 0;
  e->fused_swaps = 0;
  e->fused_swap_bytes = 0;
  e->multi_remaps = 0;
  e->out_of_place_remaps = 0;
  e->gates_cancelled = 0;
  e->fused_swap_pass_ms = 0;
  e->pass_flops_per_amp = 0;
  return QCS_CUDA_OK;
}

int qcs_cuda_marker_record(qcs_cuda_engine *e, int slot) {
  if (!e || slot < 0 || slot >= 8) return set_error(QCS_CUDA_ERR_INVALID, "bad marker slot");
  RC(require_data(*e));
  if (!e->markers[slot]) CK(cudaEventCreate(&e->markers[slot]));
  CK(cudaEventRecord(e->markers[slot], e->stream));
  return QCS_CUDA_OK;
}

int qcs_cuda_marker_elapsed_ms(qcs_cuda_engine *e, int from_slot, int to_slot, double *ms) {
  if (!e || !ms || from_slot < 0 || from_slot >= 8 || to_slot < 0 || to_slot >= 8 ||
      !e->markers[from_slot] || !e->markers[to_slot])
    return set_error(QCS_CUDA_ERR_INVALID, "marker not recorded");
  CK(cudaEventSynchronize(e->markers[to_slot]));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, e->markers[from_slot], e->markers[to_slot]));
  *ms = (double)f;
  return QCS_CUDA_OK;
}

int qcs_cuda_probe_fp64(double *ops_per_second) {
  if (!ops_per_second) return set_error(QCS_CUDA_ERR_INVALID, "null argument");
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, iters = 1 << 14;
  double *out = nullptr;
  CK(cudaMalloc(&out, (size_t)blocks * 256 * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  cudaError_t err = cudaSuccess;
  for (int rep = 0; rep < 4 && err == cudaSuccess; rep++) {  // first repetition warms up
    cudaEventRecord(e0, 0);
    err = launch_fp64_probe(out, blocks, iters, 0);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  CK(err);
  *ops_per_second = 16.0 * iters * 256.0 * blocks / (best * 1e-3);
  return QCS_CUDA_OK;
}

long qcs_cuda_pass_descriptor_bytes(void) { return (long)sizeof(PassParams); }

long qcs_cuda_trace_read(qcs_cuda_engine *e, long index, double out[12]) {
  if (!e) return 0;
  if (out && index >= 0 && (size_t)index < e->trace.size())
    std::memcpy(out, e->trace[(size_t)index].v, sizeof(double) * 12);
  return (long)e->trace.size();
}

long qcs_cuda_describe_last_plan(qcs_cuda_engine *e, char *buf, long cap) {
  if (!e) return 0;
  std::string s = describe_plan(e->last_plan);
  if (buf && cap > 0) {
    long n = (long)s.size() < cap - 1 ? (long)s.size() : cap - 1;
    std::memcpy(buf, s.data(), (size_t)n);
    buf[n] = 0;
  }
  return (long)s.size() + 1;
}

long qcs_cuda_last_plan_swap(qcs_cuda_engine *e, long pass_index, int *lpos, int *gpos) {
  if (!e || pass_index < 0 || pass_index >= (long)e->last_plan.size()) return 0;
  const PassPlan &p = e->last_plan[(size_t)pass_index];
  for (int i = 0; i < p.n_swaps && i < 3; i++) {
    if (lpos) lpos[i] = p.swap_lpos[i];
    if (gpos) gpos[i] = p.swap_gpos[i];
  }
  return p.n_swaps;
}

long qcs_cuda_last_plan_pass_info(qcs_cuda_engine *e, long pass_index, double *ms, double *flops_per_amp,
                                  int *tile_bits, int *segments) {
  if (!e) return 0;
  fold_events(*e);
  if (pass_index < 0 || pass_index >= (long)e->last_plan.size()) return (long)e->last_plan.size();
  const PassPlan &p = e->last_plan[(size_t)pass_index];
  if (ms) *ms = (size_t)pass_index < e->last_plan_ms.size() ? e->last_plan_ms[(size_t)pass_index] : -1.0;
  if (flops_per_amp) *flops_per_amp = p.flops_per_amp;
  if (tile_bits) *tile_bits = p.params.tile_bits;
  if (segments) *segments = p.params.n_segments;
  return (long)e->last_plan.size();
}

long qcs_cuda_last_plan_tables(qcs_cuda_engine *e, long pass_index, double *buf, long cap) {
  if (!e || pass_index < 0 || pass_index >= (long)e->last_plan.size()) return 0;
  const std::vector<double> &t = e->last_plan[(size_t)pass_index].thread_tables;
  if (buf && cap >= (long)t.size() && !t.empty()) std::memcpy(buf, t.data(), t.size() * sizeof(double));
  return (long)t.size();
}

long qcs_cuda_last_plan_raw(qcs_cuda_engine *e, long pass_index, void *buf, long cap) {
  if (!e) return 0;
  if (buf && pass_index >= 0 && pass_index < (long)e->last_plan.size() && cap >= (long)sizeof(PassParams))
    std::memcpy(buf, &e->last_plan[(size_t)pass_index].params, sizeof(PassParams));
  return (long)e->last_plan.size();
}

}  // extern "C"
