// planner.cpp -- see planner.h.  Pure host code.
#include "planner.h"

#include "fused_ops.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>

namespace qcs {

static inline bool is_zero(double x) { return x == 0.0; }  // true for -0.0 too
static inline bool is_one_c(const double *c) { return c[0] == 1.0 && is_zero(c[1]); }
static inline bool is_zero_c(const double *c) { return is_zero(c[0]) && is_zero(c[1]); }

Classified classify_gate(const double m[8], bool controlled, Semantics sem) {
  Classified r;
  std::memcpy(r.m, m, sizeof(r.m));
  r.flags = 0;
  const double *g0 = m, *g1 = m + 2, *g2 = m + 4, *g3 = m + 6;
  const double ctrl_frac = controlled ? 0.5 : 1.0;

  if (controlled && sem == SEM_REFERENCE) {
    // As written in the reference (src/q_gates.c:276-292, defect D1) only the
    // target=0 member of a control=1 pair changes: n0 = g0*v0 + g1*v1.
    r.flags |= GF_ROW0_ONLY;
    if (is_zero_c(g1)) {
      // n0 = g0*v0 (+ exact zero): elementwise on the target=0 element.
      r.m[2] = r.m[3] = r.m[4] = r.m[5] = 0.0;
      r.m[6] = 1.0;
      r.m[7] = 0.0;
      r.flags |= GF_D1_IDENT;
      if (is_one_c(g0)) {
        r.kind = GK_NOP;  // e.g. CPHASE under reference semantics
        r.flags |= GF_D0_IDENT;
        r.flops_per_amp = 0.0;
      } else {
        r.kind = GK_DIAG;
        r.flops_per_amp = 6.0 * 0.25;
      }
      return r;
    }
  } else if (is_zero_c(g1) && is_zero_c(g2)) {
    r.kind = GK_DIAG;
    if (is_one_c(g0)) r.flags |= GF_D0_IDENT;
    if (is_one_c(g3)) r.flags |= GF_D1_IDENT;
    int touched = ((r.flags & GF_D0_IDENT) ? 0 : 1) + ((r.flags & GF_D1_IDENT) ? 0 : 1);
    if (touched == 0) r.kind = GK_NOP;
    r.flops_per_amp = 6.0 * 0.5 * touched * ctrl_frac;
    return r;
  }

  const bool all_real = is_zero(m[1]) && is_zero(m[3]) && is_zero(m[5]) && is_zero(m[7]);
  if (all_real && is_zero(g0[0]) && is_zero(g3[0]) && g1[0] == 1.0 && g2[0] == 1.0) {
    r.kind = GK_PAIR_SWAP;
    r.flops_per_amp = 0.5 * ctrl_frac;
  } else if (all_real && g0[0] == g2[0] && g3[0] == -g1[0]) {
    r.kind = GK_PAIR_HSYM;
    r.flops_per_amp = 4.0 * ctrl_frac;
  } else if (all_real) {
    r.kind = GK_PAIR_REAL;
    r.flops_per_amp = 6.0 * ctrl_frac;
  } else {
    r.kind = GK_PAIR_GENERIC;
    r.flops_per_amp = 14.0 * ctrl_frac;
  }
  if (r.flags & GF_ROW0_ONLY) r.flops_per_amp *= 0.5;
  return r;
}

static bool is_pairing(uint8_t kind) {
  return kind == GK_PAIR_GENERIC || kind == GK_PAIR_REAL || kind == GK_PAIR_HSYM ||
         kind == GK_PAIR_SWAP;
}

// Assigns the 12 tile bits to roles for one segment.  `regbits` (tile-bit indices) become the
// register roles.  Five of the rest become lane bits: a gate whose control (or diagonal target) is
// a lane bit runs with half of the warp idle, so the five bits with the least such work
// (`lane_cost`, flops per amplitude) are chosen; among equals the lowest bits win, which gives
// lane bits = tile bits 0..4 = fully coalesced 512-byte global rows per warp, and `force_low`
// demands exactly that (first / last segment of a pass).  Lane roles 0..2 get three bits with
// distinct residues mod 3 when possible so that, with the XOR swizzle used by the kernel
// (slot ^ (slot>>3) ^ (slot>>6) ^ (slot>>9) on the low 3 bits), every quarter-warp of a 16-byte
// shared-memory access hits 8 distinct bank groups.  Returns the lane cost of the choice.
static double assign_roles(int T, const std::vector<int> &regbits,
                           const double lane_cost[QCS_TILE_BITS], bool force_low,
                           uint8_t role_tilebit[QCS_TILE_BITS]) {
  bool is_reg[QCS_TILE_BITS] = {false};
  for (int b : regbits) is_reg[b] = true;
  std::vector<int> rest;
  for (int b = 0; b < T; b++)
    if (!is_reg[b]) rest.push_back(b);
  const int n = (int)rest.size();
  auto residues_ok = [&](unsigned subset) {
    bool r[3] = {false, false, false};
    for (int i = 0; i < n; i++)
      if ((subset >> i) & 1u) r[rest[i] % 3] = true;
    return r[0] && r[1] && r[2];
  };
  unsigned best = 0;
  double best_cost = 0.0;
  int best_bad = 0, best_sum = 0;
  bool have = false;
  for (unsigned subset = 0; subset < (1u << n); subset++) {
    if (__builtin_popcount(subset) != QCS_LANE_BITS) continue;
    double cost = 0.0;
    int sum = 0;
    bool low_only = true;
    for (int i = 0; i < n; i++)
      if ((subset >> i) & 1u) {
        cost += lane_cost[rest[i]];
        sum += rest[i];
        if (rest[i] >= QCS_LANE_BITS) low_only = false;
      }
    if (force_low && !low_only) continue;
    const int bad = residues_ok(subset) ? 0 : 1;
    if (!have || cost < best_cost - 1e-9 ||
        (cost < best_cost + 1e-9 && (bad < best_bad || (bad == best_bad && sum < best_sum)))) {
      have = true;
      best = subset;
      best_cost = cost;
      best_bad = bad;
      best_sum = sum;
    }
  }
  if (!have) return assign_roles(T, regbits, lane_cost, false, role_tilebit);  // a register bit below 5
  std::vector<int> lanes, warps;
  for (int i = 0; i < n; i++) (((best >> i) & 1u) ? lanes : warps).push_back(rest[i]);
  std::vector<int> lane_lo;
  bool used_res[3] = {false, false, false}, taken[QCS_TILE_BITS] = {false};
  for (int b : lanes)
    if ((int)lane_lo.size() < 3 && !used_res[b % 3]) {
      lane_lo.push_back(b);
      used_res[b % 3] = true;
      taken[b] = true;
    }
  for (int b : lanes)
    if ((int)lane_lo.size() < 3 && !taken[b]) {
      lane_lo.push_back(b);
      taken[b] = true;
    }
  int role = 0;
  for (int b : lane_lo) role_tilebit[role++] = (uint8_t)b;
  for (int b : lanes)
    if (!taken[b]) role_tilebit[role++] = (uint8_t)b;          // lane roles 3, 4
  for (int b : warps) role_tilebit[role++] = (uint8_t)b;      // warp roles
  for (int b : regbits) role_tilebit[role++] = (uint8_t)b;
  return best_cost;
}

// math=fast cost of a gate in FP64 instructions per amplitude (fused multiply-adds; a controlled
// phase is one table lookup shared by four fan entries plus its share of one complex multiply).
static double fast_flops(const PhysGate &g) {
  const double ctrl_frac = g.cpos >= 0 ? 0.5 : 1.0;
  const double row = (g.c.flags & GF_ROW0_ONLY) ? 0.5 : 1.0;
  switch (g.c.kind) {
    case GK_PAIR_HSYM: return 3.0 * ctrl_frac * row;
    case GK_PAIR_REAL: return 4.0 * ctrl_frac * row;
    case GK_PAIR_GENERIC: return 8.0 * ctrl_frac * row;
    case GK_DIAG: {
      const int touched = ((g.c.flags & GF_D0_IDENT) ? 0 : 1) + ((g.c.flags & GF_D1_IDENT) ? 0 : 1);
      if (g.cpos >= 0 && touched == 1) return 0.5;
      return 4.0 * 0.5 * touched * ctrl_frac;
    }
    default: return g.c.flops_per_amp;
  }
}

// math=fast layout of a controlled-phase fan (header at record h, K entries behind it): the kernel
// multiplies the phases of the entries a thread takes part in into ONE factor before it touches an
// amplitude, four entries per step.  Group i = entries 4i..4i+3; the 16 products over the subsets of
// the group (product j = phases of the entries whose bit is set in j; product 0 = 1) replace the
// matrix slots of the group's records: product j sits in record 4i + (j >> 2), m[2 (j & 3)..+1].
// A last group of r < 4 entries needs 2^r <= 4 r slots, so everything fits in place.  Products are
// formed in long double and rounded once.
static void write_fan_tables(PassParams &pp, int h) {
  const int K = pp.gate[h].tsel;
  long double ph[QCS_MAX_FAN_ENTRIES][2];
  for (int k = 0; k < K; k++) {
    ph[k][0] = pp.gate[h + 1 + k].m[6];
    ph[k][1] = pp.gate[h + 1 + k].m[7];
  }
  for (int k = 0; k < K; k++) std::memset(pp.gate[h + 1 + k].m, 0, sizeof(pp.gate[0].m));
  for (int i = 0; 4 * i < K; i++) {
    const int r = std::min(4, K - 4 * i);
    for (int j = 0; j < (1 << r); j++) {
      long double re = 1.0L, im = 0.0L;
      for (int b = 0; b < r; b++)
        if ((j >> b) & 1) {
          const long double nre = re * ph[4 * i + b][0] - im * ph[4 * i + b][1];
          const long double nim = re * ph[4 * i + b][1] + im * ph[4 * i + b][0];
          re = nre;
          im = nim;
        }
      DGate &rec = pp.gate[h + 1 + 4 * i + (j >> 2)];
      rec.m[2 * (j & 3)] = (double)re;
      rec.m[2 * (j & 3) + 1] = (double)im;
    }
  }
}

namespace {

struct PassBuilder {
  const PlannerConfig &cfg;
  std::vector<int> tile;          // physical positions in the tile (unsorted until close)
  std::vector<PhysGate> gates;
  std::vector<int> ids;           // input indices of everything add()ed, NOPs included
  int n_api = 0;
  double flops = 0.0;

  explicit PassBuilder(const PlannerConfig &c) : cfg(c) {
    for (int p = 0; p < cfg.fixed_low; p++) tile.push_back(p);
  }
  bool has(int pos) const { return std::find(tile.begin(), tile.end(), pos) != tile.end(); }

  // one of the five lowest positions of the tile as it stands (a conservative stand-in for "lane
  // bit of the I/O segments": a lower position may still join the tile later)
  bool low5(int pos) const {
    int below = 0;
    for (int p : tile)
      if (p < pos) below++;
    return below < QCS_LANE_BITS;
  }

  // Counts the segments the current gate list needs (plus the I/O segments).
  int segments_needed(const PhysGate *extra) const {
    int segs = 1, in_r = 0;
    int r[QCS_MAX_REG_BITS];
    bool first_low = false, last_low = false;
    auto feed = [&](const PhysGate &g) {
      if (!is_pairing(g.c.kind)) return;
      for (int k = 0; k < in_r; k++)
        if (r[k] == g.tpos) return;
      if (in_r == cfg.reg_bits) {
        segs++;
        in_r = 0;
        last_low = false;
      }
      r[in_r++] = g.tpos;
      if (low5(g.tpos)) {
        if (segs == 1) first_low = true;
        last_low = true;
      }
    };
    for (const auto &g : gates) feed(g);
    if (extra) feed(*extra);
    if (cfg.direct_io) segs += (first_low ? 1 : 0) + (last_low ? 1 : 0);
    return segs;
  }

  bool fits(const PhysGate &g) const {
    if ((int)gates.size() + 1 > QCS_MAX_PASS_GATES) return false;
    if (!gates.empty() && flops + g.c.flops_per_amp > cfg.pass_flops_budget) return false;
    if (is_pairing(g.c.kind) && !has(g.tpos)) {
      if ((int)tile.size() >= cfg.tile_bits_max) return false;
      // A pass that already carries enough arithmetic to be FP64-bound gains nothing from more
      // qubits (its memory sweep is hidden anyway) and loses the finer CTA interleaving of a
      // smaller tile: stop growing the tile, not the pass.
      if ((int)tile.size() >= cfg.tile_bits_min && flops >= cfg.compute_bound_flops) return false;
    }
    if (segments_needed(&g) > QCS_MAX_PASS_SEGMENTS) return false;
    return true;
  }

  void add(const PhysGate &g, int id) {
    n_api++;
    ids.push_back(id);
    if (g.c.kind == GK_NOP) return;
    if (is_pairing(g.c.kind) && !has(g.tpos)) tile.push_back(g.tpos);
    gates.push_back(g);
    flops += g.c.flops_per_amp;
  }

  PassPlan close() {
    PassPlan plan;
    std::memset(&plan.params, 0, sizeof(plan.params));
    // The smallest tile that holds every pairing target; filled up with the lowest unused local
    // positions (keeps rows long).
    const int T = std::max(cfg.tile_bits_min, (int)tile.size());
    for (int p = cfg.fixed_low; (int)tile.size() < T && p < cfg.n_local; p++)
      if (!has(p)) tile.push_back(p);
    std::sort(tile.begin(), tile.end());
    plan.tile_positions = tile;
    PassParams &pp = plan.params;
    pp.shard_base = cfg.shard_base;
    pp.tile_bits = T;
    uint64_t tile_mask = 0;
    for (int b = 0; b < T; b++) {
      pp.tile_pos[b] = (uint8_t)tile[b];
      tile_mask |= 1ull << tile[b];
    }
    pp.nontile_mask = ((cfg.n_local >= 64 ? ~0ull : ((1ull << cfg.n_local) - 1))) & ~tile_mask;
    auto tilebit_of = [&](int pos) -> int {
      for (int b = 0; b < T; b++)
        if (tile[b] == pos) return b;
      return -1;
    };

    // Cut into segments: a segment's register set grows on demand up to 4 bits.
    struct Seg { std::vector<int> regbits; int begin, end; };
    std::vector<Seg> segs;
    segs.push_back(Seg{{}, 0, 0});
    for (int gi = 0; gi < (int)gates.size(); gi++) {
      const PhysGate &g = gates[gi];
      if (is_pairing(g.c.kind)) {
        int tb = tilebit_of(g.tpos);
        Seg &cur = segs.back();
        if (std::find(cur.regbits.begin(), cur.regbits.end(), tb) == cur.regbits.end()) {
          if ((int)cur.regbits.size() == cfg.reg_bits) {
            cur.end = gi;
            segs.push_back(Seg{{tb}, gi, gi});
          } else {
            cur.regbits.push_back(tb);
          }
        }
      }
      segs.back().end = gi + 1;
    }
    // Complete register sets with spare tile bits (highest first, never below 5
    // unless forced) so that lanes keep the low bits.
    for (auto &s : segs) {
      for (int b = T - 1; b >= 0 && (int)s.regbits.size() < cfg.reg_bits; b--)
        if (std::find(s.regbits.begin(), s.regbits.end(), b) == s.regbits.end())
          s.regbits.push_back(b);
    }
    // Work a tile bit would waste as a lane bit of segment s: gates that test it once per thread
    // (a control or a diagonal target that is not a register bit) run with half the warp idle.
    auto lane_costs = [&](const Seg &s, double cost[QCS_TILE_BITS]) {
      for (int b = 0; b < T; b++) cost[b] = 0.0;
      auto is_reg = [&](int tb) {
        return std::find(s.regbits.begin(), s.regbits.end(), tb) != s.regbits.end();
      };
      for (int gi = s.begin; gi < s.end; gi++) {
        const PhysGate &g = gates[gi];
        const int cb = g.cpos >= 0 ? tilebit_of(g.cpos) : -1;
        if (cb >= 0 && !is_reg(cb)) cost[cb] += g.c.flops_per_amp;
        const int tb = tilebit_of(g.tpos);
        if (g.c.kind == GK_DIAG && tb >= 0 && !is_reg(tb)) cost[tb] += g.c.flops_per_amp;
      }
    };
    // The first and the last segment load / store global memory with their lane bits as the fastest
    // index, so they must keep tile bits 0..4 on the lanes.  Where that is impossible (a register
    // bit below 5) or wasteful (low bits busy as controls / diagonal targets), a gate-less segment
    // is added for the I/O and the gates get the lanes that suit them: one more transposition
    // through shared memory, worth about kSegmentCost flops per amplitude.
    const double kSegmentCost = 6.0;
    if (cfg.direct_io) {
      auto wants_own_io = [&](const Seg &s) {
        for (int b : s.regbits)
          if (b < QCS_LANE_BITS) return true;
        if ((int)segs.size() + 2 > QCS_MAX_PASS_SEGMENTS) return false;
        double cost[QCS_TILE_BITS];
        uint8_t scratch[QCS_TILE_BITS];
        lane_costs(s, cost);
        return assign_roles(T, s.regbits, cost, true, scratch) >
               assign_roles(T, s.regbits, cost, false, scratch) + kSegmentCost;
      };
      std::vector<int> io_regs;
      for (int k = cfg.reg_bits; k >= 1; k--) io_regs.push_back(T - k);
      const bool front = wants_own_io(segs.front()), back = wants_own_io(segs.back());
      if (front) segs.insert(segs.begin(), Seg{io_regs, 0, 0});
      if (back) {
        int e = (int)gates.size();
        segs.push_back(Seg{io_regs, e, e});
      }
    }
    pp.n_segments = (int)segs.size();
    pp.reg_bits = cfg.reg_bits;
    auto is_cphase = [](const PhysGate &g) {
      return g.c.kind == GK_DIAG && g.cpos >= 0 && (g.c.flags & GF_D0_IDENT) &&
             !(g.c.flags & GF_D1_IDENT);
    };
    // symbolic op id -> case label of the generated switch for this register-file shape
    auto case_label = [&](int sym) -> uint16_t {
      return cfg.reg_bits == 3 ? QCS3_CASE_LABEL[sym] : QCS4_CASE_LABEL[sym];
    };
    int out_n = 0, n_fans = 0;
    for (int si = 0; si < (int)segs.size(); si++) {
      DSegment &ds = pp.seg[si];
      {
        double cost[QCS_TILE_BITS];
        lane_costs(segs[si], cost);
        const bool io = cfg.direct_io && (si == 0 || si == (int)segs.size() - 1);
        assign_roles(T, segs[si].regbits, cost, io, ds.role_tilebit);
      }
      ds.gate_begin = (uint16_t)out_n;
      auto reg_of = [&](int pos) -> int {
        const int tbit = pos >= 0 ? tilebit_of(pos) : -1;
        for (int k = 0; k < cfg.reg_bits; k++)
          if (tbit >= 0 && segs[si].regbits[k] == tbit) return k;
        return -1;
      };
      // writes the record of one gate
      auto emit_record = [&](const PhysGate &g) {
        const int treg = reg_of(g.tpos), creg = reg_of(g.cpos);
        DGate &dg = pp.gate[out_n++];
        std::memset(&dg, 0, sizeof(dg));
        std::memcpy(dg.m, g.c.m, sizeof(dg.m));
        dg.kind = g.c.kind;
        dg.flags = g.c.flags;
        dg.tpos = (int8_t)g.tpos;
        dg.cpos = (int8_t)g.cpos;
        dg.treg_creg = (uint8_t)((treg + 1) | ((creg + 1) << 4));
        // a control that is not a register bit is tested once per thread on its own basis index
        const bool thread_ctl = g.cpos >= 0 && creg < 0;
        dg.csel = thread_ctl ? (uint8_t)g.cpos : 0xFF;
        dg.tsel = 0xFF;
        const int row0 = (g.c.flags & GF_ROW0_ONLY) ? 1 : 0;
        int op = QCS_OP_NONE;
        switch (g.c.kind) {
          case GK_PAIR_HSYM: op = qcs_op_id_pair(0, row0, treg, creg); break;
          case GK_PAIR_REAL: op = qcs_op_id_pair(1, row0, treg, creg); break;
          case GK_PAIR_SWAP: op = qcs_op_id_pair(2, row0, treg, creg); break;
          case GK_PAIR_GENERIC: op = qcs_op_id_pair(3, row0, treg, creg); break;
          case GK_DIAG: {
            const int halves = ((g.c.flags & GF_D0_IDENT) ? 0 : 1) | ((g.c.flags & GF_D1_IDENT) ? 0 : 2);
            if (!halves) break;
            if (treg < 0) {
              op = qcs_op_id_diag_free(creg, halves);
              dg.tsel = (uint8_t)g.tpos;
            } else {
              op = qcs_op_id_diag_reg(treg, creg, halves);
            }
            break;
          }
          default: break;
        }
        if (op != QCS_OP_NONE && thread_ctl) op |= QCS_OP_TCTL;
        dg.op = case_label(op);
      };
      // writes a fan header for the `count` entries that follow; uniform = every control outside the tile
      auto emit_fan_header = [&](const PhysGate &first, int count, bool consecutive, bool uniform) {
        const int treg = reg_of(first.tpos);
        const int h = out_n;
        DGate &hd = pp.gate[out_n++];
        std::memset(&hd, 0, sizeof(hd));
        hd.op = case_label((uniform ? QCS_OP_UFAN_BASE : QCS_OP_FAN_BASE) + treg + 1);
        hd.kind = GK_DIAG;
        hd.flags = GF_FAN_HEADER;
        hd.csel = uniform ? (uint8_t)pp.n_ufans : (consecutive ? (uint8_t)first.cpos : 0xFF);
        hd.tsel = (uint8_t)count;  // number of entries that follow
        hd.tpos = (int8_t)first.tpos;
        hd.cpos = (int8_t)first.cpos;
        hd.treg_creg = (uint8_t)(treg + 1);
        hd.pad[0] = consecutive ? 1 : 0;  // uniform fans: entry k is controlled by position cpos + k
        if (uniform) pp.ufan_header[pp.n_ufans++] = (uint16_t)h;
        n_fans++;
      };
      // Controlled-phase fan: a run of consecutive controlled diag(1, e^{ia}) gates on one target
      // whose controls are not register bits gets a header so the kernel walks it in a tight loop
      // (QFT: one or two fans per round).  Entries controlled by a register bit stay ordinary gates
      // and split the run.
      auto fan_entry = [&](const PhysGate &x) { return is_cphase(x) && reg_of(x.cpos) < 0; };
      for (int gi = segs[si].begin; gi < segs[si].end; gi++) {
        const PhysGate &g = gates[gi];
        if (fan_entry(g) && n_fans + (cfg.fast_math ? 2 : 1) <= QCS_MAX_PASS_FANS) {
          int run = 1;
          while (gi + run < segs[si].end && fan_entry(gates[gi + run]) &&
                 gates[gi + run].tpos == g.tpos && run < QCS_MAX_FAN_ENTRIES)
            run++;
          if (run >= 2 && !cfg.fast_math) {
            // The fast walk needs entry k to be controlled by position cpos + k (the thread's mask is one
            // shift of its basis index); any other order makes the kernel gather the mask entry by
            // entry, ~10 instructions each.  On a sharded engine the layout drifts (after a remap
            // logical qubit 30 may sit below qubit 29), so a QFT's run of controls 3..30 stops being
            // ascending at its very end: the whole 28-entry fan fell back to the gather and the first
            // QFT pass took 39.5 instead of 31 ms (profiles/r2j_bench_2gpu.json).  Order must be kept
            // (bit-exactness), so the run is cut where it stops ascending by one: every ascending
            // stretch of two or more entries is a consecutive fan, a stretch of one an ordinary gate.
            int k0 = 0;
            while (k0 < run) {
              int len = 1;
              while (k0 + len < run && gates[gi + k0 + len].cpos == gates[gi + k0].cpos + len) len++;
              if (len >= 2 && n_fans + 1 <= QCS_MAX_PASS_FANS) {
                emit_fan_header(gates[gi + k0], len, true, false);
                for (int k = 0; k < len; k++) emit_record(gates[gi + k0 + k]);
              } else {
                for (int k = 0; k < len; k++) emit_record(gates[gi + k0 + k]);
              }
              k0 += len;
            }
            gi += run - 1;
            continue;
          }
          if (run >= 2) {
            // math=fast: the entries commute (all diagonal), so the run is sorted into the entries a
            // thread decides for itself (control on a tile bit) and the ones the whole CTA shares
            // (control outside the tile: common.h QCS_OP_UFAN_BASE); each part of two or more entries
            // becomes a fan, single entries stay ordinary gates
            std::vector<int> own, shared;
            for (int k = 0; k < run; k++)
              (tilebit_of(gates[gi + k].cpos) >= 0 ? own : shared).push_back(gi + k);
            // (the entries commute and the mode merges their phases anyway: sorted by control position
            // the shift-based mask applies whenever the positions are contiguous, whatever order the
            // circuit -- or a drifted layout -- listed them in)
            auto by_cpos = [&](int a, int b) { return gates[a].cpos < gates[b].cpos; };
            std::sort(own.begin(), own.end(), by_cpos);
            std::sort(shared.begin(), shared.end(), by_cpos);
            // Thread-table fan (common.h QCS_OP_TFAN_BASE): what the own entries do to a thread depends
            // on its lane and warp bits only, so the product of their phases over the controls a
            // thread has set is tabulated here, once for every CTA of the pass: one coalesced load per
            // thread replaces the walk over product tables (per-thread constant loads and ~25
            // instructions per group of four entries -- most of the first QFT pass's instructions).
            const bool table = cfg.thread_tables >= 2 && (int)own.size() >= cfg.thread_tables && pp.n_thread_tables < QCS_MAX_PASS_FANS &&
                               n_fans + 2 <= QCS_MAX_PASS_FANS && out_n + 2 + run <= QCS_MAX_PASS_GATES + QCS_MAX_PASS_FANS;
            if (table) {
              const int CT = 1 << (T - cfg.reg_bits);
              const int treg = reg_of(g.tpos);
              const int h = out_n;
              DGate &hd = pp.gate[out_n++];
              std::memset(&hd, 0, sizeof(hd));
              hd.op = case_label(QCS_OP_TFAN_BASE + treg + 1);
              hd.kind = GK_DIAG;
              hd.flags = GF_TFAN_HEADER;
              hd.tsel = (uint8_t)pp.n_thread_tables;
              hd.csel = shared.size() >= 2 ? (uint8_t)pp.n_ufans : 0xFF;  // the uniform fan emitted below gets this slot
              hd.tpos = (int8_t)g.tpos;
              hd.cpos = -1;
              hd.treg_creg = (uint8_t)(treg + 1);
              hd.pad[2] = (uint8_t)own.size();
              n_fans++;
              // W[tid] = product of the phases of the own entries whose control bit thread tid has set
              // (tid -> tile index through the segment's role tables, exactly as the kernel maps it)
              const size_t base = plan.thread_tables.size();
              plan.thread_tables.resize(base + 2 * (size_t)CT);
              for (int tid = 0; tid < CT; tid++) {
                uint32_t tb = 0;
                for (int role = 0; role < T - cfg.reg_bits; role++)
                  if ((tid >> role) & 1) tb |= 1u << ds.role_tilebit[role];
                long double re = 1.0L, im = 0.0L;
                for (int id : own) {
                  const int cb = tilebit_of(gates[id].cpos);
                  if (!((tb >> cb) & 1u)) continue;
                  const long double pr = gates[id].c.m[6], pi = gates[id].c.m[7];
                  const long double nre = re * pr - im * pi, nim = re * pi + im * pr;
                  re = nre;
                  im = nim;
                }
                plan.thread_tables[base + 2 * (size_t)tid] = (double)re;
                plan.thread_tables[base + 2 * (size_t)tid + 1] = (double)im;
              }
              pp.n_thread_tables++;
              for (int id : own) emit_record(gates[id]);
              if (shared.size() >= 2) {
                bool consecutive = true;
                for (size_t k = 1; k < shared.size(); k++)
                  if (gates[shared[k]].cpos != gates[shared[0]].cpos + (int)k) consecutive = false;
                emit_fan_header(gates[shared[0]], (int)shared.size(), consecutive, true);
                for (int id : shared) emit_record(gates[id]);
                pp.gate[h].pad[1] = (uint8_t)(own.size() + 1 + shared.size());
              } else {
                pp.gate[h].pad[1] = (uint8_t)own.size();
                for (int id : shared) emit_record(gates[id]);  // a single out-of-tile control: an ordinary gate
              }
              gi += run - 1;
              continue;
            }
            for (int part = 0; part < 2; part++) {
              const std::vector<int> &ids = part == 0 ? own : shared;
              if (ids.size() >= 2) {
                bool consecutive = true;
                for (size_t k = 1; k < ids.size(); k++)
                  if (gates[ids[k]].cpos != gates[ids[0]].cpos + (int)k) consecutive = false;
                emit_fan_header(gates[ids[0]], (int)ids.size(), consecutive, part == 1);
              }
              for (int id : ids) emit_record(gates[id]);
            }
            gi += run - 1;
            continue;
          }
        }
        emit_record(g);
      }
      ds.gate_end = (uint16_t)out_n;
    }
    pp.n_gates = out_n;
    plan.n_fan_headers = n_fans;
    if (cfg.fast_math)
      for (int h = 0; h < out_n; h++)
        if (pp.gate[h].flags & GF_FAN_HEADER) write_fan_tables(pp, h);
    // ---- tables that spare the kernel its per-tile index arithmetic ----------------------
    auto swz = [](uint32_t slot) { return slot ^ (((slot >> 3) ^ (slot >> 6) ^ (slot >> 9)) & 7u); };
    const int thread_roles = T - cfg.reg_bits;
    for (int si = 0; si < (int)segs.size(); si++) {
      DSegment &ds = pp.seg[si];
      for (int g = 0; g < 3; g++)
        for (int v = 0; v < 8; v++) {
          uint32_t tb = 0;
          for (int j = 0; j < 3; j++) {
            const int role = 3 * g + j;
            if (role < thread_roles && ((v >> j) & 1)) tb |= 1u << ds.role_tilebit[role];
          }
          ds.tb_lut[g][v] = (uint16_t)tb;
        }
      for (int k = 0; k < cfg.reg_bits; k++) {
        const uint32_t off = 1u << ds.role_tilebit[thread_roles + k];
        ds.toff[k] = (uint16_t)off;
        ds.stoff[k] = (uint16_t)swz(off);
      }
    }
    for (int si = 0; si < (int)segs.size(); si++) {
      DSegment &ds = pp.seg[si];
      for (int g = 0; g < 3; g++)
        for (int v = 0; v < 8; v++) {
          uint64_t off = 0;
          for (int b = 0; b < T; b++)
            if ((ds.tb_lut[g][v] >> b) & 1) off |= 1ull << pp.tile_pos[b];
          ds.gb_lut[g][v] = off;
        }
    }
    for (int which = 0; which < 2; which++) {
      const DSegment &ds = pp.seg[which == 0 ? 0 : (int)segs.size() - 1];
      for (int k = 0; k < cfg.reg_bits; k++)
        pp.goff[which][k] = 1ull << pp.tile_pos[ds.role_tilebit[thread_roles + k]];
    }
    {
      int n_runs = 0, src = 0, p = 0;
      while (p < cfg.n_local) {
        if (!((pp.nontile_mask >> p) & 1ull)) { p++; continue; }
        int len = 0;
        while (p + len < cfg.n_local && ((pp.nontile_mask >> (p + len)) & 1ull)) len++;
        pp.tile_run[n_runs].src = (uint8_t)src;
        pp.tile_run[n_runs].len = (uint8_t)len;
        pp.tile_run[n_runs].dst = (uint8_t)p;
        n_runs++;
        src += len;
        p += len;
      }
      pp.n_tile_runs = n_runs;  // <= 8: seven tile bits above position 4 split the rest into <= 8 runs
    }
    plan.n_gates_api = n_api;
    plan.api_ids = ids;
    plan.flops_per_amp = flops;
    return plan;
  }
};

}  // namespace

// ---- math=fast: commutation-aware scheduling ------------------------------------------------------
// Two gates commute when every qubit they share is a control or a diagonal target of BOTH (both are
// then block-diagonal in those qubits and act on disjoint qubits inside each block).  In floating
// point the two orders round differently, so the bit-exact mode never uses this; math=fast does.
// A pass becomes a subsequence of the queue: what can run with pairing targets inside one tile,
// segment after segment, following the circuit several layers deep where the in-order cut would
// stop at the first qubit outside the tile (a brickwork layer touches every qubit).
namespace {

// One in-order sweep over the pending gates: which of them can run now if only positions in `rmask`
// may be paired?  A gate that cannot run leaves its qubits pending: a pending pairing use blocks
// every later use of that position, a pending diagonal-like use blocks later pairing uses only.
struct Sweep {
  std::vector<int> picked;  // indices into the gate list, ascending
  int pairing = 0;          // pairing gates among them
};

// exact: only gates that move amplitudes without arithmetic (GK_PAIR_SWAP) may pass a pending gate or be
// passed by one -- a gate that rounds never overtakes another gate that rounds, not even on other qubits
// (A then B and B then A round different intermediates).
Sweep sweep_executable(const std::vector<PhysGate> &g, const std::vector<char> &done, size_t first,
                       uint64_t rmask, int n_positions, size_t window, bool collect, bool exact = false) {
  Sweep s;
  uint8_t state[64] = {0};
  int n_full = 0;
  bool rounding_gate_pending = false;
  size_t seen = 0, last_pick = 0;
  const size_t patience = 16 * (size_t)n_positions + 16;  // gates scanned without a pick before giving up
  for (size_t i = first; i < g.size() && seen < window && n_full < n_positions; i++) {
    if (done[i]) continue;
    if (seen - last_pick > patience) break;
    seen++;
    const PhysGate &x = g[i];
    if (x.c.kind == GK_NOP) {
      if (collect) s.picked.push_back((int)i);
      continue;
    }
    const bool pairing = is_pairing(x.c.kind);
    bool ok = pairing ? (((rmask >> x.tpos) & 1ull) && state[x.tpos] == 0) : state[x.tpos] <= 1;
    if (ok && x.cpos >= 0 && state[x.cpos] > 1) ok = false;
    const bool moves_only = x.c.kind == GK_PAIR_SWAP;
    if (ok && exact && !moves_only && rounding_gate_pending) ok = false;
    if (!ok && !moves_only) rounding_gate_pending = true;
    if (ok) {
      if (collect) s.picked.push_back((int)i);
      s.pairing += pairing ? 1 : 0;
      last_pick = seen;
      continue;
    }
    const uint8_t t_use = pairing ? 2 : 1;
    if (state[x.tpos] < t_use) {
      if (t_use == 2) n_full++;
      state[x.tpos] = t_use;
    }
    if (x.cpos >= 0 && state[x.cpos] < 1) state[x.cpos] = 1;
  }
  return s;
}

std::vector<PassPlan> plan_passes_reordered(const std::vector<PhysGate> &gates, const PlannerConfig &cfg_in) {
  PlannerConfig cfg = cfg_in;
  const bool exact = !cfg.fast_math;
  // the tile is chosen here, not by the builder's growth rule (stopping heavy passes at 10 bits as the
  // in-order planner does was measured: QFT unchanged, 31-qubit random circuit 3 % slower)
  if (!exact) cfg.compute_bound_flops = 1e30;
  const int n_positions = cfg.n_local + cfg.rank_bits;
  const size_t window = 4096;
  std::vector<PassPlan> out;
  std::vector<char> done(gates.size(), 0);
  size_t first = 0, left = gates.size();
  std::vector<int> carry_ids;
  while (left > 0) {
    while (first < gates.size() && done[first]) first++;
    std::unique_ptr<PassBuilder> b(new PassBuilder(cfg));
    uint64_t tile_mask = 0;
    int tile_size = 0;
    for (int p = 0; p < cfg.fixed_low && p < cfg.n_local; p++) {
      tile_mask |= 1ull << p;
      tile_size++;
    }
    bool closed = false;
    size_t taken = 0;
    for (int seg = 0; seg < cfg.reorder_segments && !closed; seg++) {
      // register set of this segment: grown one position at a time by marginal gain
      uint64_t rmask = 0;
      int base_pairing = 0;
      for (int k = 0; k < cfg.reg_bits; k++) {
        int best = -1, best_gain = 0;
        bool best_in_tile = false;
        for (int q = 0; q < cfg.n_local; q++) {
          if ((rmask >> q) & 1ull) continue;
          const bool in_tile = (tile_mask >> q) & 1ull;
          if (!in_tile && tile_size + __builtin_popcountll(rmask & ~tile_mask) >= cfg.tile_bits_max) continue;
          const int gain = sweep_executable(gates, done, first, rmask | (1ull << q), n_positions, window, false, exact).pairing -
                           base_pairing;
          // ties: a position the tile already holds (keeps the tile small / leaves room), then the lowest
          if (gain > best_gain || (gain == best_gain && gain > 0 && in_tile && !best_in_tile)) {
            best = q;
            best_gain = gain;
            best_in_tile = in_tile;
          }
        }
        if (best < 0) break;
        rmask |= 1ull << best;
        base_pairing += best_gain;
      }
      Sweep sw = sweep_executable(gates, done, first, rmask, n_positions, window, true, exact);
      if (sw.picked.empty()) break;
      size_t added = 0;
      for (int i : sw.picked) {
        if (gates[(size_t)i].c.kind != GK_NOP && !b->fits(gates[(size_t)i])) {
          closed = true;  // descriptor full: the rest of this sweep stays pending, in order
          break;
        }
        b->add(gates[(size_t)i], i);
        done[(size_t)i] = 1;
        added++;
      }
      taken += added;
      left -= added;
      for (int p : b->tile)
        if (!((tile_mask >> p) & 1ull)) {
          tile_mask |= 1ull << p;
          tile_size++;
        }
      if (rmask == 0) break;  // nothing left to pair in reach: only diagonal gates were pending
    }
    if (taken == 0) {
      // The first pending gate pairs on a position this shard does not hold: nothing more can run
      // under the current layout.  The caller swaps the position in and plans the rest again.
      if (is_pairing(gates[first].c.kind) && gates[first].tpos >= cfg.n_local) break;
      // cannot happen otherwise (the first pending gate is always executable with its own target as
      // register bit); never loop forever on a planner bug
      b.reset(new PassBuilder(cfg));
      b->add(gates[first], (int)first);
      done[first] = 1;
      left--;
    }
    if (!b->gates.empty()) {
      PassPlan plan = b->close();
      plan.n_gates_api += (int)carry_ids.size();  // value-preserving gates met before the first real one
      plan.api_ids.insert(plan.api_ids.begin(), carry_ids.begin(), carry_ids.end());
      carry_ids.clear();
      out.push_back(plan);
    } else {
      carry_ids.insert(carry_ids.end(), b->ids.begin(), b->ids.end());
    }
  }
  if (!carry_ids.empty() && !out.empty()) {
    out.back().n_gates_api += (int)carry_ids.size();
    out.back().api_ids.insert(out.back().api_ids.end(), carry_ids.begin(), carry_ids.end());
  }
  return out;
}

}  // namespace

std::vector<PassPlan> plan_passes(const std::vector<PhysGate> &gates_in, const PlannerConfig &cfg) {
  std::vector<PassPlan> out;
  std::vector<PhysGate> gates = gates_in;
  if (cfg.fast_math)
    for (PhysGate &g : gates) g.c.flops_per_amp = fast_flops(g);
  if (cfg.fast_math && cfg.reorder) return plan_passes_reordered(gates, cfg);
  if (!cfg.fast_math && cfg.reorder_exact && cfg.sem == SEM_CORRECTED) {
    // worth trying only when something in the queue may move at all; kept only when it saves passes
    bool any = false;
    for (const PhysGate &g : gates) any = any || g.c.kind == GK_PAIR_SWAP;
    if (any) {
      PlannerConfig in_order = cfg;
      in_order.reorder_exact = false;
      std::vector<PassPlan> a = plan_passes(gates_in, in_order), b2 = plan_passes_reordered(gates, cfg);
      return b2.size() < a.size() ? b2 : a;
    }
  }
  PassBuilder *b = new PassBuilder(cfg);
  for (size_t i = 0; i < gates.size(); i++) {
    const PhysGate &g = gates[i];
    if (g.c.kind != GK_NOP && !b->fits(g)) {
      out.push_back(b->close());
      delete b;
      b = new PassBuilder(cfg);
    }
    b->add(g, (int)i);
  }
  if (!b->gates.empty()) {
    out.push_back(b->close());
  } else if (b->n_api > 0 && !out.empty()) {
    out.back().n_gates_api += b->n_api;  // trailing NOPs ride on the previous pass
    out.back().api_ids.insert(out.back().api_ids.end(), b->ids.begin(), b->ids.end());
  }
  delete b;
  return out;
}

std::string describe_plan(const std::vector<PassPlan> &passes) {
  std::string s;
  char buf[256];
  static const char *kind_name[] = {"generic", "real", "hsym", "swap", "diag", "nop"};
  for (size_t pi = 0; pi < passes.size(); pi++) {
    const PassParams &pp = passes[pi].params;
    std::snprintf(buf, sizeof(buf), "pass %zu: gates=%d api_gates=%d segments=%d flops/amp=%.1f tile=[",
                  pi, pp.n_gates, passes[pi].n_gates_api, pp.n_segments, passes[pi].flops_per_amp);
    s += buf;
    for (int b = 0; b < pp.tile_bits; b++) {
      std::snprintf(buf, sizeof(buf), "%s%d", b ? "," : "", pp.tile_pos[b]);
      s += buf;
    }
    s += "]\n";
    for (int si = 0; si < pp.n_segments; si++) {
      const DSegment &ds = pp.seg[si];
      std::snprintf(buf, sizeof(buf), "  seg %d: regs(pos)=[", si);
      s += buf;
      for (int k = 0; k < pp.reg_bits; k++) {
        std::snprintf(buf, sizeof(buf), "%s%d", k ? "," : "",
                      pp.tile_pos[ds.role_tilebit[pp.tile_bits - pp.reg_bits + k]]);
        s += buf;
      }
      std::snprintf(buf, sizeof(buf), "] gates %d..%d:", ds.gate_begin, ds.gate_end);
      s += buf;
      for (int gi = ds.gate_begin; gi < ds.gate_end; gi++) {
        const DGate &g = pp.gate[gi];
        if (g.flags & GF_FAN_HEADER) {
          std::snprintf(buf, sizeof(buf), " fan[%d]", (int)g.tsel);
          s += buf;
          continue;
        }
        if (g.flags & GF_TFAN_HEADER) {
          std::snprintf(buf, sizeof(buf), " tfan[%d]", (int)g.pad[2]);
          s += buf;
          continue;
        }
        if (g.cpos >= 0)
          std::snprintf(buf, sizeof(buf), " %s(c%d,t%d)", kind_name[g.kind], g.cpos, g.tpos);
        else
          std::snprintf(buf, sizeof(buf), " %s(t%d)", kind_name[g.kind], g.tpos);
        s += buf;
      }
      s += "\n";
    }
  }
  return s;
}

}  // namespace qcs
