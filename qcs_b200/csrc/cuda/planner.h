// planner.h -- the gate-fusion scheduler (host side, no CUDA dependency).
//
// Replaces the role the north star gives to qc_optimize/qc_run ("a
// gate-fusion scheduler that batches consecutive gates per pass"); the
// reference itself applies every gate eagerly with one full sweep per gate
// (reference src/qcs.c:159-164 -> src/q_gates.c:131-148).  The planner takes
// the deferred queue and cuts it, IN ORDER, into passes and segments
// (common.h).  Gates are never algebraically merged and two gates that both
// round never trade places, so the arithmetic each amplitude sees is the
// reference's, operation by operation; gates that only MOVE amplitudes (X,
// CNOT) may be scheduled past gates they commute with, which changes no bit
// (PlannerConfig::reorder_exact).  (The opt-in math=fast mode gives the rest
// up: FMA arithmetic, fan tables and -- PlannerConfig::reorder -- every
// commuting pair scheduled out of order.)
#pragma once
#include <string>
#include <vector>

#include "common.h"

namespace qcs {

enum Semantics { SEM_REFERENCE = 0, SEM_CORRECTED = 1 };

// A queued gate in LOGICAL qubit numbering.
struct HostGate {
  double m[8];
  int target;
  int control;  // -1: uncontrolled
};

struct Classified {
  uint8_t kind;
  uint8_t flags;
  double m[8];      // possibly normalised copy (identity rows made explicit)
  double flops_per_amp;
};

// Classifies a 2x2 (or controlled 2x2) by its exact zero / one pattern.
Classified classify_gate(const double m[8], bool controlled, Semantics sem);

struct PassPlan {
  PassParams params;
  int n_gates_api;          // API-level gates folded into this pass (incl. NOPs)
  std::vector<int> api_ids; // which ones: indices into plan_passes()'s input, in execution order
                            // (consecutive unless PlannerConfig::reorder)
  int n_fan_headers = 0;    // fan header records among params.n_gates (not gates)
  int n_swaps = 0;          // set by the engine on the executed copy: this pass trades local positions
  int swap_lpos[3] = {-1, -1, -1}, swap_gpos[3] = {-1, -1, -1};  // swap_lpos[i] for global positions swap_gpos[i]
                            // on its stores, all at once (params.n_segments == 0: a stand-alone swap, no pass)
  double flops_per_amp;     // planner's cost estimate
  std::vector<int> tile_positions;
  // math=fast: params.n_thread_tables tables of 2^(tile_bits - reg_bits) complex numbers {re, im}
  // (common.h QCS_OP_TFAN_BASE); the engine copies them to the device in front of the launch
  std::vector<double> thread_tables;
};

struct PlannerConfig {
  int n_local;              // local qubits of this shard (nl)
  int rank_bits;            // log2(world)
  uint64_t shard_base;      // rank << nl
  Semantics sem;
  double pass_flops_budget; // close a pass when its estimated flops/amp exceed this
  bool direct_io;           // true: first/last segment must keep tile bits 0..4 as lane bits
  int reg_bits;             // register-role bits per thread: 4 (16 amplitudes) or 3 (8)
  int tile_bits_max = QCS_TILE_BITS;      // a pass may pair on at most this many positions ...
  int tile_bits_min = QCS_TILE_BITS;      // ... and runs on the smallest tile >= this that holds them
  double compute_bound_flops = 90.0;      // flops per amplitude above which a pass is FP64-bound
  bool fast_math = false;   // math=fast: plan for the fused-multiply-add interpreter -- cost estimates count
                            // FMA instructions, fan entries carry product tables (write_fan_tables)
  bool reorder = false;     // math=fast only: gates that commute (every shared qubit is a control or a diagonal
                            // target of both) may trade places, so a pass is a SUBSEQUENCE of the queue
                            // chosen to keep a tile busy over several circuit layers (plan_passes_reordered)
  bool reorder_exact = false;  // bit-exact mode, corrected semantics: the same scheduler under the one commutation
                            // that is exact in floating point -- a gate that only MOVES amplitudes (X, CNOT:
                            // GK_PAIR_SWAP) trades places with any gate it commutes with; two gates that both
                            // round never do, whatever qubits they act on.  Used when it saves passes.
  int reorder_segments = 8; // segments a reordered pass may spend before the next pass starts
  int thread_tables = 2;  // math=fast: runs of controlled phases with in-tile controls become one table
                            // lookup per thread (common.h QCS_OP_TFAN_BASE) instead of a walk over product tables
  int fixed_low = QCS_LANE_BITS;  // positions 0..fixed_low-1 belong to every tile: global rows of
                            // 16 << fixed_low contiguous bytes; the remaining tile bits are free
};

// Plans a run of gates whose targets (if pairing) are all LOCAL physical
// positions.  `phys_target/phys_control` give physical positions (after the
// logical->physical permutation); controls and diagonal targets may be global.
struct PhysGate {
  Classified c;
  int tpos;
  int cpos;
};

// With PlannerConfig::reorder the input may contain pairing gates on GLOBAL positions: they (and
// whatever depends on them) are left out of every pass -- api_ids tells the caller what ran.
std::vector<PassPlan> plan_passes(const std::vector<PhysGate> &gates,
                                  const PlannerConfig &cfg);

std::string describe_plan(const std::vector<PassPlan> &passes);

}  // namespace qcs
