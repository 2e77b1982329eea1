// common.h -- types shared by the host-side planner and the device kernels.
//
// Vocabulary:
//   amplitude   one complex double, 16 bytes, interleaved {re, im}
//   shard       the 2^nl amplitudes this rank holds (nl = local qubits)
//   position    a bit of the physical basis index; positions < nl are local,
//               positions >= nl are the rank's bits (global)
//   tile        2^T amplitudes (T = 10, 11 or 12, chosen per pass) one CTA holds in
//               registers during a pass; its index bits are `tile_pos` (always contains 0..4)
//   pass        one fused kernel launch: every amplitude is read once and
//               written once, a run of gates is applied in between
//   segment     a stretch of a pass during which the assignment of tile bits
//               to {lane, warp, register} roles is fixed; a gate can only pair
//               amplitudes across a *register* bit, so changing the pairing
//               qubit set means a shared-memory transposition between segments
#pragma once
#include <stdint.h>

#define QCS_TILE_BITS 12     // largest tile: 4096 amplitudes = 64 KiB (array bounds)
#define QCS_MIN_TILE_BITS 10 // smallest tile: 1024 amplitudes = 16 KiB.  Smaller tiles mean more CTAs per SM
                             // (2 / 4 / 8 at 12 / 11 / 10 bits) whose load, compute and store phases
                             // interleave -- measured 88-97 % instead of 67-82 % of the HBM rate per pass --
                             // but fewer qubits per pass; the planner takes the smallest tile a pass fits.
#define QCS_LANE_BITS 5
#define QCS_MAX_REG_BITS 4   // 16 amplitudes per thread (256-thread CTA) or 8 (512-thread CTA)
#define QCS_MAX_PASS_GATES 240
#define QCS_MAX_PASS_SEGMENTS 12

// Arithmetic classes.  Every class performs the reference's operations in the
// reference's order (c_mul then c_add, src/complex.c:23-57) except that
// products with an exact-zero or exact-one matrix component are dropped; the
// dropped terms are exact zeros / exact copies, so values compare equal (==)
// with the oracle (DESIGN.md "Bit-exactness").
enum GateKind : uint8_t {
  GK_PAIR_GENERIC = 0,  // 28 flops per pair
  GK_PAIR_REAL = 1,     // all four imaginary parts are zero: 12 flops per pair
  GK_PAIR_HSYM = 2,     // real, U00==U10 and U11==-U01 (Hadamard): 8 flops per pair
  GK_PAIR_SWAP = 3,     // exact X: moves only
  GK_DIAG = 4,          // U01==U10==0: elementwise, 6 flops per touched amplitude
  GK_NOP = 5            // provably value-preserving (e.g. reference-semantics CPHASE)
};

enum GateFlags : uint8_t {
  GF_ROW0_ONLY = 1,   // reference-semantics controlled gate: only the target=0 row is written
  GF_D0_IDENT = 2,    // diagonal entry 0 is exactly (1,0): untouched
  GF_D1_IDENT = 4,    // diagonal entry 1 is exactly (1,0): untouched
  GF_TFAN_HEADER = 0x40,  // not a gate: header record of a thread-table fan (QCS_OP_TFAN_BASE below)
  GF_FAN_HEADER = 0x80    // not a gate: header record of a controlled-phase fan (below)
};

// Dispatch ids.  The formulas below give SYMBOLIC ids; DGate.op holds the case label the generator
// assigned to that id (fused_ops.h, QCS<R>_CASE_LABEL[symbolic]): the index into the interpreter's
// brx.idx jump table.
//   kind_index: 0 HSYM, 1 REAL, 2 SWAP, 3 GENERIC;  treg/creg: register-role index, creg -1 = none
static inline int qcs_op_id_pair(int kind_index, int row0, int treg, int creg) {
  return ((kind_index * 2 + row0) * 4 + treg) * 5 + (creg + 1);
}
// diagonal whose target is NOT a register bit: the thread picks entry 0 or 1 from bit `tsel` of its
// own basis index; halves: 1 = only entry 0 differs from the identity, 2 = only entry 1, 3 = both
static inline int qcs_op_id_diag_free(int creg, int halves) { return 160 + (creg + 1) * 3 + (halves - 1); }
// diagonal whose target is register bit treg; halves as above
static inline int qcs_op_id_diag_reg(int treg, int creg, int halves) {
  return 175 + (treg * 5 + creg + 1) * 3 + (halves - 1);
}
// Header of a controlled-phase fan, op = QCS_OP_FAN_BASE + treg + 1: the next `tsel` (<= 32) gate
// records are diag(1, e^{ia}) gates on ONE target (tpos; register role treg, -1 = not a register
// bit), each controlled by a bit that is NOT a register bit of the segment (so it has one value for
// all amplitudes of a thread).  csel = control position of the first entry when entry k is
// controlled by position csel + k (the QFT's shape: a thread's participation mask is one shift of
// its own basis index), else 0xFF (the mask is gathered from the entries' cpos).  The interpreter
// walks the entries in a tight loop: one predicate + two constant loads around each FP64 block.
#define QCS_OP_FAN_BASE 240
#define QCS_MAX_FAN_ENTRIES 32
#define QCS_MAX_PASS_FANS 48
// math=fast only.  Header of a UNIFORM fan, op = QCS_OP_UFAN_BASE + treg + 1: same record layout as a
// fan, but every entry is controlled by a position outside the tile (a non-tile local position or a
// rank bit), so the product of the phases is one number for the whole CTA.  One thread per uniform
// fan computes it in the kernel prologue (while the tile's loads are in flight) into shared-memory
// slot `csel`; the interpreter fetches it with one ld.shared.  PassParams::ufan_header lists the
// headers.  (QFT: 20 of the 29 controls of the first target are outside a 10-bit tile.)
#define QCS_OP_UFAN_BASE 248
// math=fast only.  Header of a THREAD-TABLE fan, op = QCS_OP_TFAN_BASE + treg + 1.  The entries of a run of
// controlled phases on one target whose controls are tile positions that are not register bits (lane and
// warp bits) multiply a thread's amplitudes by a product that depends on the thread's lane / warp bits
// only -- the same CT numbers for every CTA of the pass.  The planner builds that table on the host
// (PassPlan::thread_tables, table `tsel`, one complex per thread id; the engine uploads it next to the
// pass), the kernel fetches its entry with one coalesced load, multiplies it with the per-CTA factor of
// the run's out-of-tile controls (uniform-fan slot `csel`, 0xFF: none) and applies the product once.
// pad[1] = records behind the header that the kernel skips: pad[2] own entries (ordinary records, kept
// so that a plan can be read back -- and checked -- gate by gate) and, when csel != 0xFF, the uniform
// fan's header and entries (listed in ufan_header for the prologue, never dispatched).
#define QCS_OP_TFAN_BASE 235
// or-ed into a pairing / diagonal id: the gate's control is a bit the thread tests once for all its
// amplitudes (a lane / warp tile bit, a position outside the tile, a rank bit); csel = its position
#define QCS_OP_TCTL 0x100
#define QCS_OP_NONE 255      // symbolic id of "execute nothing"

// A gate as the device sees it (inside a pass descriptor).  The kernel reads
// m[] and the first four header bytes; the rest documents the plan.
struct DGate {
  double m[8];      // row-major {re,im} x 4
  uint16_t op;      // dispatch id (above)
  uint8_t csel;     // QCS_OP_TCTL ops: physical position of the control bit; fans: see above; else 0xFF
  uint8_t tsel;     // diag_free ops: physical position of the target bit; fans: entry count; else 0xFF
  uint8_t kind;     // GateKind
  uint8_t flags;    // GateFlags
  int8_t tpos;      // physical position of the target bit
  int8_t cpos;      // physical position of the control bit, -1 if none
  uint8_t treg_creg; // (treg + 1) | (creg + 1) << 4
  uint8_t pad[7];
};

// gen_fused_lists.py hard-codes this layout (OFF_* / SIZEOF_DGATE)
static_assert(sizeof(DGate) == 80, "DGate layout is part of the generated interpreter");
static_assert(__builtin_offsetof(DGate, op) == 64 && __builtin_offsetof(DGate, csel) == 66 &&
              __builtin_offsetof(DGate, tsel) == 67 && __builtin_offsetof(DGate, tpos) == 70 &&
              __builtin_offsetof(DGate, cpos) == 71 && __builtin_offsetof(DGate, pad) == 73,
              "DGate layout is part of the generated interpreter");

struct DSegment {
  // role r -> tile bit index; roles 0-4 lane bits, then warp bits, the last reg_bits are register bits
  uint8_t role_tilebit[QCS_TILE_BITS];
  uint16_t gate_begin, gate_end;
  // Host-precomputed per-thread arithmetic (the kernel would otherwise redo it for every tile):
  uint16_t tb_lut[3][8];               // tile-index bits contributed by thread-id bits [3g, 3g+3)
  uint64_t gb_lut[3][8];               // the same bits as element offsets inside the shard (physical positions)
  uint16_t toff[QCS_MAX_REG_BITS];     // tile-index offset of each register role bit
  uint16_t stoff[QCS_MAX_REG_BITS];    // the same, XOR-swizzled
};

// One contiguous run of non-tile positions: tile-number bits [src, src+len) land at position dst.
struct DTileRun {
  uint8_t src, len, dst, pad;
};

// Kernel parameter block of one pass (passed by value, __grid_constant__).
struct PassParams {
  uint64_t shard_base;                 // rank << nl : high bits of every global index in this shard
  uint8_t tile_pos[QCS_TILE_BITS];     // physical position of each tile bit, ascending, [0..4] = 0..4
  int32_t tile_bits;                   // T: tile_pos[0..T) are valid
  uint64_t nontile_mask;               // local positions that are not tile bits
  int32_t n_segments;
  int32_t n_gates;
  int32_t reg_bits;                    // 4: 256 threads x 16 amplitudes, 3: 512 threads x 8
  int32_t n_tile_runs;
  DTileRun tile_run[12];               // tile number -> element offset of the tile (bit deposit as runs)
  // element offsets of the register role bits for the global loads (first segment) and stores (last)
  uint64_t goff[2][QCS_MAX_REG_BITS];  // [first/last][register role bit]
  DSegment seg[QCS_MAX_PASS_SEGMENTS];
  DGate gate[QCS_MAX_PASS_GATES + QCS_MAX_PASS_FANS];  // gates + fan headers
  // math=fast: record indices of the uniform-fan headers (QCS_OP_UFAN_BASE); slot k of the kernel's
  // shared-memory factor table belongs to ufan_header[k]
  uint16_t ufan_header[QCS_MAX_PASS_FANS];
  int32_t n_ufans;
  int32_t n_thread_tables;             // math=fast: tables of 2^(tile_bits - reg_bits) complex numbers behind the pass
};
