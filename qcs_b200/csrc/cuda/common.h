// common.h -- types shared by the host-side planner and the device kernels.
//
// Vocabulary:
//   amplitude   one complex double, 16 bytes, interleaved {re, im}
//   shard       the 2^nl amplitudes this rank holds (nl = local qubits)
//   position    a bit of the physical basis index; positions < nl are local,
//               positions >= nl are the rank's bits (global)
//   tile        2^QCS_TILE_BITS amplitudes one CTA holds in registers during a
//               pass; its index bits are `tile_pos` (always contains 0..4)
//   pass        one fused kernel launch: every amplitude is read once and
//               written once, a run of gates is applied in between
//   segment     a stretch of a pass during which the assignment of tile bits
//               to {lane, warp, register} roles is fixed; a gate can only pair
//               amplitudes across a *register* bit, so changing the pairing
//               qubit set means a shared-memory transposition between segments
#pragma once
#include <stdint.h>

#define QCS_TILE_BITS 12     // 4096 amplitudes = 64 KiB per tile
#define QCS_LANE_BITS 5
#define QCS_MAX_REG_BITS 4   // 16 amplitudes per thread (256-thread CTA) or 8 (512-thread CTA)
#define QCS_MAX_PASS_GATES 120
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
  GF_D1_IDENT = 4     // diagonal entry 1 is exactly (1,0): untouched
};

// Dispatch ids: the `op` byte of a DGate selects one straight-line block of the
// generated switch (gen_fused_lists.py emits the cases with the same formulas).
//   kind_index: 0 HSYM, 1 REAL, 2 SWAP, 3 GENERIC;  treg/creg: register-role index, creg -1 = none
static inline int qcs_op_id_pair(int kind_index, int row0, int treg, int creg) {
  return ((kind_index * 2 + row0) * 4 + treg) * 5 + (creg + 1);
}
// diagonal whose target is NOT a register bit: the thread picks entry 0 or 1 from bit `tsel`
static inline int qcs_op_id_diag_free(int creg) { return 160 + creg + 1; }
// diagonal whose target is register bit treg; halves: 1 = entry 0 only, 2 = entry 1 only, 3 = both
static inline int qcs_op_id_diag_reg(int treg, int creg, int halves) {
  return 165 + (treg * 5 + creg + 1) * 3 + (halves - 1);
}
#define QCS_OP_DIAG_FREE_FIRST 160
#define QCS_OP_DIAG_FREE_LAST 164
#define QCS_OP_NONE 255
// Header of a controlled-phase fan: the next `tsel` gate records are diag(1, e^{ia}) gates
// controlled by various qubits on ONE target; `ctest` = selector of the target bit when it is not
// a register bit (0xFF otherwise), treg in treg_creg.  The interpreter runs them in a tight loop.
#define QCS_OP_FAN 250
#define QCS_MAX_PASS_FANS 24
// DGate.ctest / DGate.tsel: 0..11 = tile-bit index (per-thread test), QCS_SEL_OUTSIDE | pos =
// physical position outside the tile (same value for the whole tile), 0xFF = none
#define QCS_SEL_OUTSIDE 0x40

// A gate as the device sees it (inside a pass descriptor).  The kernel reads
// m[] and the first four header bytes; the rest documents the plan.
struct DGate {
  double m[8];      // row-major {re,im} x 4
  uint8_t op;       // dispatch id (above)
  uint8_t flags;    // GateFlags
  uint8_t ctest;    // control that is not a register bit: selector (QCS_SEL_*) of the bit that must
                    // be 1 for this thread to take part; 0xFF = no test
  uint8_t tsel;     // for diag_free ops: selector of the target bit, else 0xFF
  uint8_t kind;     // GateKind
  int8_t tpos;      // physical position of the target bit
  int8_t cpos;      // physical position of the control bit, -1 if none
  uint8_t treg_creg; // (treg + 1) | (creg + 1) << 4
};

struct DSegment {
  // role r -> tile bit index; roles 0-4 lane bits, then warp bits, the last reg_bits are register bits
  uint8_t role_tilebit[QCS_TILE_BITS];
  uint16_t gate_begin, gate_end;
  // Host-precomputed per-thread arithmetic (the kernel would otherwise redo it for every tile):
  uint16_t tb_lut[3][8];               // tile-index bits contributed by thread-id bits [3g, 3g+3)
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
  uint64_t nontile_mask;               // local positions that are not tile bits
  int32_t n_segments;
  int32_t n_gates;
  int32_t reg_bits;                    // 4: 256 threads x 16 amplitudes, 3: 512 threads x 8
  int32_t n_tile_runs;
  DTileRun tile_run[12];               // tile number -> element offset of the tile (bit deposit as runs)
  // element offsets for the global loads (first segment) and stores (last segment)
  uint64_t gb_lut[2][3][8];            // [first/last][thread-id bit group][value]
  uint64_t goff[2][QCS_MAX_REG_BITS];  // [first/last][register role bit]
  DSegment seg[QCS_MAX_PASS_SEGMENTS];
  DGate gate[QCS_MAX_PASS_GATES + QCS_MAX_PASS_FANS + 1];  // fan headers; +1: one-ahead prefetch sentinel
};
