// kernels_fused.cu -- the fused tile pass: one HBM read + one HBM write of the
// whole shard while a run of gates is applied to amplitudes held in registers.
//
// Stands in for the reference's one-sweep-per-gate loops
// (q_apply_1q_gate, reference src/q_gates.c:131-143; q_apply_2q_gate,
// :276-292), which move 64 bytes per amplitude PER GATE.  Here a CTA owns a
// tile of 2^12 amplitudes (64 KiB) and every thread keeps 2^R of them in
// registers (R = 4 with 256 threads, R = 3 with 512).  The 12 tile-index bits
// are split into 5 lane bits, warp bits and R register bits; a gate pairs
// amplitudes across one register bit, so it is pure register arithmetic.  When
// the run moves on to qubits that are not register bits the tile is transposed
// through shared memory (one 16-byte store + one 16-byte load per amplitude,
// XOR-swizzled so quarter-warps hit 8 distinct bank groups) and the roles are
// re-assigned (planner.cpp).  Diagonal gates (Z, P, RZ, CPHASE) and controls
// need no pairing: they read the amplitude's own index bits, whatever role
// those bits currently have.
//
// The amplitudes live in PTX registers DECLARED BY NAME (qr0.., qi0..) and
// every operation on them is inline PTX.  The gate list is interpreted at run
// time -- one table-driven branch per gate into straight-line FP64 code
// (fused_lists.inc, generated).  When the amplitudes were C++ variables the
// compiler merged them at every branch join with ~64 register moves per gate,
// which left the kernel issue-bound at 25 % of the HBM rate
// (profiles/r1_fused_kernel_history.md); named registers have no SSA joins.
// Every arithmetic instruction carries an explicit .rn, so nothing is
// contracted into an FMA (bit-exactness, see gate_math.cuh).
//
// Variants (launch_fused_pass):
//   0 "ldg"    256 threads x 16 amplitudes, one CTA per tile, plain 128-bit
//              global loads/stores, 2 CTAs per SM
//   1 "tma16"  persistent CTA per SM: 8 compute warps (16 amplitudes each) +
//              1 copy warp staging tiles through shared memory with TMA bulk copies
//   2 "tma"    same, 16 compute warps x 8 amplitudes: twice the warps hide the
//              per-gate dispatch latency
//   3 "ldg8"   8 amplitudes per thread, plain loads; 12-, 11- or 10-bit tiles = 2 x 512, 4 x 256 or
//              8 x 128 threads per SM (32 warps per SM); the default
// math=fast (opt-in): the same ldg8 / ldg kernel bodies around the QCS<R>F
// interpreter (fused multiply-adds, fan product tables, per-thread pending factor, uniform-fan
// prologue): fused_pass_ldg_r3f{,_t11,_t10}, fused_pass_ldg_r4f{,_t11}.
//
// Roofline: HBM.  Algorithmic bytes per launch = 32 * 2^nl (every amplitude
// read once, written once); the planner caps the fused FP64 work per pass so
// the pass stays bandwidth-bound (DESIGN.md "Kernels").
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.h"
#include "fused_lists.inc"
#include "kernels.h"

namespace qcs {

namespace {

__device__ __forceinline__ uint32_t swizzle(uint32_t slot) {
  return slot ^ (((slot >> 3) ^ (slot >> 6) ^ (slot >> 9)) & 7u);
}

// Software bit-deposit: spreads the low bits of v over the set bits of mask.
__device__ __forceinline__ uint64_t deposit_bits(uint64_t v, uint64_t mask) {
  uint64_t r = 0;
  while (mask) {
    uint64_t low = mask & (~mask + 1);
    if (v & 1) r |= low;
    v >>= 1;
    mask ^= low;
  }
  return r;
}

// ---------------------------------------------------------------------------
// Inline-PTX operations on the named register file (qr<k>, qi<k>).
// ---------------------------------------------------------------------------
// (load/store/transpose operations are defined per instantiation in fused_body.inc)

// ---- pairing updates; a = register with target bit 0, b = target bit 1 --------
// Operation order = c_add(c_mul(g0,v0), c_mul(g1,v1)) (reference src/q_gates.c:140-141,
// src/complex.c:23-57); *_R0 = reference-semantics controlled update (row 0 only).

// GK_PAIR_HSYM: real, U00==U10=m0, U01=m2, U11=-m2: the four products are shared.
#define OP_HSYM(a, b)                                                               \
  asm volatile("{\n\t.reg .f64 t0,t1,t2,t3;\n\t"                                    \
               "mul.rn.f64 t0, %0, qr" #a ";\n\tmul.rn.f64 t1, %0, qi" #a ";\n\t"   \
               "mul.rn.f64 t2, %1, qr" #b ";\n\tmul.rn.f64 t3, %1, qi" #b ";\n\t"   \
               "add.rn.f64 qr" #a ", t0, t2;\n\tadd.rn.f64 qi" #a ", t1, t3;\n\t"   \
               "sub.rn.f64 qr" #b ", t0, t2;\n\tsub.rn.f64 qi" #b ", t1, t3;\n\t}"  \
               ::"d"(m0), "d"(m2));
#define OP_HSYM_R0(a, b)                                                            \
  asm volatile("{\n\t.reg .f64 t0,t1,t2,t3;\n\t"                                    \
               "mul.rn.f64 t0, %0, qr" #a ";\n\tmul.rn.f64 t1, %0, qi" #a ";\n\t"   \
               "mul.rn.f64 t2, %1, qr" #b ";\n\tmul.rn.f64 t3, %1, qi" #b ";\n\t"   \
               "add.rn.f64 qr" #a ", t0, t2;\n\tadd.rn.f64 qi" #a ", t1, t3;\n\t}"  \
               ::"d"(m0), "d"(m2));

// GK_PAIR_REAL: all imaginary parts zero.
#define OP_REAL(a, b)                                                               \
  asm volatile("{\n\t.reg .f64 t0,t1,t2,t3,t4,t5,t6,t7;\n\t"                        \
               "mul.rn.f64 t0, %0, qr" #a ";\n\tmul.rn.f64 t1, %0, qi" #a ";\n\t"   \
               "mul.rn.f64 t2, %1, qr" #b ";\n\tmul.rn.f64 t3, %1, qi" #b ";\n\t"   \
               "mul.rn.f64 t4, %2, qr" #a ";\n\tmul.rn.f64 t5, %2, qi" #a ";\n\t"   \
               "mul.rn.f64 t6, %3, qr" #b ";\n\tmul.rn.f64 t7, %3, qi" #b ";\n\t"   \
               "add.rn.f64 qr" #a ", t0, t2;\n\tadd.rn.f64 qi" #a ", t1, t3;\n\t"   \
               "add.rn.f64 qr" #b ", t4, t6;\n\tadd.rn.f64 qi" #b ", t5, t7;\n\t}"  \
               ::"d"(m0), "d"(m2), "d"(m4), "d"(m6));
#define OP_REAL_R0(a, b)                                                            \
  asm volatile("{\n\t.reg .f64 t0,t1,t2,t3;\n\t"                                    \
               "mul.rn.f64 t0, %0, qr" #a ";\n\tmul.rn.f64 t1, %0, qi" #a ";\n\t"   \
               "mul.rn.f64 t2, %1, qr" #b ";\n\tmul.rn.f64 t3, %1, qi" #b ";\n\t"   \
               "add.rn.f64 qr" #a ", t0, t2;\n\tadd.rn.f64 qi" #a ", t1, t3;\n\t}"  \
               ::"d"(m0), "d"(m2));

// GK_PAIR_SWAP: exact X.
#define OP_SWAP(a, b)                                                               \
  asm volatile("{\n\t.reg .f64 t0,t1;\n\t"                                          \
               "mov.f64 t0, qr" #a ";\n\tmov.f64 t1, qi" #a ";\n\t"                 \
               "mov.f64 qr" #a ", qr" #b ";\n\tmov.f64 qi" #a ", qi" #b ";\n\t"     \
               "mov.f64 qr" #b ", t0;\n\tmov.f64 qi" #b ", t1;\n\t}" ::);
#define OP_SWAP_R0(a, b)                                                            \
  asm volatile("mov.f64 qr" #a ", qr" #b ";\n\tmov.f64 qi" #a ", qi" #b ";" ::);

// GK_PAIR_GENERIC: the full 28-flop update.
//   x = g*v : x.r = g.r*v.r - g.i*v.i ; x.i = g.r*v.i + g.i*v.r
#define QCS_CMUL(xr, xi, gr, gi, v)                                                 \
  "mul.rn.f64 " xr ", " gr ", qr" v ";\n\tmul.rn.f64 u0, " gi ", qi" v ";\n\t"      \
  "sub.rn.f64 " xr ", " xr ", u0;\n\t"                                              \
  "mul.rn.f64 " xi ", " gr ", qi" v ";\n\tmul.rn.f64 u0, " gi ", qr" v ";\n\t"      \
  "add.rn.f64 " xi ", " xi ", u0;\n\t"
#define OP_GENERIC(a, b)                                                            \
  asm volatile("{\n\t.reg .f64 t0,t1,t2,t3,t4,t5,t6,t7,u0;\n\t"                     \
               QCS_CMUL("t0", "t1", "%0", "%1", #a)                                 \
               QCS_CMUL("t2", "t3", "%2", "%3", #b)                                 \
               QCS_CMUL("t4", "t5", "%4", "%5", #a)                                 \
               QCS_CMUL("t6", "t7", "%6", "%7", #b)                                 \
               "add.rn.f64 qr" #a ", t0, t2;\n\tadd.rn.f64 qi" #a ", t1, t3;\n\t"   \
               "add.rn.f64 qr" #b ", t4, t6;\n\tadd.rn.f64 qi" #b ", t5, t7;\n\t}"  \
               ::"d"(m0), "d"(m1), "d"(m2), "d"(m3), "d"(m4), "d"(m5), "d"(m6), "d"(m7));
#define OP_GENERIC_R0(a, b)                                                         \
  asm volatile("{\n\t.reg .f64 t0,t1,t2,t3,u0;\n\t"                                 \
               QCS_CMUL("t0", "t1", "%0", "%1", #a)                                 \
               QCS_CMUL("t2", "t3", "%2", "%3", #b)                                 \
               "add.rn.f64 qr" #a ", t0, t2;\n\tadd.rn.f64 qi" #a ", t1, t3;\n\t}"  \
               ::"d"(m0), "d"(m1), "d"(m2), "d"(m3));

// GK_DIAG: amplitude r *= (dr, di), c_mul order.  DIAG uses the per-thread
// selected entry (dr, di); DIAG0 / DIAG1 use matrix entry 0 (m0, m1) / 1 (m6, m7).
#define QCS_DIAG_ASM(r, er, ei)                                                     \
  asm volatile("{\n\t.reg .f64 t0,t1,t2,t3;\n\t"                                    \
               "mul.rn.f64 t0, %0, qr" #r ";\n\tmul.rn.f64 t1, %1, qi" #r ";\n\t"   \
               "mul.rn.f64 t2, %0, qi" #r ";\n\tmul.rn.f64 t3, %1, qr" #r ";\n\t"   \
               "sub.rn.f64 qr" #r ", t0, t1;\n\tadd.rn.f64 qi" #r ", t2, t3;\n\t}"  \
               ::"d"(er), "d"(ei));
#define OP_DIAG(r) QCS_DIAG_ASM(r, dr, di)
#define OP_DIAG0(r) QCS_DIAG_ASM(r, m0, m1)
#define OP_DIAG1(r) QCS_DIAG_ASM(r, m6, m7)

// per-case matrix loads (constant bank): only what the class needs
#define QCS_LD_HSYM const double m0 = g.m[0], m2 = g.m[2];
#define QCS_LD_HSYM_R0 QCS_LD_HSYM
#define QCS_LD_REAL const double m0 = g.m[0], m2 = g.m[2], m4 = g.m[4], m6 = g.m[6];
#define QCS_LD_REAL_R0 QCS_LD_HSYM
#define QCS_LD_SWAP
#define QCS_LD_SWAP_R0
#define QCS_LD_GENERIC_R0 const double m0 = g.m[0], m1 = g.m[1], m2 = g.m[2], m3 = g.m[3];
#define QCS_LD_GENERIC QCS_LD_GENERIC_R0 const double m4 = g.m[4], m5 = g.m[5], m6 = g.m[6], m7 = g.m[7];
#define QCS_LD_DIAG0 const double m0 = g.m[0], m1 = g.m[1];
#define QCS_LD_DIAG1 const double m6 = g.m[6], m7 = g.m[7];

#define COMPUTE_BAR(nthreads) asm volatile("bar.sync 1, %0;" ::"n"(nthreads) : "memory")

// ---- TMA / mbarrier primitives ---------------------------------------------------
constexpr int kSlots = 3;
constexpr uint32_t kTileBytes = 16u << QCS_TILE_BITS;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "QCS_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra QCS_DONE;\n\t"
      "bra QCS_WAIT;\n"
      "QCS_DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes,
                                          uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(src_smem), "r"(bytes)
               : "memory");
}

// ---------------------------------------------------------------------------
// Kernel bodies, instantiated for R = 4 (16 amplitudes/thread) and R = 3 (8).
// ---------------------------------------------------------------------------
#define QCS_SWAP 0
#define QCS_FAST 0
#define QCS_R 4
#define QCS_CT (1 << (QCS_T - QCS_R))
#define QCS_NREG_STR "16"
#define QCS_LIST(x) QCS4_##x
#define QCS_WITH_LDG 1

#define QCS_T 12
#define QCS_MIN_CTAS 2
#define QCS_NAME(x) x##_r4
#define QCS_WITH_TMA 1
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME
#undef QCS_WITH_TMA

// 16 amplitudes per thread on small tiles: 4 x 128 or 8 x 64 threads per SM.  Half the warps of ldg8,
// but every per-gate cost (dispatch, fan-entry bookkeeping) is spread over twice the amplitudes and a
// segment pairs on four positions.
#define QCS_WITH_TMA 0
#define QCS_T 11
#define QCS_MIN_CTAS 3
#define QCS_NAME(x) x##_r4_t11
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME

#define QCS_T 10
#define QCS_MIN_CTAS 8
#define QCS_NAME(x) x##_r4_t10
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME
#undef QCS_WITH_TMA

#undef QCS_R
#undef QCS_CT
#undef QCS_NREG_STR
#undef QCS_LIST
#undef QCS_WITH_LDG

#define QCS_R 3
#define QCS_NREG_STR "8"
#define QCS_LIST(x) QCS3_##x
#define QCS_WITH_LDG 1
#define QCS_CT (1 << (QCS_T - QCS_R))

#define QCS_T 12
#define QCS_MIN_CTAS 2
#define QCS_NAME(x) x##_r3
#define QCS_WITH_TMA 1
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME
#undef QCS_WITH_TMA

// the same three shapes once more with the remap machinery compiled in (QCS_SWAP): x_r3s...
#define QCS_WITH_TMA 0
#undef QCS_SWAP
#define QCS_SWAP 1
#define QCS_T 12
#define QCS_MIN_CTAS 2
#define QCS_NAME(x) x##_r3s
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME
#define QCS_T 11
#define QCS_MIN_CTAS 4
#define QCS_NAME(x) x##_r3s_t11
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME
#define QCS_T 10
#define QCS_MIN_CTAS 8
#define QCS_NAME(x) x##_r3s_t10
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME
#undef QCS_SWAP
#define QCS_SWAP 0
#undef QCS_WITH_TMA

#define QCS_T 11
#define QCS_MIN_CTAS 4
#define QCS_NAME(x) x##_r3_t11
#define QCS_WITH_TMA 0
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME

#define QCS_T 10
#define QCS_MIN_CTAS 8
#define QCS_NAME(x) x##_r3_t10
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME
#undef QCS_WITH_TMA

// math=fast: the same three ldg8 kernels around the fused-multiply-add interpreter (QCS3F_*)
#undef QCS_LIST
#undef QCS_FAST
#define QCS_LIST(x) QCS3F_##x
#define QCS_FAST 1
#define QCS_WITH_TMA 0

#define QCS_T 12
#define QCS_MIN_CTAS 2
#define QCS_NAME(x) x##_r3f
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME

#define QCS_T 11
#define QCS_MIN_CTAS 4
#define QCS_NAME(x) x##_r3f_t11
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME

#define QCS_T 10
#define QCS_MIN_CTAS 8
#define QCS_NAME(x) x##_r3f_t10
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME

#undef QCS_SWAP
#define QCS_SWAP 1
#define QCS_T 12
#define QCS_MIN_CTAS 2
#define QCS_NAME(x) x##_r3fs
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME
#define QCS_T 11
#define QCS_MIN_CTAS 4
#define QCS_NAME(x) x##_r3fs_t11
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME
#define QCS_T 10
#define QCS_MIN_CTAS 8
#define QCS_NAME(x) x##_r3fs_t10
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME
#undef QCS_SWAP
#define QCS_SWAP 0
#undef QCS_WITH_TMA

#undef QCS_R
#undef QCS_CT
#undef QCS_NREG_STR
#undef QCS_LIST

// math=fast with 16 amplitudes per thread (tile_kernel=ldg): what a thread pays per GATE (dispatch,
// fan walks, table lookups) is spread over twice the amplitudes, and a segment pairs on four
// positions instead of three.  The bit-exact mode is FP64-bound and gains nothing from this shape;
// the fast mode is bound by exactly those per-gate costs.
#define QCS_R 4
#define QCS_NREG_STR "16"
#define QCS_LIST(x) QCS4F_##x
#define QCS_CT (1 << (QCS_T - QCS_R))
#define QCS_WITH_TMA 0

#define QCS_T 12
#define QCS_MIN_CTAS 2
#define QCS_NAME(x) x##_r4f
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME

#define QCS_T 11
#define QCS_MIN_CTAS 4
#define QCS_NAME(x) x##_r4f_t11
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME

#define QCS_T 10
#define QCS_MIN_CTAS 8
#define QCS_NAME(x) x##_r4f_t10
#include "fused_body.inc"
#undef QCS_T
#undef QCS_MIN_CTAS
#undef QCS_NAME
#undef QCS_WITH_TMA

#undef QCS_R
#undef QCS_CT
#undef QCS_NREG_STR
#undef QCS_LIST
#undef QCS_WITH_LDG
#undef QCS_FAST

static int tile_row_bits(const PassParams &p) {
  int b = 0;
  while (b < QCS_TILE_BITS && p.tile_pos[b] == b) b++;
  return b;  // >= 5: positions 0..4 are always tile bits
}

}  // namespace

using LdgKernel = void (*)(double2 *, const PassParams, uint32_t, const SwapStore, const PassExtras, uint32_t);

// [math=fast][16 amplitudes per thread][tile bits - 10]
// the ldg8 shapes with the remap machinery compiled in: [math=fast][tile bits - 10]
static LdgKernel ldg8_remap_kernel(bool fast, int T) {
  static const LdgKernel table[2][3] = {
      {fused_pass_ldg_r3s_t10, fused_pass_ldg_r3s_t11, fused_pass_ldg_r3s},
      {fused_pass_ldg_r3fs_t10, fused_pass_ldg_r3fs_t11, fused_pass_ldg_r3fs}};
  return table[fast ? 1 : 0][T - QCS_MIN_TILE_BITS];
}

static LdgKernel ldg_kernel(bool fast, bool r4, int T) {
  static const LdgKernel table[2][2][3] = {
      {{fused_pass_ldg_r3_t10, fused_pass_ldg_r3_t11, fused_pass_ldg_r3},
       {fused_pass_ldg_r4_t10, fused_pass_ldg_r4_t11, fused_pass_ldg_r4}},
      {{fused_pass_ldg_r3f_t10, fused_pass_ldg_r3f_t11, fused_pass_ldg_r3f},
       {fused_pass_ldg_r4f_t10, fused_pass_ldg_r4f_t11, fused_pass_ldg_r4f}}};
  return table[fast ? 1 : 0][r4 ? 1 : 0][T - QCS_MIN_TILE_BITS];
}

// Per-device launch state (engines on different GPUs may live in one process).
struct DeviceLaunchState {
  int sm_count = 0;
  bool ldg_configured[2][2][3] = {};
  bool remap_configured[2][3] = {};
  bool tma_configured[2] = {false, false};
};
static DeviceLaunchState *device_launch_state(cudaError_t *err) {
  static DeviceLaunchState states[64];
  int dev = 0;
  *err = cudaGetDevice(&dev);
  if (*err != cudaSuccess) return nullptr;
  if (dev < 0 || dev >= 64) {
    *err = cudaErrorInvalidDevice;
    return nullptr;
  }
  DeviceLaunchState *s = &states[dev];
  if (s->sm_count == 0) {
    int sms = 0;
    *err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (*err != cudaSuccess) return nullptr;
    s->sm_count = sms;
  }
  return s;
}

int device_sm_count() {
  cudaError_t err;
  DeviceLaunchState *s = device_launch_state(&err);
  return s ? s->sm_count : 148;
}

cudaError_t launch_fused_pass(double2 *state, const PassParams &params, int n_local,
                              cudaStream_t stream, int variant, const SwapStore *swap, bool fast,
                              uint32_t pass_flags, const PassExtras *extras) {
  if (variant < 0 || variant > 3) return cudaErrorInvalidValue;
  const bool ldg = variant == 0 || variant == 3;
  if ((swap || fast || pass_flags) && !ldg) return cudaErrorInvalidValue;  // plain-load kernels only
  if (swap && swap->k && variant != 3) return cudaErrorInvalidValue;        // remaps ride on ldg8 passes only
  SwapStore sw{};
  if (swap) sw = *swap;
  PassExtras ex{};
  if (extras) ex = *extras;
  if ((pass_flags & QCS_PASS_ARGMAX) && (!ex.argmax_p || !ex.argmax_idx)) return cudaErrorInvalidValue;
  const int T = params.tile_bits;
  if (T < QCS_MIN_TILE_BITS || T > QCS_TILE_BITS || T > n_local) return cudaErrorInvalidValue;
  if (T != QCS_TILE_BITS && !ldg) return cudaErrorInvalidValue;  // the TMA kernels exist for 12-bit tiles only
  if ((variant >= 2) != (params.reg_bits == 3)) return cudaErrorInvalidValue;
  const unsigned n_tiles = 1u << (n_local - T);
  cudaError_t e;
  DeviceLaunchState *dls = device_launch_state(&e);
  if (!dls) return e;
  const int sm_count = dls->sm_count;
  if (ldg) {
    const bool r4 = variant == 0;
    // math=fast: the tile + one 16-byte factor per uniform fan behind it
    const size_t smem = ((size_t)16 << T) + (fast ? 16 * QCS_MAX_PASS_FANS : 0);
    const bool remap = swap && sw.k;
    LdgKernel k = remap ? ldg8_remap_kernel(fast, T) : ldg_kernel(fast, r4, T);
    bool &configured = remap ? dls->remap_configured[fast ? 1 : 0][T - QCS_MIN_TILE_BITS]
                             : dls->ldg_configured[fast ? 1 : 0][r4 ? 1 : 0][T - QCS_MIN_TILE_BITS];
    if (!configured) {
      e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured = true;
    }
    // One CTA per tile.  (A remap-carrying pass was also tried PERSISTENT -- as many CTAs as are resident,
    // each looping over its tiles with the bulk stores of tile i draining while it loads and computes
    // tile i + 1, QCS_CUDA_PERSISTENT_REMAP=1 -- and was 15-40 % slower: CTAs that all start together
    // and loop stay in lock step, so loads, FP64 work and stores no longer overlap ACROSS CTAs, which
    // is worth more than the overlap gained inside one; profiles/r2g_swap_bw_persistent.log.)
    unsigned grid = n_tiles;
    static const bool persistent_remap = getenv("QCS_CUDA_PERSISTENT_REMAP") && atoi(getenv("QCS_CUDA_PERSISTENT_REMAP")) != 0;
    if (swap && sw.k && persistent_remap) {
      const unsigned resident = (unsigned)sm_count * (r4 ? (T == 12 ? 2u : T == 11 ? 3u : 8u) : (8u >> (T - QCS_MIN_TILE_BITS)));
      if (grid > resident) grid = resident;
    }
    k<<<grid, 1u << (T - params.reg_bits), smem, stream>>>(state, params, pass_flags, sw, ex, n_tiles);
    return cudaGetLastError();
  }
  const size_t smem_tma = (size_t)kSlots * kTileBytes + 128;  // slots + barriers + tile origins
  bool &configured = dls->tma_configured[variant - 1];
  if (!configured) {
    e = cudaFuncSetAttribute(variant == 1 ? fused_pass_tma_r4 : fused_pass_tma_r3,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const unsigned grid = n_tiles < (unsigned)sm_count ? n_tiles : (unsigned)sm_count;
  if (variant == 1)
    fused_pass_tma_r4<<<grid, 256 + 32, smem_tma, stream>>>(state, params, n_tiles, (uint32_t)tile_row_bits(params));
  else
    fused_pass_tma_r3<<<grid, 512 + 32, smem_tma, stream>>>(state, params, n_tiles, (uint32_t)tile_row_bits(params));
  return cudaGetLastError();
}

}  // namespace qcs
