// kernels.h -- launchers implemented in the .cu files (all asynchronous on `stream`).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"

namespace qcs {

// kernels_fused.cu ----------------------------------------------------------
// variant 0: one CTA per tile, plain 128-bit global loads/stores
// variant 1: persistent CTAs, tiles staged through shared memory by TMA bulk copies
// swap (plain-load kernels only, multi-GPU): the pass also trades up to QCS_MAX_REMAP local positions
// for global ones on its way out -- an all-to-all among the 2^k ranks that differ in the traded rank
// bits.  An amplitude whose bits at the traded local positions are b = (b_0 .. b_k-1) lands in the
// shard of the rank whose traded rank bits are b (the other rank bits are mine), at its own index
// with those local bits replaced by MY traded rank bits; (1 - 2^-k) of the shard crosses NVLink,
// against k halves for k separate swaps.
#define QCS_MAX_REMAP 3
struct SwapStore {
  uint32_t k;                        // position pairs traded (0 = plain pass)
  uint32_t lpos[QCS_MAX_REMAP];      // the local positions
  uint32_t my_gbits;                 // bit i = my rank's bit at the global position traded for lpos[i]
  uint32_t in_tile;                  // bit i: lpos[i] is one of the pass's tile positions
  uint32_t n_out;                    // pairs whose local position is NOT a tile position ...
  uint32_t out_pair[QCS_MAX_REMAP];  // ... which pairs, by ascending bit of the tile number ...
  uint32_t out_tile_bit[QCS_MAX_REMAP];  // ... and which bit of the tile NUMBER each of them is: the low n_out
                                     // bits of the block index supply these, so tiles bound for different
                                     // ranks (and tiles that stay) alternate in launch order and the transfer
                                     // runs underneath the whole pass
  double2 *peer[1 << QCS_MAX_REMAP];       // [b]: shard of the rank whose traded rank bits are b; [my_gbits] = mine
  uint32_t *peer_flags[1 << QCS_MAX_REMAP];  // [b]: where that rank expects MY "tile loaded" signals
  const uint32_t *my_flags;          // my flag words: [s * flag_stride + block] is written by the rank with traded
                                     // bits s once its CTA `block` holds, in registers, the tile my stores overwrite
  uint32_t flag_stride;
  uint32_t epoch;                    // value of this pass (strictly increasing per engine)
  uint32_t bulk;                     // 1: tiles with data for other ranks leave through shared memory and TMA bulk
                                     // stores (cp.async.bulk.global.shared::cta, one per contiguous row);
                                     // 0: 16-byte st.global from the compute threads (same speed, measured)
  uint32_t row_bits;                 // low tile positions 0..row_bits-1 are contiguous: a row = 16 << row_bits bytes
  uint32_t *abort_flag;              // my own device word: set when a CTA gave up waiting for a partner (the tile
                                     // is then NOT stored and the engine reports QCS_CUDA_ERR_CUDA); later CTAs bail out
  unsigned long long spin_limit;     // clock64 ticks a CTA waits for a partner's signal before giving up
  uint32_t out_of_place;             // 1: peer[] are the ranks' SECOND shard buffers (Engine::alt) -- every tile of
                                     // the pass, the ones that stay included, is stored there and the buffers
                                     // trade roles behind the pass.  Nothing a partner still has to read is
                                     // overwritten, so there are no flags, no signal and no wait: the ranks run
                                     // the pass at their own pace.  0: in place, with the per-tile handshake.
  double2 *own_out;                  // out_of_place: peer[my_gbits], my own second buffer (no indexed parameter read)
};
// fast (ldg8 only): the fused-multiply-add interpreter; `params` must come from a planner run with
// PlannerConfig::fast_math (fan entries carry product tables instead of single phases).
// pass_flags (plain-load kernels only): QCS_PASS_SYNTH_ZERO_KET = do not read the shard, it is the
// never-written |0...0> of qc_create (amplitude 0 of the whole register is 1, everything else 0).
//   QCS_PASS_ARGMAX = every CTA also leaves (max |a|^2, lowest index holding it) of the tile it wrote in
//   extras->argmax_p / argmax_idx [block index]: qc_find_most_likely_state right behind a run of gates
//   then reduces 2^(nl-T) pairs (launch_argmax_pairs) instead of sweeping the shard once more.
enum { QCS_PASS_SYNTH_ZERO_KET = 1, QCS_PASS_ARGMAX = 2 };
struct PassExtras {
  double *argmax_p;
  long long *argmax_idx;
  const double2 *thread_tables;  // math=fast: the pass's thread-table fans (common.h QCS_OP_TFAN_BASE), device memory
};
cudaError_t launch_fused_pass(double2 *state, const PassParams &params, int n_local,
                              cudaStream_t stream, int variant, const SwapStore *swap = nullptr,
                              bool fast = false, uint32_t pass_flags = 0, const PassExtras *extras = nullptr);

// multiprocessor count of the current device (cached per device); grids are sized from it
int device_sm_count();

// kernels_simple.cu: one launch per gate (fusion off, shards below one tile) ---
cudaError_t launch_simple_gate(double2 *state, const DGate &g, int n_local,
                               uint64_t shard_base, cudaStream_t stream);

cudaError_t launch_swap_local_bits(double2 *state, int n_local, int a, int b, cudaStream_t stream);
// `out` must hold blocks * 256 doubles; every thread performs 16 * iters separately rounded FP64 ops
cudaError_t launch_fp64_probe(double *out, int blocks, int iters, cudaStream_t stream);

// kernels_reduce.cu -----------------------------------------------------------
struct ReduceWorkspace {
  double *partials;        // device, >= 4 * REDUCE_MAX_BLOCKS doubles
  long long *ipartials;    // device, >= REDUCE_MAX_BLOCKS
  double *result;          // device, RES_COUNT doubles
  long long *iresult;      // device, 2 long longs
  // exact sequential-prefix machinery (chunked)
  double *chunk_sum;       // device, n_chunks
  double *chunk_sum2;      // device, n_chunks: the imaginary parts' chunk sums (diffusion)
  double *chunk_abs;       // device, n_chunks: sums of |re| per chunk (signed sums: "adds nothing" test)
  double *chunk_abs2;      // device, n_chunks: sums of |im| per chunk
  double *chunk_approx;    // device, n_chunks + 1
  double *chunk_delta;     // device, n_chunks
  unsigned char *chunk_flag;  // device, n_chunks
  double *chunk_exact;     // device, n_chunks + 1
  double *group_sum;       // device, n_chunks / QCS_RESOLVE_GROUP + 2 (two-level resolve of large shards)
  double *group_exact;     // device, n_chunks / QCS_RESOLVE_GROUP + 2
  unsigned char *group_flag;  // device, n_chunks / QCS_RESOLVE_GROUP + 2
  size_t n_chunks_cap;
};
enum { REDUCE_MAX_BLOCKS = 4096, SEQ_CHUNK = 1024 };
// chunks per group of the two-level resolve.  Small groups: a group is walked chunk by chunk as soon
// as ONE of its chunks is flagged or it spans a binade, and the one warp that walks pays ~1 us per 32
// chunks -- with 1024-chunk groups the five binade crossings a 24-qubit Grover state has beyond its
// first group cost 5 x 32 us per sum; the clean groups themselves cost ~20 cycles each (their records
// are fetched 32 groups at a time).
#define QCS_RESOLVE_GROUP 64
// What the exact sequential sums add up (the `mask_pos` argument below): a position >= 0 = |a_i|^2 over
// the indices with that bit clear, SEL_ALL = every |a_i|^2, SEL_RE / SEL_IM = the real / imaginary parts
// of the amplitudes themselves (signed; q_apply_diffusion, reference src/q_gates.c:334-336).
enum { SEL_ALL = -1, SEL_RE = -2, SEL_IM = -3 };

cudaError_t launch_init_state(double2 *state, uint64_t n_amps, bool set_one, cudaStream_t s);
cudaError_t launch_complex_sum(const double2 *state, uint64_t n_amps, ReduceWorkspace &ws,
                               cudaStream_t s);  // result[RES_LOCAL_SUM_RE..IM] = this shard's sum
// result[0..1] (raw sum of the whole state) -> result[2..3] = 2*mean, per semantics (device-side, no host sync)
cudaError_t launch_diffusion_mean(ReduceWorkspace &ws, bool corrected, double n_total,
                                  cudaStream_t s);
// dst[i] = 2*mean - src[i]   (dst may alias src)
cudaError_t launch_diffusion_write(const double2 *src, double2 *dst, uint64_t n_amps,
                                   const ReduceWorkspace &ws, cudaStream_t s);
// the same and result[RES_LOCAL_SUM_RE..IM] = sum of the values written (carried to the next iteration)
cudaError_t launch_diffusion_write_sum(const double2 *src, double2 *dst, uint64_t n_amps,
                                       ReduceWorkspace &ws, cudaStream_t s);
cudaError_t launch_scale(double2 *state, uint64_t n_amps, double factor, cudaStream_t s);
cudaError_t launch_zero_half(double2 *state, uint64_t n_amps, int pos, int keep_bit,
                             cudaStream_t s);
// dst[index] = -src[index]; carried_sum (device, may be null) -= 2 * src[index]
cudaError_t launch_negate_one(double2 *dst, const double2 *src, uint64_t index,
                              double *carried_sum, cudaStream_t s);
cudaError_t launch_argmax(const double2 *state, uint64_t n_amps, ReduceWorkspace &ws,
                          cudaStream_t s);  // result[0] = max prob, iresult[0] = first index
// the same from n (probability, index) candidates (QCS_PASS_ARGMAX); ties go to the lowest index
cudaError_t launch_argmax_pairs(const double *p, const long long *idx, uint64_t n, ReduceWorkspace &ws,
                                cudaStream_t s);
// Local (physical) index -> logical basis index under the current qubit layout, one table per index
// byte (shards hold at most 2^32 amplitudes); `base` = the contribution of this rank's bits.
struct LogicalIndexLut {
  unsigned long long base;
  unsigned long long t[4][256];
};
// same, ties resolved by LOGICAL index; iresult[0] is a logical (global) index
cudaError_t launch_argmax_permuted(const double2 *state, uint64_t n_amps,
                                   const LogicalIndexLut &lut, ReduceWorkspace &ws, cudaStream_t s);

// Exact replay of the reference's left-to-right rounded accumulation of
// |a_i|^2 (optionally only over indices with bit `mask_pos` clear; -1 = all).
// Three stream-ordered phases so that ranks can chain their shards:
//   launch_chunk_sums    K1: tree sum per 1024-term chunk; result[RES_APPROX_TOTAL] = shard total
//   launch_chunk_deltas  K2+K3: approximate prefix from *approx_start_dev, per-chunk increments
//   launch_chunk_resolve K4: exact walk from *exact_start_dev; chunk_exact[k] = running sum
//                        before chunk k (k = 0..n_chunks), result[RES_EXACT_TOTAL] = after the shard
enum { RES_SUM_RE = 0, RES_SUM_IM = 1, RES_TWO_MEAN_RE = 2, RES_TWO_MEAN_IM = 3,
       RES_APPROX_TOTAL = 4, RES_EXACT_TOTAL = 5, RES_ZERO = 6, RES_START = 7,
       RES_LOCAL_SUM_RE = 9, RES_LOCAL_SUM_IM = 10, RES_COUNT = 16 };
cudaError_t launch_chunk_sums(const double2 *state, uint64_t n_amps, int mask_pos,
                              ReduceWorkspace &ws, cudaStream_t s);
// K1 for SEL_RE and SEL_IM in one read of the shard (no result[RES_APPROX_TOTAL])
cudaError_t launch_chunk_sums_complex(const double2 *state, uint64_t n_amps, ReduceWorkspace &ws,
                                      cudaStream_t s);
// skip_if_zero (may be null): device word that is 0 when every term is an exact zero
// (launch_chunk_sums_complex leaves "some imaginary part is non-zero" in ws.iresult[2]): the kernels
// return at once and the total is the start value.  total_only: the caller reads
// result[RES_EXACT_TOTAL] and nothing else (chunk_exact[] is then not filled in for clean groups).
cudaError_t launch_chunk_deltas(const double2 *state, uint64_t n_amps, int mask_pos,
                                const double *approx_start_dev, ReduceWorkspace &ws,
                                cudaStream_t s, const long long *skip_if_zero = nullptr);
cudaError_t launch_chunk_resolve(const double2 *state, uint64_t n_amps, int mask_pos,
                                 const double *exact_start_dev, ReduceWorkspace &ws,
                                 cudaStream_t s, bool total_only = false,
                                 const long long *skip_if_zero = nullptr);
// Per shot: smallest i with u < running sum; -1 if none.  Needs launch_chunk_resolve first.
cudaError_t launch_sample(const double2 *state, uint64_t n_amps, const ReduceWorkspace &ws,
                          const double *u_dev, int shots, long long *idx_dev,
                          long long index_offset, cudaStream_t s);

}  // namespace qcs
