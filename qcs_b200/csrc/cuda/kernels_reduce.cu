// kernels_reduce.cu -- reductions, collapse, diffusion and the exact
// sequential-prefix machinery behind measurement and sampling.
//
// Reference loops replaced here:
//   prob_0 masked sum + collapse ....... reference src/qcs.c:253-281
//   q_state_normalize .................. reference src/q_utils.c:103-117
//   q_apply_diffusion .................. reference src/q_gates.c:334-355
//   q_apply_phase_flip ................. reference src/q_gates.c:311-316
//   qc_find_most_likely_state .......... reference src/qcs.c:464-478
//   qc_run_shots CDF scan .............. reference src/qcs.c:589-605
//
// Exactness.  The reference accumulates |a_i|^2 left to right in one double,
// and its decisions (`u <= prob_0`, `u < cumulative`, `total != 1.0`) depend
// on that rounded running sum.  A parallel tree sum differs in the last bits,
// so instead the running sum is REPLAYED exactly, in parallel:
//   while the running sum S stays inside one binade [2^e, 2^(e+1)) every
//   addition S (+) p rounds to a multiple of q = 2^(e-52), and (ties aside)
//   S (+) p - S = rne_q(p) does not depend on S.  So the increment a chunk of
//   1024 terms adds, D = (A (+) p_0 (+) ... (+) p_1023) - A, is the same for
//   any start A in that binade -- in particular for an approximate prefix A
//   from a parallel scan and for the true sequential S.
// Kernel 3 computes D for every chunk from the approximate prefix (twice,
// from A and from nextafter(A): the two runs have opposite parity, so they
// disagree exactly when a round-half-even tie occurred); kernel 4 walks the
// chunks in order with one addition each (exact: all values are multiples of
// q) and replays a chunk term by term only when a flag says the shortcut is
// unsafe (tie, binade crossing, wrong binade guess).  The result is the
// reference's sequence of partial sums, bit for bit, at chunk boundaries;
// the sampler then replays inside one chunk per shot.
//
// Signed sums.  Grover's diffusion sums the amplitudes themselves, real and
// imaginary parts, left to right (src/q_gates.c:334-336).  The same argument
// holds for signed terms as long as the running sum keeps its sign and binade
// THROUGHOUT a chunk; that is guaranteed when all terms of the chunk have one
// sign (the partial sums are then monotone and lie between the chunk's end
// points, which are checked), so a chunk whose terms have mixed signs is
// flagged and replayed term by term (`sel` = SEL_RE / SEL_IM below).  Grover's states have
// all non-solution amplitudes equal, so the fast path covers them; after a
// generic circuit most chunks replay (one warp, ~2^n dependent additions).
// A tree sum is NOT good enough here: the reference's sequential sum of 2^n
// equal terms drifts from the true mean (6e-10 relative after the 804
// iterations at 20 qubits), and parity is with the reference.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "gate_math.cuh"
#include "kernels.h"

namespace qcs {

namespace {

constexpr int RB = 256;  // reduction block size

__device__ __forceinline__ double norm_sq(double2 a) {
  return __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y));  // c_norm_sq, complex.c:89-93
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// same sign and same binade (for the non-negative sums of |a|^2 this is the exponent test)
__device__ __forceinline__ bool same_binade(double a, double b) {
  return (((unsigned)__double2hiint(a) ^ (unsigned)__double2hiint(b)) >> 20) == 0u;
}

unsigned grid_for(uint64_t items, int per_block) {
  uint64_t b = (items + per_block - 1) / per_block;
  if (b > REDUCE_MAX_BLOCKS) b = REDUCE_MAX_BLOCKS;
  if (b == 0) b = 1;
  return (unsigned)b;
}

// ---------------------------------------------------------------- fill / init
__global__ void init_state_kernel(double2 *state, uint64_t n, int set_one) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    state[i] = make_double2((set_one && i == 0) ? 1.0 : 0.0, 0.0);
}

// ---------------------------------------------------------------- complex sum
__global__ void __launch_bounds__(RB) complex_sum_kernel(const double2 *__restrict__ state,
                                                         uint64_t n, double *partials) {
  double sr = 0.0, si = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double2 a = __ldcs(state + i);
    sr += a.x;
    si += a.y;
  }
  __shared__ double sh[2][RB / 32];
  sr = warp_sum(sr);
  si = warp_sum(si);
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = sr;
    sh[1][threadIdx.x >> 5] = si;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tr = 0.0, ti = 0.0;
    for (int w = 0; w < RB / 32; w++) {
      tr += sh[0][w];
      ti += sh[1][w];
    }
    partials[2 * blockIdx.x] = tr;
    partials[2 * blockIdx.x + 1] = ti;
  }
}

__global__ void __launch_bounds__(1024) complex_sum_final_kernel(const double *partials, int nb,
                                                                 double *result) {
  double sr = 0.0, si = 0.0;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    sr += partials[2 * b];
    si += partials[2 * b + 1];
  }
  __shared__ double sh[2][32];
  sr = warp_sum(sr);
  si = warp_sum(si);
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = sr;
    sh[1][threadIdx.x >> 5] = si;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tr = 0.0, ti = 0.0;
    for (int w = 0; w < 32; w++) {
      tr += sh[0][w];
      ti += sh[1][w];
    }
    result[0] = tr;
    result[1] = ti;
  }
}

// ---------------------------------------------------------------- streaming writes
// mean of q_apply_diffusion (q_gates.c:337-346) from the raw sum, on the device
// so that Grover iterations never synchronise with the host.  sqrt and / are
// IEEE correctly rounded in double, like the host's.
__global__ void diffusion_mean_kernel(double *result, int corrected, double n_total) {
  double sr = result[0], si = result[1];
  if (corrected) {
    sr = __ddiv_rn(sr, n_total);
    si = __ddiv_rn(si, n_total);
  } else {
    const double mag = __dsqrt_rn(__dadd_rn(__dmul_rn(sr, sr), __dmul_rn(si, si)));
    if (mag > 1e-10) {
      sr = __ddiv_rn(sr, mag);
      si = __ddiv_rn(si, mag);
    }
  }
  result[2] = __dmul_rn(2.0, sr);
  result[3] = __dmul_rn(2.0, si);
}

__global__ void __launch_bounds__(256)
diffusion_write_kernel(const double2 *src, double2 *dst, uint64_t n,
                       const double *__restrict__ two_mean) {
  const double tr = two_mean[0], ti = two_mean[1];
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double2 a = __ldcs(src + i);
    __stcs(dst + i, make_double2(__dsub_rn(tr, a.x), __dsub_rn(ti, a.y)));
  }
}

// The same write, also summing the values it writes (per-block partial sums): the next Grover
// iteration's mean then needs no read pass of its own -- 32 instead of 48 bytes per amplitude per
// iteration.  (The reference sums sequentially, src/q_gates.c:334-336; this sum, like
// complex_sum_kernel's, is a tree: Grover parity is stated to 1e-12, not bit-exact.)
__global__ void __launch_bounds__(RB)
diffusion_write_sum_kernel(const double2 *src, double2 *dst, uint64_t n,
                           const double *__restrict__ two_mean, double *partials) {
  const double tr = two_mean[0], ti = two_mean[1];
  double sr = 0.0, si = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double2 a = __ldcs(src + i);
    const double2 v = make_double2(__dsub_rn(tr, a.x), __dsub_rn(ti, a.y));
    __stcs(dst + i, v);
    sr += v.x;
    si += v.y;
  }
  __shared__ double sh[2][RB / 32];
  sr = warp_sum(sr);
  si = warp_sum(si);
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = sr;
    sh[1][threadIdx.x >> 5] = si;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tr2 = 0.0, ti2 = 0.0;
    for (int w = 0; w < RB / 32; w++) {
      tr2 += sh[0][w];
      ti2 += sh[1][w];
    }
    partials[2 * blockIdx.x] = tr2;
    partials[2 * blockIdx.x + 1] = ti2;
  }
}

__global__ void __launch_bounds__(256) scale_kernel(double2 *state, uint64_t n, double f) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double2 a = __ldcs(state + i);
    __stcs(state + i, make_double2(__dmul_rn(a.x, f), __dmul_rn(a.y, f)));
  }
}

__global__ void __launch_bounds__(256)
zero_half_kernel(double2 *state, uint64_t n_items, int pos, uint64_t zero_bit) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += stride) {
    const uint64_t low = it & ((1ull << pos) - 1ull);
    const uint64_t i = (((it >> pos) << (pos + 1)) | low) | zero_bit;
    state[i] = make_double2(0.0, 0.0);
  }
}

// carried_sum (may be null): running sum of all amplitudes kept across Grover iterations; negating
// one amplitude takes it out twice
__global__ void negate_one_kernel(double2 *dst, const double2 *src, uint64_t index,
                                  double *carried_sum) {
  const double2 a = src[index];
  dst[index] = make_double2(-a.x, -a.y);
  if (carried_sum) {
    carried_sum[0] = (carried_sum[0] - a.x) - a.x;
    carried_sum[1] = (carried_sum[1] - a.y) - a.y;
  }
}

// ---------------------------------------------------------------- argmax
struct Best {
  double p;
  long long idx;
};
__device__ __forceinline__ Best better(Best a, Best b) {
  if (b.p > a.p || (b.p == a.p && b.idx < a.idx)) return b;
  return a;
}
__device__ __forceinline__ Best warp_best(Best v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best w;
    w.p = __shfl_xor_sync(0xffffffffu, v.p, o);
    w.idx = __shfl_xor_sync(0xffffffffu, v.idx, o);
    v = better(v, w);
  }
  return v;
}

__global__ void __launch_bounds__(RB) argmax_kernel(const double2 *__restrict__ state, uint64_t n,
                                                    double *partials, long long *ipartials) {
  Best b{0.0, 0x7fffffffffffffffll};
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double p = norm_sq(__ldcs(state + i));
    if (p > b.p) {  // strict: first maximum wins (qcs.c:472)
      b.p = p;
      b.idx = (long long)i;
    }
  }
  __shared__ double shp[RB / 32];
  __shared__ long long shi[RB / 32];
  b = warp_best(b);
  if ((threadIdx.x & 31) == 0) {
    shp[threadIdx.x >> 5] = b.p;
    shi[threadIdx.x >> 5] = b.idx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Best t{shp[0], shi[0]};
    for (int w = 1; w < RB / 32; w++) t = better(t, Best{shp[w], shi[w]});
    partials[blockIdx.x] = t.p;
    ipartials[blockIdx.x] = t.idx;
  }
}

// Second level of the argmax a fused pass started (QCS_PASS_ARGMAX): candidates instead of amplitudes.
__global__ void __launch_bounds__(RB)
argmax_pairs_kernel(const double *__restrict__ cand_p, const long long *__restrict__ cand_i, uint64_t n,
                    double *partials, long long *ipartials) {
  Best b{0.0, 0x7fffffffffffffffll};
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Best c{cand_p[i], cand_i[i]};
    if (c.p > 0.0) b = better(b, c);
  }
  __shared__ double shp[RB / 32];
  __shared__ long long shi[RB / 32];
  b = warp_best(b);
  if ((threadIdx.x & 31) == 0) {
    shp[threadIdx.x >> 5] = b.p;
    shi[threadIdx.x >> 5] = b.idx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Best t{shp[0], shi[0]};
    for (int w = 1; w < RB / 32; w++) t = better(t, Best{shp[w], shi[w]});
    partials[blockIdx.x] = t.p;
    ipartials[blockIdx.x] = t.idx;
  }
}

// The same over a shard whose index bits are permuted (multi-GPU layouts after position swaps):
// candidates are compared by their LOGICAL index, assembled from four byte-indexed tables
// (LogicalIndexLut, built by the engine from the current layout), so "first maximum wins" holds
// without restoring the identity layout first.
__global__ void __launch_bounds__(RB)
argmax_permuted_kernel(const double2 *__restrict__ state, uint64_t n,
                       const __grid_constant__ LogicalIndexLut lut, double *partials,
                       long long *ipartials) {
  __shared__ unsigned long long tab[4][256];
  for (int k = threadIdx.x; k < 4 * 256; k += blockDim.x) tab[k >> 8][k & 255] = lut.t[k >> 8][k & 255];
  __syncthreads();
  Best b{0.0, 0x7fffffffffffffffll};
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double p = norm_sq(__ldcs(state + i));
    if (p > 0.0 && p >= b.p) {
      const long long logical = (long long)(lut.base | tab[0][i & 255] | tab[1][(i >> 8) & 255] |
                                            tab[2][(i >> 16) & 255] | tab[3][(i >> 24) & 255]);
      if (p > b.p || logical < b.idx) {
        b.p = p;
        b.idx = logical;
      }
    }
  }
  __shared__ double shp[RB / 32];
  __shared__ long long shi[RB / 32];
  b = warp_best(b);
  if ((threadIdx.x & 31) == 0) {
    shp[threadIdx.x >> 5] = b.p;
    shi[threadIdx.x >> 5] = b.idx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Best t{shp[0], shi[0]};
    for (int w = 1; w < RB / 32; w++) t = better(t, Best{shp[w], shi[w]});
    partials[blockIdx.x] = t.p;
    ipartials[blockIdx.x] = t.idx;
  }
}

__global__ void argmax_final_kernel(const double *partials, const long long *ipartials, int nb,
                                    double *result, long long *iresult) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    Best t{0.0, 0x7fffffffffffffffll};
    for (int b = 0; b < nb; b++) t = better(t, Best{partials[b], ipartials[b]});
    result[0] = t.p;
    iresult[0] = (t.p > 0.0) ? t.idx : 0;  // all-zero state -> 0 (qcs.c:465-466)
  }
}

// ---------------------------------------------------------------- exact prefix
enum { FLAG_TIE = 1, FLAG_CROSS = 2, FLAG_ZERO = 4 };

// Term i of the sequential sum selected by `mask_pos` (kernels.h): >= 0 |a_i|^2 over indices with that
// bit clear, SEL_ALL |a_i|^2, SEL_RE / SEL_IM the real / imaginary part of a_i (signed).
__device__ __forceinline__ double masked_norm(const double2 *__restrict__ state, uint64_t i,
                                              int mask_pos) {
  if (mask_pos <= SEL_RE) {
    const double2 a = __ldg(state + i);
    return mask_pos == SEL_RE ? a.x : a.y;
  }
  if (mask_pos >= 0 && ((i >> mask_pos) & 1ull)) return 0.0;
  return norm_sq(__ldg(state + i));
}
// The same from an amplitude already in registers.  Kernels that are short of warps (one warp per 32
// chunks, or one warp in all) fetch a batch of amplitudes with back-to-back loads first and turn them
// into terms afterwards: with masked_norm's branches between the loads the compiler serialises them,
// one memory latency per term (the 24-qubit Grover search spent 0.57 of its 1.3 ms per iteration
// that way, profiles/r2l_launches_grover24.csv).
__device__ __forceinline__ double term_of(double2 a, uint64_t i, int mask_pos) {
  if (mask_pos <= SEL_RE) return mask_pos == SEL_RE ? a.x : a.y;
  const double p = norm_sq(a);
  return (mask_pos >= 0 && ((i >> mask_pos) & 1ull)) ? 0.0 : p;
}

// K1: plain (tree) sum of each 1024-term chunk; one warp per chunk.
__global__ void __launch_bounds__(256)
chunk_sum_kernel(const double2 *__restrict__ state, uint64_t n_chunks, int mask_pos,
                 double *chunk_sum, double *chunk_abs) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t c = warp; c < n_chunks; c += n_warps) {
    double s = 0.0, sa = 0.0;
#pragma unroll 8
    for (int j = 0; j < SEQ_CHUNK / 32; j++) {
      const double p = masked_norm(state, c * SEQ_CHUNK + (uint64_t)j * 32 + lane, mask_pos);
      s += p;
      sa += fabs(p);
    }
    s = warp_sum(s);
    if (chunk_abs) sa = warp_sum(sa);
    if (lane == 0) {
      chunk_sum[c] = s;
      if (chunk_abs) chunk_abs[c] = sa;  // signed terms: "all zero" cannot be read off a sum that may cancel
    }
  }
}

// K1 for both components of the amplitudes in one read (diffusion): tree sums and sums of magnitudes
// of the real parts (sum_re, abs_re) and of the imaginary parts (sum_im, abs_im) per chunk.
__global__ void __launch_bounds__(256)
chunk_sum_complex_kernel(const double2 *__restrict__ state, uint64_t n_chunks, double *sum_re,
                         double *abs_re, double *sum_im, double *abs_im, long long *im_nonzero) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t c = warp; c < n_chunks; c += n_warps) {
    double sr = 0.0, si = 0.0, ar = 0.0, ai = 0.0;
#pragma unroll 8
    for (int j = 0; j < SEQ_CHUNK / 32; j++) {
      const double2 a = __ldcs(state + c * SEQ_CHUNK + (uint64_t)j * 32 + lane);
      sr += a.x;
      si += a.y;
      ar += fabs(a.x);
      ai += fabs(a.y);
    }
    sr = warp_sum(sr);
    si = warp_sum(si);
    ar = warp_sum(ar);
    ai = warp_sum(ai);
    if (lane == 0) {
      sum_re[c] = sr;
      sum_im[c] = si;
      abs_re[c] = ar;
      abs_im[c] = ai;
      if (ai != 0.0) *im_nonzero = 1;  // (racing writers all store the same value)
    }
  }
}

// K2: approximate exclusive prefix over chunks (single block, 1024 threads): coalesced tiles of 1024
// chunk sums, a block-wide scan per tile (warp shuffles + one shared-memory hop), the running total
// carried from tile to tile.  (The prefix is approximate by design -- any summation order will do.)
// skip_if_zero (may be null): a device word that is 0 when every term of the sum is an exact zero
// (chunk_sum_complex_kernel's "some imaginary part is non-zero" flag): nothing to do then.
__global__ void __launch_bounds__(1024)
chunk_scan_kernel(const double *chunk_sum, uint64_t n_chunks, const double *start_dev,
                  double *approx, const long long *skip_if_zero) {
  if (skip_if_zero && *skip_if_zero == 0) return;
  __shared__ double warp_tot[32];
  __shared__ double carry_sh;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double carry = *start_dev;
  for (uint64_t base = 0; base < n_chunks; base += 1024) {
    const uint64_t c = base + threadIdx.x;
    const double v = c < n_chunks ? chunk_sum[c] : 0.0;
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
      const double t = warp_tot[lane];
      double ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      warp_tot[lane] = ti - t;  // exclusive over warps
      if (lane == 31) carry_sh = ti;
    }
    __syncthreads();
    if (c < n_chunks) approx[c] = carry + warp_tot[w] + (incl - v);
    carry += carry_sh;
    __syncthreads();
  }
}

// K3: per-chunk increment D and safety flags.  One warp handles 32 chunks;
// 32x32 blocks of |a|^2 are staged through shared memory so global loads stay
// coalesced while each lane walks ITS chunk strictly left to right.
__global__ void __launch_bounds__(128)
chunk_delta_kernel(const double2 *__restrict__ state, uint64_t n_chunks, int mask_pos,
                   const double *__restrict__ chunk_sum, const double *__restrict__ approx,
                   double *__restrict__ delta, unsigned char *__restrict__ flag,
                   const long long *skip_if_zero) {
  if (skip_if_zero && *skip_if_zero == 0) return;  // every term is an exact zero (chunk_scan_kernel)
  // chunk_sum: what decides "this chunk adds nothing" -- the chunk's sum for non-negative terms, the
  // sum of magnitudes for signed ones (the launcher passes the right array)
  const bool is_signed = mask_pos <= SEL_RE;
  __shared__ double stage[4][32][33];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const uint64_t group = (uint64_t)blockIdx.x * 4 + wib;  // 32 chunks per group
  const uint64_t c0 = group * 32;
  if (c0 >= n_chunks) return;
  const uint64_t my_chunk = c0 + lane;
  const bool have = my_chunk < n_chunks;
  const double csum = have ? chunk_sum[my_chunk] : 0.0;
  const bool any_work = __any_sync(0xffffffffu, have && csum != 0.0);
  if (!any_work) {
    if (have) {
      delta[my_chunk] = 0.0;
      flag[my_chunk] = FLAG_ZERO;
    }
    return;
  }
  const double a1 = have ? approx[my_chunk] : 1.0;
  const double a2 = __longlong_as_double(__double_as_longlong(a1) + 1);  // next double away from zero
  double s1 = a1, s2 = a2;
  bool any_pos = false, any_neg = false;
  const int n_rows = (int)((n_chunks - c0) < 32 ? (n_chunks - c0) : 32);
  for (int j = 0; j < SEQ_CHUNK / 32; j++) {
    // all 32 row loads of this step are issued before the first is used: on small shards (a 24-qubit
    // Grover search has 512 warps in this kernel) nothing else hides their latency
    double2 raw[32];
#pragma unroll
    for (int c = 0; c < 32; c++) {
      const uint64_t i = (c0 + (uint64_t)(c < n_rows ? c : 0)) * SEQ_CHUNK + (uint64_t)j * 32 + lane;
      raw[c] = __ldg(state + i);  // rows past the end re-read row 0 (discarded below): no branch between the loads
    }
#pragma unroll
    for (int c = 0; c < 32; c++) {
      const uint64_t i = (c0 + c) * SEQ_CHUNK + (uint64_t)j * 32 + lane;
      stage[wib][c][lane] = c < n_rows ? term_of(raw[c], i, mask_pos) : 0.0;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 32; k++) {
      const double p = stage[wib][lane][k];
      s1 = __dadd_rn(s1, p);
      s2 = __dadd_rn(s2, p);
      any_pos |= p > 0.0;
      any_neg |= p < 0.0;
    }
    __syncwarp();
  }
  if (have) {
    const double d1 = __dsub_rn(s1, a1);
    const double d2 = __dsub_rn(s2, a2);
    unsigned char f = 0;
    if (csum == 0.0) f |= FLAG_ZERO;
    if (d1 != d2) f |= FLAG_TIE;
    if (!same_binade(s1, a1) || !same_binade(s2, a1)) f |= FLAG_CROSS;
    // signed terms: the end points only vouch for the whole chunk when the running sum is monotone,
    // i.e. all terms of the chunk have one sign (whichever side of zero the sum is on: a negative sum
    // climbing through positive terms stays between its end points just the same)
    if (is_signed && any_pos && any_neg) f |= FLAG_CROSS;
    delta[my_chunk] = d1;
    flag[my_chunk] = f;
  }
}

// Term-by-term replay of one chunk of at most SEQ_CHUNK terms (all lanes return the same S).  The
// warp fetches the whole chunk with 32 independent coalesced loads (one memory latency instead of 32
// in a row: the walk that calls this is a single warp); the terms then reach the additions through
// shuffles that do not depend on S, so the dependent chain is the 1 024 additions alone.  Positions
// past `len` contribute +0.0 (S + 0.0 == S).
__device__ __forceinline__ double replay_chunk(const double2 *__restrict__ state, uint64_t first,
                                               uint64_t len, int mask_pos, double S, int lane) {
  double2 raw[SEQ_CHUNK / 32];
#pragma unroll
  for (int j = 0; j < SEQ_CHUNK / 32; j++) {
    const uint64_t off = (uint64_t)j * 32 + lane;
    raw[j] = __ldg(state + first + (off < len ? off : 0));  // no branch between the loads
  }
  double p[SEQ_CHUNK / 32];
#pragma unroll
  for (int j = 0; j < SEQ_CHUNK / 32; j++) {
    const uint64_t off = (uint64_t)j * 32 + lane;
    p[j] = off < len ? term_of(raw[j], first + off, mask_pos) : 0.0;
  }
#pragma unroll
  for (int j = 0; j < SEQ_CHUNK / 32; j++) {
#pragma unroll
    for (int l = 0; l < 32; l++) S = __dadd_rn(S, __shfl_sync(0xffffffffu, p[j], l));
  }
  return S;
}

// ---- K4, two levels (shards with many chunks) ---------------------------------------------
// A one-warp walk over 2^20 chunks (30 qubits) costs ~100 ms.  The shortcut the walk takes per chunk
// composes: while nothing is flagged and the running sum stays in one binade, every step is an
// EXACT addition of a multiple of that binade's quantum, so a whole group of chunks advances the
// sum by the exact total of its increments.  K4a sums the increments of each group of
// RESOLVE_GROUP chunks in parallel and marks groups that contain a flagged chunk or span a binade
// of the approximate prefix; K4b walks the (few) groups, falling back to the per-chunk walk inside
// marked groups or when the running sum leaves the binade; K4c fills the per-chunk running sums
// of the clean groups in parallel (the sampler needs them).
constexpr int RESOLVE_GROUP = QCS_RESOLVE_GROUP;

__global__ void __launch_bounds__(256)
group_sum_kernel(const double *__restrict__ delta, const double *__restrict__ approx,
                 const unsigned char *__restrict__ flag, uint64_t n_chunks, uint64_t n_groups,
                 double *__restrict__ gsum, unsigned char *__restrict__ gflag, const long long *skip_if_zero) {
  if (skip_if_zero && *skip_if_zero == 0) return;
  const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_groups) return;
  const uint64_t first = g * RESOLVE_GROUP;
  const uint64_t last = first + RESOLVE_GROUP < n_chunks ? first + RESOLVE_GROUP : n_chunks;
  double s = 0.0;
  int f = 0, signs = 0;
  for (uint64_t c = first + lane; c < last; c += 32) {
    const int fc = flag[c];
    f |= fc;
    if (!(fc & FLAG_ZERO)) {
      const double d = delta[c];
      s += d;  // exact: multiples of one quantum, total below the binade's top
      signs |= (d > 0.0 ? 1 : 0) | (d < 0.0 ? 2 : 0);
    }
  }
  s = warp_sum(s);
  f = __reduce_or_sync(0xffffffffu, f);
  signs = __reduce_or_sync(0xffffffffu, signs);
  if (lane == 0) {
    unsigned char out = 0;
    if (f & (FLAG_TIE | FLAG_CROSS)) out = 1;
    if (signs == 3) out = 1;  // signed sums: the running sum is not monotone across the group
    if (!same_binade(approx[first], approx[last - 1])) out = 1;
    gsum[g] = s;
    gflag[g] = out;
  }
}

// the per-chunk walk over chunks [c_begin, c_end), shared by the one-level kernel and K4b
__device__ __forceinline__ double walk_chunks(const double2 *__restrict__ state, uint64_t c_begin,
                                              uint64_t c_end, uint64_t chunk_len, int mask_pos,
                                              const double *__restrict__ approx,
                                              const double *__restrict__ delta,
                                              const unsigned char *__restrict__ flag,
                                              int have_deltas, double *__restrict__ exact, double S,
                                              int lane, long long &n_replay) {
  double nx_d = 0.0, nx_a = 0.0;
  int nx_f = FLAG_CROSS;
  if (have_deltas && c_begin + lane < c_end) {
    nx_d = delta[c_begin + lane];
    nx_a = approx[c_begin + lane];
    nx_f = flag[c_begin + lane];
  }
  for (uint64_t base = c_begin; base < c_end; base += 32) {
    const uint64_t mine = base + lane;
    const double my_d = nx_d, my_a = nx_a;
    const int my_f = nx_f;
    // the next block's records are fetched while this one is processed (one warp walks alone here)
    nx_d = 0.0; nx_a = 0.0; nx_f = FLAG_CROSS;
    if (have_deltas && mine + 32 < c_end) {
      nx_d = delta[mine + 32];
      nx_a = approx[mine + 32];
      nx_f = flag[mine + 32];
    }
    double my_exact = 0.0;
    const int lim = (int)((c_end - base) < 32 ? (c_end - base) : 32);
    if (have_deltas) {
      // Shortcut over the whole block of 32 chunks: when none of them is flagged, their approximate
      // prefixes share the running sum's binade and their increments share one sign, the walk below
      // would perform 32 EXACT additions of multiples of the binade's quantum (monotone, so the end
      // point vouches for every step) -- the same numbers come out of one warp scan.
      const bool live = mine < c_end && !(my_f & FLAG_ZERO);
      const unsigned nz = __ballot_sync(0xffffffffu, live);
      if (nz == 0u) {
        if (mine < c_end) exact[mine] = S;
        continue;
      }
      const double a_ref = __shfl_sync(0xffffffffu, my_a, __ffs((int)nz) - 1);
      const unsigned bad = __ballot_sync(0xffffffffu, live && ((my_f & (FLAG_TIE | FLAG_CROSS)) || !same_binade(my_a, a_ref)));
      const unsigned pos = __ballot_sync(0xffffffffu, live && my_d > 0.0);
      const unsigned neg = __ballot_sync(0xffffffffu, live && my_d < 0.0);
      if (bad == 0u && !(pos && neg) && same_binade(S, a_ref)) {
        const double dz = live ? my_d : 0.0;
        double incl = dz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const double Snew = __dadd_rn(S, __shfl_sync(0xffffffffu, incl, 31));
        if (same_binade(Snew, S)) {
          if (mine < c_end) exact[mine] = S + (incl - dz);
          S = Snew;
          continue;
        }
      }
    }
    for (int j = 0; j < lim; j++) {
      const double d = __shfl_sync(0xffffffffu, my_d, j);
      const double a = __shfl_sync(0xffffffffu, my_a, j);
      const int f = __shfl_sync(0xffffffffu, my_f, j);
      if (lane == j) my_exact = S;
      if (f & FLAG_ZERO) continue;
      bool ok = !(f & (FLAG_TIE | FLAG_CROSS)) && same_binade(S, a);
      double Snew = S;
      if (ok) {
        Snew = __dadd_rn(S, d);
        ok = same_binade(Snew, S);
      }
      if (ok) {
        S = Snew;
      } else {
        S = replay_chunk(state, (base + j) * chunk_len, chunk_len, mask_pos, S, lane);
        n_replay++;
      }
    }
    if (mine < c_end) exact[mine] = my_exact;
  }
  return S;
}

// K4b: one warp walks the groups; gexact[g] = running sum before group g, gflag[g] = 2 when the
// per-chunk walk already wrote that group's chunk_exact entries.  The groups' records (flag, sum,
// approximate prefix of the first chunk) are fetched 32 groups at a time, one per lane; a block of
// 32 clean groups whose prefixes share the running sum's binade and whose sums share one sign is
// one warp scan (the argument of walk_chunks' block shortcut, one level up: exact additions of
// multiples of the binade's quantum, monotone, so the end point vouches for every step).
// skip_if_zero: see chunk_scan_kernel -- the total is then the start value (adding exact zeros to it
// changes nothing: the start is +0 or a non-zero running sum).
__global__ void __launch_bounds__(32)
group_resolve_kernel(const double2 *__restrict__ state, uint64_t n_chunks, uint64_t n_groups,
                     uint64_t chunk_len, int mask_pos, const double *start_dev,
                     const double *__restrict__ approx, const double *__restrict__ delta,
                     const unsigned char *__restrict__ flag, const double *__restrict__ gsum,
                     unsigned char *__restrict__ gflag, double *__restrict__ gexact,
                     double *__restrict__ exact, double *total_out, long long *replays,
                     const long long *skip_if_zero) {
  const int lane = threadIdx.x;
  double S = *start_dev;
  if (skip_if_zero && *skip_if_zero == 0) {
    if (lane == 0) {
      *total_out = S;
      if (replays) *replays = 0;
    }
    return;
  }
  long long n_replay = 0;
  for (uint64_t gb = 0; gb < n_groups; gb += 32) {
    const uint64_t g_mine = gb + lane;
    const bool have = g_mine < n_groups;
    int my_gf = 1;
    double my_gs = 0.0, my_ga = 0.0, my_gexact = 0.0;
    if (have) {
      my_gf = gflag[g_mine];
      my_gs = gsum[g_mine];
      my_ga = approx[g_mine * RESOLVE_GROUP];
    }
    const int lim = (int)((n_groups - gb) < 32 ? (n_groups - gb) : 32);
    {
      const unsigned bad = __ballot_sync(0xffffffffu, have && (my_gf != 0 || !same_binade(my_ga, S)));
      const unsigned pos = __ballot_sync(0xffffffffu, have && my_gs > 0.0);
      const unsigned neg = __ballot_sync(0xffffffffu, have && my_gs < 0.0);
      if (bad == 0u && !(pos && neg)) {
        const double dz = have ? my_gs : 0.0;
        double incl = dz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const double Snew = __dadd_rn(S, __shfl_sync(0xffffffffu, incl, 31));
        if (same_binade(Snew, S)) {
          if (have) gexact[g_mine] = S + (incl - dz);
          S = Snew;
          continue;
        }
      }
    }
    for (int j = 0; j < lim; j++) {
      const int gf = __shfl_sync(0xffffffffu, my_gf, j);
      const double gs = __shfl_sync(0xffffffffu, my_gs, j);
      const double ga = __shfl_sync(0xffffffffu, my_ga, j);
      bool clean = gf == 0 && same_binade(S, ga);
      double Snew = S;
      if (clean) {
        Snew = __dadd_rn(S, gs);
        clean = same_binade(Snew, S);
      }
      if (lane == j) my_gexact = S;
      if (clean) {
        S = Snew;  // K4c fills this group's chunk_exact entries
      } else {
        const uint64_t first = (gb + (uint64_t)j) * RESOLVE_GROUP;
        const uint64_t last = first + RESOLVE_GROUP < n_chunks ? first + RESOLVE_GROUP : n_chunks;
        S = walk_chunks(state, first, last, chunk_len, mask_pos, approx, delta, flag, 1, exact, S, lane,
                        n_replay);
        if (lane == j) my_gf = 2;  // done here
      }
    }
    if (have) {
      gexact[g_mine] = my_gexact;
      if (my_gf == 2) gflag[g_mine] = 2;
    }
  }
  if (lane == 0) {
    exact[n_chunks] = S;
    *total_out = S;
    if (replays) *replays = n_replay;
  }
}

// K4c: exact[c] = gexact[g] + (increments of the group's earlier chunks) for the clean groups.
__global__ void __launch_bounds__(256)
group_fill_kernel(const double *__restrict__ delta, const unsigned char *__restrict__ flag,
                  uint64_t n_chunks, uint64_t n_groups, const unsigned char *__restrict__ gflag,
                  const double *__restrict__ gexact, double *__restrict__ exact) {
  const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_groups || gflag[g] == 2) return;
  const uint64_t first = g * RESOLVE_GROUP;
  const uint64_t last = first + RESOLVE_GROUP < n_chunks ? first + RESOLVE_GROUP : n_chunks;
  // lane l owns chunks [first + 32 l, first + 32 l + 32)
  const uint64_t lo = first + (uint64_t)lane * (RESOLVE_GROUP / 32);
  double mine = 0.0;
  for (uint64_t c = lo; c < lo + RESOLVE_GROUP / 32 && c < last; c++)
    if (!(flag[c] & FLAG_ZERO)) mine += delta[c];
  double incl = mine;  // inclusive scan over lanes (exact additions, any order)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  double run = gexact[g] + (incl - mine);
  for (uint64_t c = lo; c < lo + RESOLVE_GROUP / 32 && c < last; c++) {
    exact[c] = run;
    if (!(flag[c] & FLAG_ZERO)) run += delta[c];
  }
}

// K4 (one level): walk the chunks in order (one warp; every lane carries the same S).
__global__ void __launch_bounds__(32)
chunk_resolve_kernel(const double2 *__restrict__ state, uint64_t n_amps, uint64_t n_chunks,
                     uint64_t chunk_len, int mask_pos, const double *start_dev,
                     const double *__restrict__ approx, const double *__restrict__ delta,
                     const unsigned char *__restrict__ flag, int have_deltas,
                     double *__restrict__ exact, double *total_out, long long *replays,
                     const long long *skip_if_zero) {
  const int lane = threadIdx.x;
  if (skip_if_zero && *skip_if_zero == 0) {
    if (lane == 0) {
      *total_out = *start_dev;
      if (replays) *replays = 0;
    }
    return;
  }
  long long n_replay = 0;
  const double S = walk_chunks(state, 0, n_chunks, chunk_len, mask_pos, approx, delta, flag,
                               have_deltas, exact, *start_dev, lane, n_replay);
  if (lane == 0) {
    exact[n_chunks] = S;
    *total_out = S;
    if (replays) *replays = n_replay;
  }
  (void)n_amps;
}

// Total of the chunk sums (approximate shard total, for cross-rank offsets).
__global__ void __launch_bounds__(1024)
chunk_total_kernel(const double *chunk_sum, uint64_t n_chunks, double *out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (uint64_t c = threadIdx.x; c < n_chunks; c += blockDim.x) s += chunk_sum[c];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; w++) t += sh[w];
    *out = t;
  }
}

// K5: one thread per shot.
__global__ void __launch_bounds__(128)
sample_kernel(const double2 *__restrict__ state, uint64_t n_chunks, uint64_t chunk_len,
              const double *__restrict__ exact, const double *__restrict__ u_dev, int shots,
              long long *__restrict__ idx_out, long long index_offset) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= shots) return;
  const double u = u_dev[k];
  long long found = -1;
  // Not ours if an earlier shard already satisfied u < sum, or no prefix here exceeds u.
  if (!(u < exact[0]) && (u < exact[n_chunks])) {
    // smallest c with u < exact[c + 1]
    uint64_t lo = 0, hi = n_chunks - 1;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (u < exact[mid + 1]) hi = mid; else lo = mid + 1;
    }
    double S = exact[lo];
    const uint64_t first = lo * chunk_len;
    for (uint64_t i = 0; i < chunk_len; i++) {
      S = __dadd_rn(S, norm_sq(__ldg(state + first + i)));
      if (u < S) {  // strict, qcs.c:600
        found = (long long)(first + i) + index_offset;
        break;
      }
    }
  }
  idx_out[k] = found;
}

}  // namespace

// =============================================================== launchers
cudaError_t launch_init_state(double2 *state, uint64_t n, bool set_one, cudaStream_t s) {
  init_state_kernel<<<grid_for(n, 256 * 8), 256, 0, s>>>(state, n, set_one ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_complex_sum(const double2 *state, uint64_t n, ReduceWorkspace &ws,
                               cudaStream_t s) {
  const unsigned nb = grid_for(n, RB * 8);
  complex_sum_kernel<<<nb, RB, 0, s>>>(state, n, ws.partials);
  complex_sum_final_kernel<<<1, 1024, 0, s>>>(ws.partials, (int)nb, ws.result + RES_LOCAL_SUM_RE);
  return cudaGetLastError();
}

cudaError_t launch_diffusion_mean(ReduceWorkspace &ws, bool corrected, double n_total,
                                  cudaStream_t s) {
  diffusion_mean_kernel<<<1, 1, 0, s>>>(ws.result, corrected ? 1 : 0, n_total);
  return cudaGetLastError();
}

cudaError_t launch_diffusion_write_sum(const double2 *src, double2 *dst, uint64_t n,
                                       ReduceWorkspace &ws, cudaStream_t s) {
  const unsigned nb = grid_for(n, RB * 8);
  diffusion_write_sum_kernel<<<nb, RB, 0, s>>>(src, dst, n, ws.result + 2, ws.partials);
  complex_sum_final_kernel<<<1, 1024, 0, s>>>(ws.partials, (int)nb, ws.result + RES_LOCAL_SUM_RE);
  return cudaGetLastError();
}

cudaError_t launch_diffusion_write(const double2 *src, double2 *dst, uint64_t n,
                                   const ReduceWorkspace &ws, cudaStream_t s) {
  diffusion_write_kernel<<<grid_for(n, 256 * 4), 256, 0, s>>>(src, dst, n, ws.result + 2);
  return cudaGetLastError();
}

cudaError_t launch_scale(double2 *state, uint64_t n, double f, cudaStream_t s) {
  scale_kernel<<<grid_for(n, 256 * 4), 256, 0, s>>>(state, n, f);
  return cudaGetLastError();
}

cudaError_t launch_zero_half(double2 *state, uint64_t n, int pos, int keep_bit, cudaStream_t s) {
  const uint64_t items = n >> 1;
  const uint64_t zero_bit = keep_bit ? 0ull : (1ull << pos);
  zero_half_kernel<<<grid_for(items, 256 * 4), 256, 0, s>>>(state, items, pos, zero_bit);
  return cudaGetLastError();
}

cudaError_t launch_negate_one(double2 *dst, const double2 *src, uint64_t index,
                              double *carried_sum, cudaStream_t s) {
  negate_one_kernel<<<1, 1, 0, s>>>(dst, src, index, carried_sum);
  return cudaGetLastError();
}

cudaError_t launch_argmax(const double2 *state, uint64_t n, ReduceWorkspace &ws, cudaStream_t s) {
  const unsigned nb = grid_for(n, RB * 8);
  argmax_kernel<<<nb, RB, 0, s>>>(state, n, ws.partials, ws.ipartials);
  argmax_final_kernel<<<1, 32, 0, s>>>(ws.partials, ws.ipartials, (int)nb, ws.result, ws.iresult);
  return cudaGetLastError();
}

cudaError_t launch_argmax_pairs(const double *p, const long long *idx, uint64_t n, ReduceWorkspace &ws,
                                cudaStream_t s) {
  const unsigned nb = grid_for(n, RB * 8);
  argmax_pairs_kernel<<<nb, RB, 0, s>>>(p, idx, n, ws.partials, ws.ipartials);
  argmax_final_kernel<<<1, 32, 0, s>>>(ws.partials, ws.ipartials, (int)nb, ws.result, ws.iresult);
  return cudaGetLastError();
}

cudaError_t launch_argmax_permuted(const double2 *state, uint64_t n, const LogicalIndexLut &lut,
                                   ReduceWorkspace &ws, cudaStream_t s) {
  const unsigned nb = grid_for(n, RB * 8);
  argmax_permuted_kernel<<<nb, RB, 0, s>>>(state, n, lut, ws.partials, ws.ipartials);
  argmax_final_kernel<<<1, 32, 0, s>>>(ws.partials, ws.ipartials, (int)nb, ws.result, ws.iresult);
  return cudaGetLastError();
}

// the arrays a selector works on: tree sums of the chunks, and what decides "adds nothing"
static double *sums_of(ReduceWorkspace &ws, int sel) { return sel == SEL_IM ? ws.chunk_sum2 : ws.chunk_sum; }
static double *zero_test_of(ReduceWorkspace &ws, int sel) {
  return sel == SEL_IM ? ws.chunk_abs2 : sel == SEL_RE ? ws.chunk_abs : ws.chunk_sum;
}

cudaError_t launch_chunk_sums(const double2 *state, uint64_t n, int mask_pos,
                              ReduceWorkspace &ws, cudaStream_t s) {
  if (n < (uint64_t)SEQ_CHUNK) return cudaSuccess;  // single short chunk: resolved by replay
  const uint64_t n_chunks = n / SEQ_CHUNK;
  uint64_t blocks = (n_chunks + 7) / 8;
  if (blocks > (unsigned)device_sm_count() * 32) blocks = (unsigned)device_sm_count() * 32;
  chunk_sum_kernel<<<(unsigned)blocks, 256, 0, s>>>(state, n_chunks, mask_pos, sums_of(ws, mask_pos),
                                                    mask_pos <= SEL_RE ? zero_test_of(ws, mask_pos) : nullptr);
  chunk_total_kernel<<<1, 1024, 0, s>>>(sums_of(ws, mask_pos), n_chunks, ws.result + RES_APPROX_TOTAL);
  return cudaGetLastError();
}

cudaError_t launch_chunk_sums_complex(const double2 *state, uint64_t n, ReduceWorkspace &ws, cudaStream_t s) {
  // shards below one chunk have no chunk sums: the flag says "do not skip" (any non-zero bytes)
  if (n < (uint64_t)SEQ_CHUNK) return cudaMemsetAsync(ws.iresult + 2, 1, sizeof(long long), s);
  const uint64_t n_chunks = n / SEQ_CHUNK;
  uint64_t blocks = (n_chunks + 7) / 8;
  if (blocks > (unsigned)device_sm_count() * 32) blocks = (unsigned)device_sm_count() * 32;
  cudaError_t err = cudaMemsetAsync(ws.iresult + 2, 0, sizeof(long long), s);
  if (err != cudaSuccess) return err;
  chunk_sum_complex_kernel<<<(unsigned)blocks, 256, 0, s>>>(state, n_chunks, ws.chunk_sum, ws.chunk_abs,
                                                            ws.chunk_sum2, ws.chunk_abs2, ws.iresult + 2);
  return cudaGetLastError();
}

cudaError_t launch_chunk_deltas(const double2 *state, uint64_t n, int mask_pos,
                                const double *approx_start_dev, ReduceWorkspace &ws,
                                cudaStream_t s, const long long *skip_if_zero) {
  if (n < (uint64_t)SEQ_CHUNK) return cudaSuccess;
  const uint64_t n_chunks = n / SEQ_CHUNK;
  chunk_scan_kernel<<<1, 1024, 0, s>>>(sums_of(ws, mask_pos), n_chunks, approx_start_dev, ws.chunk_approx,
                                       skip_if_zero);
  const uint64_t groups = (n_chunks + 31) / 32;
  const uint64_t blocks = (groups + 3) / 4;
  chunk_delta_kernel<<<(unsigned)blocks, 128, 0, s>>>(state, n_chunks, mask_pos, zero_test_of(ws, mask_pos),
                                                      ws.chunk_approx, ws.chunk_delta,
                                                      ws.chunk_flag, skip_if_zero);
  return cudaGetLastError();
}

cudaError_t launch_chunk_resolve(const double2 *state, uint64_t n, int mask_pos,
                                 const double *exact_start_dev, ReduceWorkspace &ws,
                                 cudaStream_t s, bool total_only, const long long *skip_if_zero) {
  if (n < (uint64_t)SEQ_CHUNK) {
    chunk_resolve_kernel<<<1, 32, 0, s>>>(state, n, 1, n, mask_pos, exact_start_dev, nullptr,
                                          nullptr, nullptr, 0, ws.chunk_exact,
                                          ws.result + RES_EXACT_TOTAL, ws.iresult + 1, skip_if_zero);
    return cudaGetLastError();
  }
  const uint64_t n_chunks = n / SEQ_CHUNK;
  if (n_chunks < 4 * (uint64_t)RESOLVE_GROUP) {
    chunk_resolve_kernel<<<1, 32, 0, s>>>(state, n, n_chunks, SEQ_CHUNK, mask_pos, exact_start_dev,
                                          ws.chunk_approx, ws.chunk_delta, ws.chunk_flag, 1,
                                          ws.chunk_exact, ws.result + RES_EXACT_TOTAL,
                                          ws.iresult + 1, skip_if_zero);
    return cudaGetLastError();
  }
  const uint64_t n_groups = (n_chunks + RESOLVE_GROUP - 1) / RESOLVE_GROUP;
  const unsigned blocks = (unsigned)((n_groups + 7) / 8);  // 8 warps per block, one group per warp
  group_sum_kernel<<<blocks, 256, 0, s>>>(ws.chunk_delta, ws.chunk_approx, ws.chunk_flag, n_chunks,
                                          n_groups, ws.group_sum, ws.group_flag, skip_if_zero);
  group_resolve_kernel<<<1, 32, 0, s>>>(state, n_chunks, n_groups, SEQ_CHUNK, mask_pos,
                                        exact_start_dev, ws.chunk_approx, ws.chunk_delta,
                                        ws.chunk_flag, ws.group_sum, ws.group_flag, ws.group_exact,
                                        ws.chunk_exact, ws.result + RES_EXACT_TOTAL, ws.iresult + 1,
                                        skip_if_zero);
  // the running sum in front of every chunk: what sampling and collapse read; a caller that wants the
  // total alone (the diffusion mean) skips it
  if (!total_only)
    group_fill_kernel<<<blocks, 256, 0, s>>>(ws.chunk_delta, ws.chunk_flag, n_chunks, n_groups,
                                             ws.group_flag, ws.group_exact, ws.chunk_exact);
  return cudaGetLastError();
}

cudaError_t launch_sample(const double2 *state, uint64_t n, const ReduceWorkspace &ws,
                          const double *u_dev, int shots, long long *idx_dev,
                          long long index_offset, cudaStream_t s) {
  const uint64_t chunk_len = n < (uint64_t)SEQ_CHUNK ? n : (uint64_t)SEQ_CHUNK;
  const uint64_t n_chunks = n / chunk_len;
  sample_kernel<<<(shots + 127) / 128, 128, 0, s>>>(state, n_chunks, chunk_len, ws.chunk_exact,
                                                    u_dev, shots, idx_dev, index_offset);
  return cudaGetLastError();
}

}  // namespace qcs
