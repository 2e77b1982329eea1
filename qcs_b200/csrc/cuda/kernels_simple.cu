// kernels_simple.cu -- one launch per gate, in place.
//
// The unfused path: used when QCS_CUDA_FUSION=off (parity cross-check of the
// fused scheduler) and for shards smaller than one tile (fewer than 12 local
// qubits).  Direct counterpart of the reference's per-gate sweeps
// (q_apply_1q_gate reference src/q_gates.c:131-143, q_apply_2q_gate :276-292),
// minus the memcpy: every pair is read once and written once.
// Algorithmic bytes: 32 * 2^nl (uncontrolled), 16 * 2^nl (controlled pairing),
// 8..16 * 2^nl (diagonal) -- SURVEY.md section 8(d).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "gate_math.cuh"
#include "kernels.h"

namespace qcs {

namespace {

__device__ __forceinline__ uint64_t insert_zero(uint64_t v, int pos) {
  const uint64_t low = v & ((1ull << pos) - 1ull);
  return ((v >> pos) << (pos + 1)) | low;
}

__global__ void __launch_bounds__(256)
simple_pair_kernel(double2 *__restrict__ state, const __grid_constant__ DGate g,
                   int local_control, uint64_t n_items) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_items;
       item += stride) {
    uint64_t i0;
    if (local_control) {
      const int lo = g.tpos < g.cpos ? g.tpos : g.cpos;
      const int hi = g.tpos < g.cpos ? g.cpos : g.tpos;
      i0 = insert_zero(insert_zero(item, lo), hi) | (1ull << g.cpos);
    } else {
      i0 = insert_zero(item, g.tpos);
    }
    const uint64_t i1 = i0 | (1ull << g.tpos);
    double2 v0 = state[i0];
    double2 v1 = state[i1];
    pair_update(g, v0.x, v0.y, v1.x, v1.y);
    state[i0] = v0;
    if (!(g.flags & GF_ROW0_ONLY)) state[i1] = v1;
  }
}

__global__ void __launch_bounds__(256)
simple_diag_kernel(double2 *__restrict__ state, const __grid_constant__ DGate g,
                   uint64_t shard_base, uint64_t n_amps) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
    const uint64_t full = shard_base | i;
    if (g.cpos >= 0 && !bit_of(full, g.cpos)) continue;
    const bool t = bit_of(full, g.tpos);
    if (t ? (g.flags & GF_D1_IDENT) : (g.flags & GF_D0_IDENT)) continue;
    const double2 v = state[i];
    const cplx n = cmul(t ? g.m[6] : g.m[0], t ? g.m[7] : g.m[1], v.x, v.y);
    state[i] = make_double2(n.r, n.i);
  }
}

// Swaps two LOCAL index bits a < b of the whole shard in place: amplitude (..a=1,b=0..) trades
// places with (..a=0,b=1..).  Used to restore the identity qubit layout after position swaps.
__global__ void __launch_bounds__(256)
swap_local_bits_kernel(double2 *__restrict__ state, int a, int b, uint64_t n_items) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_items;
       item += stride) {
    const uint64_t base = insert_zero(insert_zero(item, a), b);
    const uint64_t i = base | (1ull << a);
    const uint64_t j = base | (1ull << b);
    const double2 vi = state[i];
    const double2 vj = state[j];
    state[i] = vj;
    state[j] = vi;
  }
}

// Issue-rate probe: 8 independent dependency chains of mul.rn.f64 / add.rn.f64 per thread, no
// memory traffic.  bench.py reports the measured rate as the FP64 ceiling of the fused pass, whose
// arithmetic may not be contracted into FMAs (bit-exactness) and therefore cannot use the
// datasheet's FMA-counted FP64 figure.
__global__ void __launch_bounds__(256)
fp64_probe_kernel(double *__restrict__ out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
  for (int i = 0; i < iters; i++) {
    x0 = __dadd_rn(__dmul_rn(x0, a), b); x1 = __dadd_rn(__dmul_rn(x1, a), b);
    x2 = __dadd_rn(__dmul_rn(x2, a), b); x3 = __dadd_rn(__dmul_rn(x3, a), b);
    x4 = __dadd_rn(__dmul_rn(x4, a), b); x5 = __dadd_rn(__dmul_rn(x5, a), b);
    x6 = __dadd_rn(__dmul_rn(x6, a), b); x7 = __dadd_rn(__dmul_rn(x7, a), b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace

cudaError_t launch_fp64_probe(double *out, int blocks, int iters, cudaStream_t stream) {
  fp64_probe_kernel<<<blocks, 256, 0, stream>>>(out, iters, 0.9999999, 1e-9);
  return cudaGetLastError();
}

cudaError_t launch_swap_local_bits(double2 *state, int n_local, int a, int b, cudaStream_t stream) {
  if (a == b) return cudaSuccess;
  if (a > b) { int t = a; a = b; b = t; }
  const uint64_t n_items = (1ull << n_local) >> 2;
  const unsigned blocks = (unsigned)((n_items + 255) / 256 > (unsigned)device_sm_count() * 16 ? (unsigned)device_sm_count() * 16 : (n_items + 255) / 256);
  swap_local_bits_kernel<<<blocks ? blocks : 1, 256, 0, stream>>>(state, a, b, n_items ? n_items : 1);
  return cudaGetLastError();
}

cudaError_t launch_simple_gate(double2 *state, const DGate &g, int n_local, uint64_t shard_base,
                               cudaStream_t stream) {
  if (g.kind == GK_NOP) return cudaSuccess;
  const uint64_t n_amps = 1ull << n_local;
  if (g.kind == GK_DIAG) {
    const unsigned blocks = (unsigned)((n_amps + 255) / 256 > (unsigned)device_sm_count() * 16 ? (unsigned)device_sm_count() * 16 : (n_amps + 255) / 256);
    simple_diag_kernel<<<blocks, 256, 0, stream>>>(state, g, shard_base, n_amps);
    return cudaGetLastError();
  }
  // pairing gate: the target must be local (the engine remaps first otherwise)
  int local_control = 0;
  uint64_t n_items = n_amps >> 1;
  if (g.cpos >= 0) {
    if (g.cpos >= n_local) {
      if (!((shard_base >> g.cpos) & 1ull)) return cudaSuccess;  // control bit is 0 on this rank
    } else {
      local_control = 1;
      n_items = n_amps >> 2;
    }
  }
  if (n_items == 0) n_items = 1;  // 1-qubit shard with a local control cannot happen; guard anyway
  const unsigned blocks = (unsigned)((n_items + 255) / 256 > (unsigned)device_sm_count() * 16 ? (unsigned)device_sm_count() * 16 : (n_items + 255) / 256);
  simple_pair_kernel<<<blocks, 256, 0, stream>>>(state, g, local_control, n_items);
  return cudaGetLastError();
}

}  // namespace qcs
