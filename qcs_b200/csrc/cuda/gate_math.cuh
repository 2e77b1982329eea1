// gate_math.cuh -- the scalar arithmetic every gate kernel shares.
//
// Operation order is the reference's (c_mul then c_add, reference
// src/complex.c:23-57, used at src/q_gates.c:140-141).  All products and sums
// go through __dmul_rn/__dadd_rn/__dsub_rn, which the compiler never contracts
// into an FMA, so each binary operation is rounded on its own exactly like the
// oracle built with -ffp-contract=off.  Specialised classes drop only terms
// that are exact zeros or exact copies (common.h, GateKind).
#pragma once
#include <cuda_runtime.h>

#include "common.h"

namespace qcs {

struct cplx {
  double r, i;
};

// c_mul(g, v): re = g.r*v.r - g.i*v.i ; im = g.r*v.i + g.i*v.r
__device__ __forceinline__ cplx cmul(double gr, double gi, double vr, double vi) {
  cplx o;
  o.r = __dsub_rn(__dmul_rn(gr, vr), __dmul_rn(gi, vi));
  o.i = __dadd_rn(__dmul_rn(gr, vi), __dmul_rn(gi, vr));
  return o;
}

// row = c_add(c_mul(ga, v0), c_mul(gb, v1))
__device__ __forceinline__ cplx row_generic(double gar, double gai, double gbr, double gbi,
                                            double v0r, double v0i, double v1r, double v1i) {
  cplx a = cmul(gar, gai, v0r, v0i);
  cplx b = cmul(gbr, gbi, v1r, v1i);
  cplx o;
  o.r = __dadd_rn(a.r, b.r);
  o.i = __dadd_rn(a.i, b.i);
  return o;
}

// Same with purely real matrix entries: the g.i*v terms are exact zeros.
__device__ __forceinline__ cplx row_real(double ga, double gb, double v0r, double v0i,
                                         double v1r, double v1i) {
  cplx o;
  o.r = __dadd_rn(__dmul_rn(ga, v0r), __dmul_rn(gb, v1r));
  o.i = __dadd_rn(__dmul_rn(ga, v0i), __dmul_rn(gb, v1i));
  return o;
}

// Applies one pairing gate of compile-time class KIND to the pair
// (v0 = target bit 0, v1 = target bit 1); m = row-major {re,im} x 4.
// `row0_only` is the reference-semantics controlled update (n1 = v1).
template <int KIND>
__device__ __forceinline__ void pair_update_k(const double *__restrict__ m, bool row0_only,
                                              double &v0r, double &v0i, double &v1r,
                                              double &v1i) {
  if (KIND == GK_PAIR_HSYM) {
    const double h0 = m[0], h1 = m[2];
    double p0r = __dmul_rn(h0, v0r), p0i = __dmul_rn(h0, v0i);
    double p1r = __dmul_rn(h1, v1r), p1i = __dmul_rn(h1, v1i);
    v0r = __dadd_rn(p0r, p1r);
    v0i = __dadd_rn(p0i, p1i);
    if (!row0_only) {
      v1r = __dsub_rn(p0r, p1r);
      v1i = __dsub_rn(p0i, p1i);
    }
  } else if (KIND == GK_PAIR_REAL) {
    cplx n0 = row_real(m[0], m[2], v0r, v0i, v1r, v1i);
    if (!row0_only) {
      cplx n1 = row_real(m[4], m[6], v0r, v0i, v1r, v1i);
      v1r = n1.r;
      v1i = n1.i;
    }
    v0r = n0.r;
    v0i = n0.i;
  } else if (KIND == GK_PAIR_SWAP) {
    double tr = v0r, ti = v0i;
    v0r = v1r;
    v0i = v1i;
    if (!row0_only) {
      v1r = tr;
      v1i = ti;
    }
  } else {
    cplx n0 = row_generic(m[0], m[1], m[2], m[3], v0r, v0i, v1r, v1i);
    if (!row0_only) {
      cplx n1 = row_generic(m[4], m[5], m[6], m[7], v0r, v0i, v1r, v1i);
      v1r = n1.r;
      v1i = n1.i;
    }
    v0r = n0.r;
    v0i = n0.i;
  }
}

// Run-time class dispatch (per-gate kernels).
__device__ __forceinline__ void pair_update(const DGate &g, double &v0r, double &v0i,
                                            double &v1r, double &v1i) {
  const bool row0 = (g.flags & GF_ROW0_ONLY) != 0;
  switch (g.kind) {
    case GK_PAIR_HSYM: pair_update_k<GK_PAIR_HSYM>(g.m, row0, v0r, v0i, v1r, v1i); break;
    case GK_PAIR_REAL: pair_update_k<GK_PAIR_REAL>(g.m, row0, v0r, v0i, v1r, v1i); break;
    case GK_PAIR_SWAP: pair_update_k<GK_PAIR_SWAP>(g.m, row0, v0r, v0i, v1r, v1i); break;
    default: pair_update_k<GK_PAIR_GENERIC>(g.m, row0, v0r, v0i, v1r, v1i); break;
  }
}

// Bit `pos` of a full (shard_base | local) basis index.
__device__ __forceinline__ bool bit_of(uint64_t idx, int pos) { return (idx >> pos) & 1ull; }

}  // namespace qcs
