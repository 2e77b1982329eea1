/*
 * oracle/qcs_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's sequential arithmetic.  See
 * qcs_oracle.h for scope and pinning status.  Build with -ffp-contract=off
 * (oracle/Makefile does): every binary operation is rounded on its own, as in
 * the reference built with -std=c89 (reference Makefile:8).
 *
 * Conventions: qubit q <-> bit q of the basis index;
 * amplitudes are interleaved doubles {re, im}.
 */
#include "qcs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct orc_state {
  int n;
  long size;
  int sem;
  double *live;    /* reference: state->vector          (src/internal.h:46) */
  double *scratch; /* reference: state->scratch_vector  (src/internal.h:47) */
};

/* ---- scalar complex arithmetic: op order of src/complex.c:23-57 ---------- */

/* c_mul(g, v): re = g.r*v.r - g.i*v.i ; im = g.r*v.i + g.i*v.r  (complex.c:50-54) */
static void cx_mul(const double *g, const double *v, double *o) {
  double re = (g[0] * v[0]) - (g[1] * v[1]);
  double im = (g[0] * v[1]) + (g[1] * v[0]);
  o[0] = re;
  o[1] = im;
}

/* row(ga, gb; v0, v1) = c_add(c_mul(ga, v0), c_mul(gb, v1))  (q_gates.c:140-141) */
static void cx_row(const double *ga, const double *gb, const double *v0,
                   const double *v1, double *o) {
  double a[2], b[2];
  cx_mul(ga, v0, a);
  cx_mul(gb, v1, b);
  o[0] = a[0] + b[0];
  o[1] = a[1] + b[1];
}

static void swap_buffers(orc_state *s) {
  double *t = s->live;
  s->live = s->scratch;
  s->scratch = t;
}

/* ---- storage: q_state_init / q_state_free (src/q_state.c:32-115) --------- */

orc_state *orc_create(int n_qubits, int semantics) {
  orc_state *s;
  if (n_qubits <= 0 || n_qubits > 30)
    return NULL;
  s = (orc_state *)malloc(sizeof(*s));
  if (!s)
    return NULL;
  s->n = n_qubits;
  s->size = 1L << n_qubits;
  s->sem = semantics;
  s->live = (double *)calloc((size_t)s->size * 2, sizeof(double));
  s->scratch = (double *)calloc((size_t)s->size * 2, sizeof(double));
  if (!s->live || !s->scratch) {
    free(s->live);
    free(s->scratch);
    free(s);
    return NULL;
  }
  s->live[0] = 1.0; /* vector[0] = 1 (q_state.c:99) */
  return s;
}

void orc_destroy(orc_state *s) {
  if (!s)
    return;
  free(s->live);
  free(s->scratch);
  free(s);
}

int orc_num_qubits(const orc_state *s) { return s->n; }
long orc_size(const orc_state *s) { return s->size; }
double *orc_live(orc_state *s) { return s->live; }
double *orc_scratch(orc_state *s) { return s->scratch; }

/* ---- 1-qubit gate: q_apply_1q_gate sequential arm (q_gates.c:130-148) ---- *
 * Reference: memcpy(scratch, vector); overwrite every pair in scratch; swap.
 * Every index belongs to exactly one pair, so the memcpy is unobservable:
 * net effect new_live = U applied, new_scratch = complete pre-gate state.    */
int orc_apply_1q(orc_state *s, const double m[8], int target) {
  long stride, lo, blk;
  if (target < 0 || target >= s->n)
    return 1; /* reference prints and leaves the state untouched (:32-36) */
  stride = 1L << target;
  for (blk = 0; blk < s->size; blk += 2 * stride) {
    for (lo = blk; lo < blk + stride; lo++) {
      const double *v0 = s->live + 2 * lo;
      const double *v1 = s->live + 2 * (lo + stride);
      cx_row(m + 0, m + 2, v0, v1, s->scratch + 2 * lo);
      cx_row(m + 4, m + 6, v0, v1, s->scratch + 2 * (lo + stride));
    }
  }
  swap_buffers(s);
  return 0;
}

/* ---- controlled gate: q_apply_2q_gate sequential arm (q_gates.c:276-297) -- *
 * REFERENCE semantics (defect D1): for control=1,target=0 index i the loop
 * writes scratch[i] = row0 and scratch[i|t] = row1, but the later visit of
 * i|t (control=1,target=1) takes the else-arm and copies vector[i|t] back
 * over it.  Net: n0 = g0*v0 + g1*v1, n1 = v1 (unchanged).
 * CORRECTED semantics: n1 = g2*v0 + g3*v1.                                   */
int orc_apply_c1q(orc_state *s, const double m[8], int control, int target) {
  long cbit, tbit, i;
  if (control < 0 || target < 0 || control >= s->n || target >= s->n ||
      control == target)
    return 1; /* q_gates.c:165-170 */
  cbit = 1L << control;
  tbit = 1L << target;
  memcpy(s->scratch, s->live, (size_t)s->size * 2 * sizeof(double));
  for (i = 0; i < s->size; i++) {
    if ((i & cbit) && !(i & tbit)) {
      const double *v0 = s->live + 2 * i;
      const double *v1 = s->live + 2 * (i | tbit);
      cx_row(m + 0, m + 2, v0, v1, s->scratch + 2 * i);
      if (s->sem == ORC_SEM_CORRECTED)
        cx_row(m + 4, m + 6, v0, v1, s->scratch + 2 * (i | tbit));
    }
  }
  swap_buffers(s);
  return 0;
}

/* ---- phase flip: q_apply_phase_flip (q_gates.c:305-317) ------------------- *
 * REFERENCE (defect D2): swap buffers first, then live[idx] = -scratch[idx];
 * every other element of the new live buffer is whatever scratch held.       */
int orc_phase_flip(orc_state *s, long index) {
  if (index < 0 || index >= s->size)
    return 1;
  if (s->sem == ORC_SEM_REFERENCE) {
    swap_buffers(s);
    s->live[2 * index] = -s->scratch[2 * index];
    s->live[2 * index + 1] = -s->scratch[2 * index + 1];
  } else {
    s->live[2 * index] = -s->live[2 * index];
    s->live[2 * index + 1] = -s->live[2 * index + 1];
  }
  return 0;
}

/* ---- diffusion: q_apply_diffusion (q_gates.c:323-356) --------------------- *
 * Sum is a sequential ascending c_add.  REFERENCE (defect D3) divides the sum
 * by its magnitude when that exceeds 1e-10; CORRECTED divides by N.
 * Then swap, live[i] = 2*mean - scratch[i] componentwise.                    */
void orc_diffusion(orc_state *s) {
  double sr = 0.0, si = 0.0, tr, ti;
  long i;
  for (i = 0; i < s->size; i++) {
    sr = sr + s->live[2 * i];
    si = si + s->live[2 * i + 1];
  }
  if (s->sem == ORC_SEM_REFERENCE) {
    double mag = sqrt((sr * sr) + (si * si)); /* c_magnitude, complex.c:95 */
    if (mag > 1e-10) {
      sr /= mag;
      si /= mag;
    }
  } else {
    sr /= (double)s->size;
    si /= (double)s->size;
  }
  tr = 2.0 * sr;
  ti = 2.0 * si;
  swap_buffers(s);
  for (i = 0; i < s->size; i++) {
    s->live[2 * i] = tr - s->scratch[2 * i];
    s->live[2 * i + 1] = ti - s->scratch[2 * i + 1];
  }
}

/* ---- normalise: q_state_normalize sequential arm (q_utils.c:103-117) ------ */
void orc_normalize(orc_state *s) {
  double total = 0.0;
  long i;
  for (i = 0; i < s->size; i++)
    total += (s->live[2 * i] * s->live[2 * i]) +
             (s->live[2 * i + 1] * s->live[2 * i + 1]);
  if (total > 1e-12 && total != 1.0) {
    double inv = 1.0 / sqrt(total);
    for (i = 0; i < s->size; i++) {
      s->live[2 * i] *= inv;
      s->live[2 * i + 1] *= inv;
    }
  }
}

/* ---- measurement: qc_measure (qcs.c:237-284) ------------------------------ */
double orc_prob0(const orc_state *s, int qubit) {
  double p = 0.0;
  long i, bit = 1L << qubit;
  for (i = 0; i < s->size; i++)
    if (!(i & bit))
      p += (s->live[2 * i] * s->live[2 * i]) +
           (s->live[2 * i + 1] * s->live[2 * i + 1]);
  return p;
}

int orc_measure(orc_state *s, int qubit, double u) {
  long i, bit;
  int outcome;
  if (qubit < 0 || qubit >= s->n)
    return 0; /* qcs.c:245-248 */
  bit = 1L << qubit;
  outcome = (u <= orc_prob0(s, qubit)) ? 0 : 1; /* `<=`, qcs.c:263 */
  for (i = 0; i < s->size; i++) {
    int b = (i & bit) != 0;
    if (b != outcome) {
      s->live[2 * i] = 0.0;
      s->live[2 * i + 1] = 0.0;
    }
  }
  orc_normalize(s);
  return outcome;
}

/* ---- state access: qc_get_probability / qc_find_most_likely_state --------- *
 * (qcs.c:391-395, 464-478): c_norm_sq; first strict maximum, start 0.0.      */
double orc_probability(const orc_state *s, long index) {
  if (index < 0 || index >= s->size)
    return 0.0;
  return (s->live[2 * index] * s->live[2 * index]) +
         (s->live[2 * index + 1] * s->live[2 * index + 1]);
}

long orc_argmax(const orc_state *s) {
  long best = 0, i;
  double bestp = 0.0;
  for (i = 0; i < s->size; i++) {
    double p = orc_probability(s, i);
    if (p > bestp) {
      bestp = p;
      best = i;
    }
  }
  return best;
}

/* ---- sampling: qc_run_shots (qcs.c:575-607) ------------------------------- *
 * Per shot: running sum left to right, first i with u < cumulative (strict);
 * a shot with no such i is dropped.                                          */
void orc_sample(const orc_state *s, const double *u, int shots, int *results) {
  int k;
  long i;
  memset(results, 0, (size_t)s->size * sizeof(int));
  for (k = 0; k < shots; k++) {
    double cum = 0.0;
    for (i = 0; i < s->size; i++) {
      cum += orc_probability(s, i);
      if (u[k] < cum) {
        results[i]++;
        break;
      }
    }
  }
}

/* ---- gate constructors (q_gates.c:375-491); untouched entries are +0.0 ---- */
static void zero8(double m[8]) { memset(m, 0, 8 * sizeof(double)); }

void orc_gate_x(double m[8]) { zero8(m); m[2] = 1.0; m[4] = 1.0; }
void orc_gate_y(double m[8]) { zero8(m); m[3] = -1.0; m[5] = 1.0; }
void orc_gate_z(double m[8]) { zero8(m); m[0] = 1.0; m[6] = -1.0; }
void orc_gate_h(double m[8]) {
  double h = 1.0 / sqrt(2.0);
  zero8(m);
  m[0] = h; m[2] = h; m[4] = h; m[6] = -h;
}
void orc_gate_p(double m[8], double angle) {
  zero8(m);
  m[0] = 1.0;
  m[6] = cos(angle);
  m[7] = sin(angle);
}
void orc_gate_rx(double m[8], double angle) {
  double c = cos(angle / 2.0), sn = sin(angle / 2.0);
  zero8(m);
  m[0] = c; m[3] = -sn; m[5] = -sn; m[6] = c;
}
void orc_gate_ry(double m[8], double angle) {
  double c = cos(angle / 2.0), sn = sin(angle / 2.0);
  zero8(m);
  m[0] = c; m[2] = -sn; m[4] = sn; m[6] = c;
}
void orc_gate_rz(double m[8], double angle) {
  double c = cos(angle / 2.0), sn = sin(angle / 2.0);
  zero8(m);
  m[0] = c; m[1] = -sn; m[6] = c; m[7] = sn;
}

/* ---- algorithm drivers ------------------------------------------------------ */

/* q_grover_iterations (q_utils.c:126-134): floor((pi/4) * sqrt(2^n)). */
int orc_grover_iterations(int n_qubits) {
  double big_n = (double)(1 << n_qubits);
  return (int)floor((3.14159265358979323846 / 4.0) * sqrt(big_n));
}

/* qc_grover_search (qcs.c:402-419). */
void orc_grover(orc_state *s, long solution) {
  double h[8];
  int q, it, iters = orc_grover_iterations(s->n);
  orc_gate_h(h);
  for (q = 0; q < s->n; q++)
    orc_apply_1q(s, h, q);
  for (it = 0; it < iters; it++) {
    orc_phase_flip(s, solution);
    orc_diffusion(s);
  }
}

/* qc_quantum_fourier_transform (qcs.c:443-457): no final swaps. */
void orc_qft(orc_state *s) {
  double h[8], cp[8];
  int i, j;
  orc_gate_h(h);
  for (i = 0; i < s->n; i++) {
    orc_apply_1q(s, h, i);
    for (j = i + 1; j < s->n; j++) {
      orc_gate_p(cp, 3.14159265358979323846 / (double)(1 << (j - i)));
      orc_apply_c1q(s, cp, j, i);
    }
  }
}

/* qc_bernstein_vazirani (qcs.c:614-644): ancilla = top qubit. */
void orc_bv(orc_state *s, int hidden) {
  double x[8], h[8];
  int i, n = s->n - 1;
  if (n <= 0)
    return;
  orc_gate_x(x);
  orc_gate_h(h);
  orc_apply_1q(s, x, n);
  orc_apply_1q(s, h, n);
  for (i = 0; i < n; i++)
    orc_apply_1q(s, h, i);
  for (i = 0; i < n; i++)
    if ((hidden >> i) & 1)
      orc_apply_c1q(s, x, i, n);
  for (i = 0; i < n; i++)
    orc_apply_1q(s, h, i);
}

/* qc_ghz_state (qcs.c:521-535). */
void orc_ghz(orc_state *s) {
  double x[8], h[8];
  int i;
  if (s->n < 2)
    return;
  orc_gate_x(x);
  orc_gate_h(h);
  orc_apply_1q(s, h, 0);
  for (i = 0; i < s->n - 1; i++)
    orc_apply_c1q(s, x, i, i + 1);
}
