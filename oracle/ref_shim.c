/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Builds the UNMODIFIED reference (SegusFaultise/qcs) into a shared object by
 * including its translation units from where they lie under $(QCS_REF_SRC)
 * (normally /root/reference/src, passed as -I on the command line by
 * oracle/Makefile).  Nothing from the reference is copied into this repo; the
 * only code in this file is a handful of accessors that expose the reference's
 * private state vector so that tests can compare amplitudes bit-for-bit.
 *
 * One translation unit == the reference's "single-header bundle" usage
 * (scripts/bundle.py), which is the fast way users build it (c_mul / c_add
 * inline).  gpu_complex.c (needs CL/cl.h) and main.c are left out, exactly as
 * SURVEY.md section 8(c) describes.
 *
 * Compiled with -std=c89 like the reference Makefile:8, which also turns FMA
 * contraction off (ISO mode) -- the canonical oracle arithmetic.
 */
#ifndef QCS_REF_MULTI_TU
#include "complex.c"
#include "q_matrix.c"
#include "q_state.c"
#include "q_utils.c"
#include "q_gates.c"
#include "thread_pools.c"
#endif
/* QCS_REF_MULTI_TU (oracle/Makefile, libqcsref_seqlib.so): the layout of the reference's own
 * Makefile instead -- one object per source file, so c_mul / c_add (src/complex.c) are CALLS from
 * the gate loops of src/q_gates.c, not inlined.  Only qcs.c is included here (the accessors below
 * need its private struct t_q_circuit); the other files are compiled on their own and linked in.
 * A CPU baseline only (~20x slower than the single-header build). */
#include "qcs.c"

/* qc_cphase is exported by the reference (src/qcs.c:431) but not declared in
 * its header; tests call it through this shared object directly. */

long qcsref_state_size(t_q_circuit *c) { return c->state->size; }

/* Copies the live amplitude buffer (2*size doubles, interleaved re/im). */
void qcsref_copy_state(t_q_circuit *c, double *out) {
  memcpy(out, c->state->vector,
         (size_t)c->state->size * sizeof(struct t_complex));
}

/* Copies the scratch buffer -- observable through q_apply_phase_flip
 * (src/q_gates.c:311-316), see SURVEY.md appendix A.4. */
void qcsref_copy_scratch(t_q_circuit *c, double *out) {
  memcpy(out, c->state->scratch_vector,
         (size_t)c->state->size * sizeof(struct t_complex));
}

/* Overwrites the live amplitudes (for tests that start from a random state). */
void qcsref_load_state(t_q_circuit *c, const double *in) {
  memcpy(c->state->vector, in,
         (size_t)c->state->size * sizeof(struct t_complex));
}

void qcsref_load_scratch(t_q_circuit *c, const double *in) {
  memcpy(c->state->scratch_vector, in,
         (size_t)c->state->size * sizeof(struct t_complex));
}

/* Direct access to the internal (non-API) state operations used by Grover. */
void qcsref_phase_flip(t_q_circuit *c, int index) {
  q_apply_phase_flip(c->state, index);
}
void qcsref_diffusion(t_q_circuit *c) { q_apply_diffusion(c->state); }
void qcsref_normalize(t_q_circuit *c) { q_state_normalize(c->state); }

/* Arbitrary 2x2 gate through the reference's own apply functions. */
void qcsref_apply_1q(t_q_circuit *c, const double *m, int target) {
  struct t_q_matrix *g = q_matrix_init(2, 2);
  int k;
  for (k = 0; k < 4; k++) {
    g->data[k].number_real = m[2 * k];
    g->data[k].number_imaginary = m[2 * k + 1];
  }
  q_apply_1q_gate(c->state, g, target);
  q_matrix_free(g);
}
void qcsref_apply_c1q(t_q_circuit *c, const double *m, int control,
                      int target) {
  struct t_q_matrix *g = q_matrix_init(2, 2);
  int k;
  for (k = 0; k < 4; k++) {
    g->data[k].number_real = m[2 * k];
    g->data[k].number_imaginary = m[2 * k + 1];
  }
  q_apply_2q_gate(c->state, g, control, target);
  q_matrix_free(g);
}

/* History inspection (src/qcs.c:13-23) for bookkeeping parity. */
int qcsref_history_size(t_q_circuit *c) { return c->history_size; }
const char *qcsref_history_name(t_q_circuit *c, int i) {
  return c->gate_history[i];
}
int qcsref_history_target(t_q_circuit *c, int i) { return c->target_qubits[i]; }
int qcsref_history_control(t_q_circuit *c, int i) {
  return c->control_qubits[i];
}
double qcsref_history_param(t_q_circuit *c, int i) { return c->parameters[i]; }
