/*
 * oracle/qcs_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's state-vector arithmetic for the
 * gate-application / measurement / sampling hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * use it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_pinning.py checks this restatement
 *   (1) against the real reference compiled from /root/reference into
 *       oracle/_ref/libqcsref_{seq,corrected}.so (bit-for-bit, every amplitude,
 *       plus scratch buffer, measurement outcomes and shot histograms), and
 *   (2) against the golden vectors committed under tests/golden/ (generated
 *       from the real reference by tests/golden/make_golden.py) and the
 *       known-answer anchors recorded in SURVEY.md section 8(c).
 *
 * Two semantics:
 *   ORC_SEM_REFERENCE  bug-compatible with the unmodified sequential mode
 *                      (defects D1-D3)
 *   ORC_SEM_CORRECTED  D1-D3 repaired (matches oracle/make_corrected.py)
 */
#ifndef QCS_ORACLE_H
#define QCS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_SEM_REFERENCE = 0, ORC_SEM_CORRECTED = 1 };

typedef struct orc_state orc_state;

orc_state *orc_create(int n_qubits, int semantics);
void orc_destroy(orc_state *s);
int orc_num_qubits(const orc_state *s);
long orc_size(const orc_state *s);
/* Raw interleaved (re,im) views -- `live` is the reference's state->vector,
 * `scratch` its state->scratch_vector. */
double *orc_live(orc_state *s);
double *orc_scratch(orc_state *s);

/* m = row-major {re,im} x 4 = U00,U01,U10,U11 (src/q_gates.c:140-141). */
int orc_apply_1q(orc_state *s, const double m[8], int target);
int orc_apply_c1q(orc_state *s, const double m[8], int control, int target);
int orc_phase_flip(orc_state *s, long index);
void orc_diffusion(orc_state *s);
void orc_normalize(orc_state *s);

double orc_prob0(const orc_state *s, int qubit);
/* u = rand()/(double)RAND_MAX drawn by the caller (src/qcs.c:261). */
int orc_measure(orc_state *s, int qubit, double u);
double orc_probability(const orc_state *s, long index);
long orc_argmax(const orc_state *s);
/* results must hold 2^n ints; zeroed here (src/qcs.c:593). */
void orc_sample(const orc_state *s, const double *u, int shots, int *results);

/* Gate constructors (host libm, src/q_gates.c:362-491). */
void orc_gate_x(double m[8]);
void orc_gate_y(double m[8]);
void orc_gate_z(double m[8]);
void orc_gate_h(double m[8]);
void orc_gate_p(double m[8], double angle);
void orc_gate_rx(double m[8], double angle);
void orc_gate_ry(double m[8], double angle);
void orc_gate_rz(double m[8], double angle);

/* Algorithm drivers = gate sequences of src/qcs.c:402-457, 521-535, 614-644. */
int orc_grover_iterations(int n_qubits);
void orc_grover(orc_state *s, long solution);
void orc_qft(orc_state *s);
void orc_bv(orc_state *s, int hidden);
void orc_ghz(orc_state *s);

#ifdef __cplusplus
}
#endif
#endif
