/*
 * qcs_host.c -- the C89 host layer of the QCS_GPU_CUDA mode (libqcs.so).
 *
 * Implements the public API of include/qcs.h (reference include/qcs.h:28-74;
 * behaviour of reference src/qcs.c) on top of the extern "C" ABI of
 * include/qcs_cuda.h.  This layer owns only bookkeeping: the gate history
 * (names, targets, controls, parameters -- what qc_get_num_gates,
 * qc_print_circuit and qc_optimize look at), the 2x2 matrices (built with the
 * host libm exactly like reference src/q_gates.c:375-491 so the entries are
 * bit-identical), and the libc rand() draws (one per qc_measure, one per shot,
 * in the reference's order, so srand() in user code behaves identically).
 * Every touch of an amplitude goes to the device library; there is no CPU
 * arithmetic on the state here and no fallback.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/qcs.h"
#include "../../../include/qcs_cuda.h"

#define QCS_PI 3.14159265358979323846

struct hist_entry {
  char *name;
  int target;
  int control;
  double param;
};

struct t_q_circuit {
  int num_qubits;
  int num_gates;
  qcs_cuda_engine *engine;
  struct hist_entry *hist;
  int hist_len;
  int hist_cap;
};

static void report(const char *where, int rc) {
  if (rc != QCS_CUDA_OK)
    fprintf(stderr, "Error: %s (%s)\n", qcs_cuda_last_error(), where);
}

/* ---- lifetime: qc_create / qc_destroy (reference src/qcs.c:30-100) -------- */

t_q_circuit *qc_create(int num_qubits) {
  t_q_circuit *c = (t_q_circuit *)malloc(sizeof(*c));
  int rc;
  if (!c)
    return NULL;
  c->num_qubits = num_qubits;
  c->num_gates = 0;
  c->engine = NULL;
  c->hist_len = 0;
  c->hist_cap = 100;
  c->hist = (struct hist_entry *)malloc((size_t)c->hist_cap * sizeof(*c->hist));
  rc = qcs_cuda_state_create(&c->engine, num_qubits);
  if (rc != QCS_CUDA_OK || !c->hist) {
    /* The reference documents NULL as the failure result (src/qcs.c:28); with
     * no device state there is nothing a circuit could do. */
    report("qc_create", rc);
    free(c->hist);
    free(c);
    return NULL;
  }
  return c;
}

void qc_destroy(t_q_circuit *circuit) {
  int i;
  if (!circuit)
    return;
  qcs_cuda_state_destroy(circuit->engine);
  for (i = 0; i < circuit->hist_len; i++)
    free(circuit->hist[i].name);
  free(circuit->hist);
  free(circuit);
}

qcs_cuda_engine *qc_cuda_engine(t_q_circuit *circuit) {
  return circuit ? circuit->engine : NULL;
}

/* ---- history: qc_add_gate (reference src/qcs.c:132-152) -------------------- */

void qc_add_gate(t_q_circuit *circuit, const char *gate_name, int target,
                 int control, double param) {
  struct hist_entry *h;
  size_t len;
  if (circuit->hist_len >= circuit->hist_cap) {
    circuit->hist_cap *= 2;
    circuit->hist = (struct hist_entry *)realloc(
        circuit->hist, (size_t)circuit->hist_cap * sizeof(*circuit->hist));
  }
  h = &circuit->hist[circuit->hist_len++];
  h->name = NULL;
  if (gate_name) {
    len = strlen(gate_name) + 1;
    h->name = (char *)malloc(len);
    if (h->name)
      memcpy(h->name, gate_name, len);
  }
  h->target = target;
  h->control = control;
  h->param = param;
  circuit->num_gates++;
}

/* ---- gates ----------------------------------------------------------------- *
 * As in the reference (src/qcs.c:159-164), the history entry is appended even
 * when the gate was rejected as invalid; the state is then left untouched and
 * the reference's message goes to stderr.                                      */

static void submit_1q(t_q_circuit *c, const double m[8], int qubit,
                      const char *name, double param) {
  if (qcs_cuda_apply_1q(c->engine, m, qubit) != QCS_CUDA_OK)
    fprintf(stderr, "Error: %s\n", qcs_cuda_last_error());
  qc_add_gate(c, name, qubit, -1, param);
}

static void submit_c1q(t_q_circuit *c, const double m[8], int control,
                       int target, const char *name, double param) {
  if (qcs_cuda_apply_c1q(c->engine, m, control, target) != QCS_CUDA_OK)
    fprintf(stderr, "Error: %s\n", qcs_cuda_last_error());
  qc_add_gate(c, name, target, control, param);
}

static void mat_clear(double m[8]) {
  int k;
  for (k = 0; k < 8; k++)
    m[k] = 0.0;
}

/* entries as reference src/q_gates.c:375-491; unset components are +0.0 */
static void mat_x(double m[8]) { mat_clear(m); m[2] = 1.0; m[4] = 1.0; }
static void mat_p(double m[8], double angle) {
  mat_clear(m);
  m[0] = 1.0;
  m[6] = cos(angle);
  m[7] = sin(angle);
}

void qc_h(t_q_circuit *circuit, int qubit) {
  double m[8];
  double h = 1.0 / sqrt(2.0);
  mat_clear(m);
  m[0] = h; m[2] = h; m[4] = h; m[6] = -h;
  submit_1q(circuit, m, qubit, "H", 0.0);
}

void qc_x(t_q_circuit *circuit, int qubit) {
  double m[8];
  mat_x(m);
  submit_1q(circuit, m, qubit, "X", 0.0);
}

void qc_y(t_q_circuit *circuit, int qubit) {
  double m[8];
  mat_clear(m);
  m[3] = -1.0;
  m[5] = 1.0;
  submit_1q(circuit, m, qubit, "Y", 0.0);
}

void qc_z(t_q_circuit *circuit, int qubit) {
  double m[8];
  mat_clear(m);
  m[0] = 1.0;
  m[6] = -1.0;
  submit_1q(circuit, m, qubit, "Z", 0.0);
}

void qc_cnot(t_q_circuit *circuit, int control, int target) {
  double m[8];
  mat_x(m);
  submit_c1q(circuit, m, control, target, "CNOT", 0.0);
}

void qc_phase(t_q_circuit *circuit, int qubit, double angle) {
  double m[8];
  mat_p(m, angle);
  submit_1q(circuit, m, qubit, "P", angle);
}

void qc_cphase(t_q_circuit *circuit, int control, int target, double angle) {
  double m[8];
  mat_p(m, angle);
  submit_c1q(circuit, m, control, target, "CPHASE", angle);
}

void qc_rx(t_q_circuit *circuit, int qubit, double angle) {
  double m[8];
  double ch = cos(angle / 2.0), sh = sin(angle / 2.0);
  mat_clear(m);
  m[0] = ch; m[3] = -sh; m[5] = -sh; m[6] = ch;
  submit_1q(circuit, m, qubit, "RX", angle);
}

void qc_ry(t_q_circuit *circuit, int qubit, double angle) {
  double m[8];
  double ch = cos(angle / 2.0), sh = sin(angle / 2.0);
  mat_clear(m);
  m[0] = ch; m[2] = -sh; m[4] = sh; m[6] = ch;
  submit_1q(circuit, m, qubit, "RY", angle);
}

void qc_rz(t_q_circuit *circuit, int qubit, double angle) {
  double m[8];
  double ch = cos(angle / 2.0), sh = sin(angle / 2.0);
  mat_clear(m);
  m[0] = ch; m[1] = -sh; m[6] = ch; m[7] = sh;
  submit_1q(circuit, m, qubit, "RZ", angle);
}

void qc_barrier(t_q_circuit *circuit) {
  qc_add_gate(circuit, "BARRIER", -1, -1, 0.0);
}

/* ---- measurement (reference src/qcs.c:237-302, 550-567) -------------------- */

int qc_measure(t_q_circuit *circuit, int qubit) {
  double p0 = 0.0, r;
  int outcome, rc;
  if (circuit == NULL || circuit->engine == NULL || qubit < 0 ||
      qubit >= circuit->num_qubits)
    return 0;
  rc = qcs_cuda_prob0(circuit->engine, qubit, &p0);
  if (rc != QCS_CUDA_OK) {
    report("qc_measure", rc);
    return 0;
  }
  r = rand() / (double)RAND_MAX;
  outcome = (r <= p0) ? 0 : 1;
  report("qc_measure", qcs_cuda_collapse(circuit->engine, qubit, outcome));
  return outcome;
}

void qc_measure_all(t_q_circuit *circuit, int *results) {
  int q;
  if (circuit == NULL || results == NULL)
    return;
  for (q = 0; q < circuit->num_qubits; q++)
    results[q] = qc_measure(circuit, q);
  qc_add_gate(circuit, "MEASURE", -1, -1, 0.0);
}

void qc_reset(t_q_circuit *circuit, int qubit) {
  if (qc_measure(circuit, qubit) == 1)
    qc_x(circuit, qubit);
  qc_add_gate(circuit, "RESET", qubit, -1, 0.0);
}

void qc_run(t_q_circuit *circuit) {
  int *tmp = (int *)malloc((size_t)circuit->num_qubits * sizeof(int));
  if (tmp) {
    qc_measure_all(circuit, tmp);
    free(tmp);
  }
}

/* qc_run_shots (reference src/qcs.c:575-607).  The reference draws one rand()
 * per shot while scanning; nothing else consumes rand() in between, so drawing
 * the whole batch first yields the same stream.  The dense histogram is the
 * caller's (2^n ints, zeroed here as the reference does). */
void qc_run_shots(t_q_circuit *circuit, int shots, int *results) {
  double *u;
  long *idx;
  size_t n_states;
  int s, rc;
  if (!circuit || shots <= 0 || !results)
    return;
  n_states = (size_t)1 << circuit->num_qubits;
  u = (double *)malloc((size_t)shots * sizeof(double));
  idx = (long *)malloc((size_t)shots * sizeof(long));
  if (!u || !idx) {
    free(u);
    free(idx);
    return;
  }
  memset(results, 0, n_states * sizeof(int));
  for (s = 0; s < shots; s++)
    u[s] = rand() / (double)RAND_MAX;
  rc = qcs_cuda_sample(circuit->engine, u, shots, idx);
  report("qc_run_shots", rc);
  if (rc == QCS_CUDA_OK)
    for (s = 0; s < shots; s++)
      if (idx[s] >= 0)
        results[idx[s]]++;
  free(u);
  free(idx);
}

/* Sparse companion of qc_run_shots: the same draws, the same
 * decisions, but the outcome of shot s is written to indices[s] (-1: the reference drops that
 * shot) instead of a dense 2^n histogram -- at 32 qubits the histogram alone is 16 GiB that
 * the callee must zero.  Not part of include/qcs.h; declared in include/qcs_cuda.h. */
void qc_run_shots_sparse(t_q_circuit *circuit, int shots, long *indices) {
  double *u;
  int s;
  if (!circuit || shots <= 0 || !indices)
    return;
  u = (double *)malloc((size_t)shots * sizeof(double));
  if (!u)
    return;
  for (s = 0; s < shots; s++)
    u[s] = rand() / (double)RAND_MAX;
  report("qc_run_shots_sparse", qcs_cuda_sample(circuit->engine, u, shots, indices));
  free(u);
}

/* ---- state access (reference src/qcs.c:391-395, 464-478) -------------------- */

double qc_get_probability(t_q_circuit *circuit, int state) {
  double p = 0.0;
  if (state < 0 || (circuit->num_qubits < 31 && state >= (1 << circuit->num_qubits)))
    return 0.0;
  report("qc_get_probability", qcs_cuda_probability(circuit->engine, (long)state, &p));
  return p;
}

int qc_find_most_likely_state(t_q_circuit *circuit) {
  long idx = 0;
  report("qc_find_most_likely_state", qcs_cuda_argmax(circuit->engine, &idx));
  return (int)idx;
}

/* qc_print_state -> q_state_print (reference src/q_state.c:139-168): first 4
 * amplitudes (all if <= 8 states), the highlighted one, the last one. */
void qc_print_state(t_q_circuit *circuit, int solution_index) {
  long size = 1L << circuit->num_qubits;
  int max_print = size > 8 ? 4 : (int)size;
  double a[2];
  long i;
  printf("--- Quantum State (%d Qubits) ---\n", circuit->num_qubits);
  for (i = 0; i < max_print; i++) {
    a[0] = a[1] = 0.0;
    qcs_cuda_get_amplitude(circuit->engine, i, a);
    printf("|%ld>: %f + i%f%s\n", i, a[0], a[1],
           i == solution_index ? " <-- SOLUTION" : "");
  }
  if (solution_index >= max_print && solution_index < size - 1) {
    a[0] = a[1] = 0.0;
    qcs_cuda_get_amplitude(circuit->engine, (long)solution_index, a);
    printf("...\n");
    printf("|%d>: %f + i%f <-- SOLUTION\n", solution_index, a[0], a[1]);
  }
  if (size > max_print) {
    a[0] = a[1] = 0.0;
    qcs_cuda_get_amplitude(circuit->engine, size - 1, a);
    printf("...\n");
    printf("|%ld>: %f + i%f%s\n", size - 1, a[0], a[1],
           size - 1 == solution_index ? " <-- SOLUTION" : "");
  }
  printf("----------------------------------\n");
}

/* qc_print_circuit (reference src/qcs.c:308-374): same text, byte for byte. */
void qc_print_circuit(t_q_circuit *circuit) {
  int q, g, measured = 0;
  const struct hist_entry *h;
  printf("\n\xE2\x95\x94");
  for (g = 0; g < 42; g++) printf("\xE2\x95\x90");
  printf("\xE2\x95\x97\n");
  printf("\xE2\x95\x91              QUANTUM CIRCUIT             \xE2\x95\x91\n");
  printf("\xE2\x95\x9A");
  for (g = 0; g < 42; g++) printf("\xE2\x95\x90");
  printf("\xE2\x95\x9D\n");
  printf("Qubits: %d | Gates: %d\n\n", circuit->num_qubits, circuit->num_gates);

  for (g = 0; g < circuit->hist_len; g++)
    if (strcmp(circuit->hist[g].name, "MEASURE") == 0) {
      measured = 1;
      break;
    }

  for (q = 0; q < circuit->num_qubits; q++) {
    printf("q%-2d ", q);
    for (g = 0; g < circuit->hist_len; g++) {
      const char *cell = "\xE2\x94\x80\xE2\x94\x80\xE2\x94\x80"; /* --- */
      h = &circuit->hist[g];
      if (strcmp(h->name, "H") == 0 && h->target == q)
        cell = "\xE2\x94\x80H\xE2\x94\x80";
      else if (strcmp(h->name, "X") == 0 && h->target == q)
        cell = "\xE2\x94\x80X\xE2\x94\x80";
      else if (strcmp(h->name, "CNOT") == 0) {
        if (h->control == q)
          cell = "\xE2\x94\x80\xE2\x88\x99\xE2\x94\x80"; /* control dot */
        else if (h->target == q)
          cell = "\xE2\x94\x80\xE2\x8A\x95\xE2\x94\x80"; /* circled plus */
      }
      printf("%s", cell);
      if (g < circuit->hist_len - 1)
        printf("\xE2\x94\x80");
    }
    if (!measured)
      printf("\xE2\x94\x80M\xE2\x94\x80");
    printf("\xE2\x94\x80\xE2\x94\x90\n");
  }

  printf("\nGATE SEQUENCE: ");
  for (g = 0; g < circuit->hist_len; g++) {
    h = &circuit->hist[g];
    if (strcmp(h->name, "CNOT") == 0)
      printf("CNOT(%d,%d) ", h->control, h->target);
    else if (strcmp(h->name, "MEASURE") == 0)
      printf("MEASURE ");
    else
      printf("%s(%d) ", h->name, h->target);
  }
  printf("\n\n");
}

/* ---- algorithm drivers ------------------------------------------------------- */

/* q_grover_iterations (reference src/q_utils.c:126-134): floor((pi/4)*sqrt(2^n)) */
static int grover_iterations(int num_qubits) {
  double n_states = ldexp(1.0, num_qubits);
  return (int)floor((QCS_PI / 4.0) * sqrt(n_states));
}

/* qc_grover_search (reference src/qcs.c:402-419) */
void qc_grover_search(t_q_circuit *circuit, int solution_state) {
  int q, it, iterations = grover_iterations(circuit->num_qubits);
  for (q = 0; q < circuit->num_qubits; q++)
    qc_h(circuit, q);
  for (it = 0; it < iterations; it++) {
    if (qcs_cuda_phase_flip(circuit->engine, (long)solution_state) != QCS_CUDA_OK)
      fprintf(stderr, "Error: %s\n", qcs_cuda_last_error());
    qc_add_gate(circuit, "ORACLE", solution_state, -1, 0.0);
    report("qc_grover_search", qcs_cuda_diffusion(circuit->engine));
    qc_add_gate(circuit, "DIFFUSION", -1, -1, 0.0);
  }
}

/* qc_quantum_fourier_transform (reference src/qcs.c:443-457): no final swaps */
void qc_quantum_fourier_transform(t_q_circuit *circuit) {
  int n = circuit->num_qubits, i, j;
  for (i = 0; i < n; i++) {
    qc_h(circuit, i);
    for (j = i + 1; j < n; j++)
      qc_cphase(circuit, j, i, QCS_PI / ldexp(1.0, j - i));
  }
}

/* qc_bernstein_vazirani (reference src/qcs.c:614-644): ancilla = top qubit */
void qc_bernstein_vazirani(t_q_circuit *circuit, int hidden_string) {
  int n = circuit->num_qubits - 1, i;
  if (n <= 0) {
    fprintf(stderr, "Bernstein-Vazirani requires at least 2 qubits (1 input + "
                    "1 ancilla).\n");
    return;
  }
  qc_x(circuit, n);
  qc_h(circuit, n);
  for (i = 0; i < n; i++)
    qc_h(circuit, i);
  qc_barrier(circuit);
  for (i = 0; i < n; i++)
    if ((hidden_string >> i) & 1)
      qc_cnot(circuit, i, n);
  qc_barrier(circuit);
  for (i = 0; i < n; i++)
    qc_h(circuit, i);
}

/* qc_ghz_state (reference src/qcs.c:521-535) */
void qc_ghz_state(t_q_circuit *circuit) {
  int i;
  if (circuit->num_qubits < 2) {
    fprintf(stderr, "Error: GHZ state requires at least 2 qubits.\n");
    return;
  }
  qc_h(circuit, 0);
  for (i = 0; i + 1 < circuit->num_qubits; i++)
    qc_cnot(circuit, i, i + 1);
}

/* ---- introspection ------------------------------------------------------------ */

int qc_get_num_qubits(t_q_circuit *circuit) { return circuit->num_qubits; }
int qc_get_num_gates(t_q_circuit *circuit) { return circuit->num_gates; }

/* qc_optimize (reference src/qcs.c:650-698): history-only peephole.  Removes
 * an adjacent pair of identical H/X/Y/Z on one target, or of identical CNOTs,
 * then rescans from the start.  Amplitudes are never touched. */
void qc_optimize(t_q_circuit *circuit) {
  int i = 0;
  while (i + 1 < circuit->hist_len) {
    struct hist_entry *a = &circuit->hist[i], *b = &circuit->hist[i + 1];
    int same_name = strcmp(a->name, b->name) == 0;
    int self_inverse_1q = same_name && a->target == b->target &&
                          (strcmp(a->name, "H") == 0 || strcmp(a->name, "X") == 0 ||
                           strcmp(a->name, "Y") == 0 || strcmp(a->name, "Z") == 0);
    int cnot_pair = same_name && strcmp(a->name, "CNOT") == 0 &&
                    a->target == b->target && a->control == b->control;
    if (self_inverse_1q || cnot_pair) {
      free(a->name);
      free(b->name);
      memmove(a, a + 2, (size_t)(circuit->hist_len - i - 2) * sizeof(*a));
      circuit->hist_len -= 2;
      circuit->num_gates -= 2;
      i = 0;
    } else {
      i++;
    }
  }
}
