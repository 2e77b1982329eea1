/*
 * qcs.h -- public C89 circuit API of QCS, B200 edition.
 *
 * Drop-in for the reference header (reference include/qcs.h:28-74): the same
 * opaque `t_q_circuit`, the same 27 functions with the same signatures and
 * argument meaning, so a program written against the reference recompiles
 * unchanged.  What differs is the execution mode list:
 *
 *   QCS_GPU_CUDA   (this repository; the only mode it implements)
 *       amplitudes live in persistent device-resident complex-double buffers
 *       on an NVIDIA B200 (sm_100a); every gate / measurement / sampling call
 *       is forwarded through the extern "C" ABI of include/qcs_cuda.h into
 *       libqcs_cuda.so.  There is no CPU fallback: without the CUDA library
 *       and a GPU, qc_create() prints an error and returns NULL.
 *
 * The reference's other modes (sequential, QCS_MULTI_THREAD, QCS_CPU_OPENMP,
 * QCS_GPU_OPENCL, QCS_SIMD_ONLY -- reference include/qcs.h:10-23) are CPU or
 * OpenCL back-ends of the reference itself and are not part of this library.
 *
 * Run-time knobs (environment, read by qc_create):
 *   QCS_CUDA_SEMANTICS = reference | corrected   (default reference)
 *       reference: bit-compatible with the reference's sequential mode,
 *                  including its controlled-gate / phase-flip / diffusion
 *                  behaviour (DESIGN.md "Semantics").
 *       corrected: controlled gates apply both matrix rows, phase flip is in
 *                  place, diffusion uses sum/N.
 *   QCS_CUDA_FUSION    = on | off                (default on)
 *   QCS_CUDA_DEVICE    = <ordinal>               (default: current device)
 *
 * Link:  -lqcs -lqcs_cuda   
 */
#ifndef QCS_H
#define QCS_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct t_q_circuit t_q_circuit;

/* lifetime ---------------------------------------------------------------- */
t_q_circuit *qc_create(int num_qubits);
void qc_destroy(t_q_circuit *circuit);

/* fixed single-qubit gates and CNOT --------------------------------------- */
void qc_h(t_q_circuit *circuit, int qubit);
void qc_x(t_q_circuit *circuit, int qubit);
void qc_y(t_q_circuit *circuit, int qubit);
void qc_z(t_q_circuit *circuit, int qubit);
void qc_cnot(t_q_circuit *circuit, int control, int target);

/* parameterised single-qubit gates ---------------------------------------- */
void qc_phase(t_q_circuit *circuit, int qubit, double angle);
void qc_rx(t_q_circuit *circuit, int qubit, double angle);
void qc_ry(t_q_circuit *circuit, int qubit, double angle);
void qc_rz(t_q_circuit *circuit, int qubit, double angle);

/* bookkeeping ops ----------------------------------------------------------- */
void qc_barrier(t_q_circuit *circuit);
void qc_reset(t_q_circuit *circuit, int qubit);

/* collapsing measurement (uses libc rand(), one draw per measured qubit) --- */
int qc_measure(t_q_circuit *circuit, int qubit);
void qc_measure_all(t_q_circuit *circuit, int *results);

/* execution ----------------------------------------------------------------- */
void qc_run(t_q_circuit *circuit);
/* results must hold 2^num_qubits ints; it is zeroed, then one rand() draw per
 * shot selects a basis state by inverse CDF. */
void qc_run_shots(t_q_circuit *circuit, int shots, int *results);

/* state access -------------------------------------------------------------- */
int qc_find_most_likely_state(t_q_circuit *circuit);
double qc_get_probability(t_q_circuit *circuit, int state);
void qc_print_state(t_q_circuit *circuit, int solution_index);
void qc_print_circuit(t_q_circuit *circuit);

/* algorithm drivers ---------------------------------------------------------- */
void qc_grover_search(t_q_circuit *circuit, int solution_state);
void qc_quantum_fourier_transform(t_q_circuit *circuit);
void qc_bernstein_vazirani(t_q_circuit *circuit, int hidden_string);
void qc_ghz_state(t_q_circuit *circuit);

/* introspection --------------------------------------------------------------- */
int qc_get_num_qubits(t_q_circuit *circuit);
int qc_get_num_gates(t_q_circuit *circuit);
void qc_optimize(t_q_circuit *circuit);

/* Exported by the reference without a prototype (reference src/qcs.c:132 and
 * :431; README lists qc_cphase).  Declared here so callers need no extern. */
void qc_cphase(t_q_circuit *circuit, int control, int target, double angle);
void qc_add_gate(t_q_circuit *circuit, const char *gate_name, int target,
                 int control, double param);

#ifdef __cplusplus
}
#endif
#endif /* QCS_H */
