/*
 * qcs_cuda.h -- extern "C" ABI of libqcs_cuda.so, the sm_100a state-vector
 * engine behind the QCS_GPU_CUDA mode of include/qcs.h.
 *
 * This is the drop-in boundary: plain pointers, ints, longs and doubles only
 * (C89-callable; no CUDA, C++ or torch types).  Each entry point replaces one
 * host-side touch of `state->vector` in the reference; the reference file:line
 * it stands in for is cited beside it.  The reference's own GPU seam has the
 * same shape -- `q_apply_1q_gate_gpu(state, gate, target)` and
 * `q_apply_2q_gate_gpu(state, gate, control, target)` (reference
 * src/q_gates.c:13-17, called at :39 and :173) -- but it leaves every other
 * read of the state on the host; this ABI moves all of them to the device.
 *
 * Conventions
 *   - qubit q is bit q of the basis index (reference src/q_gates.c:28).
 *   - a 2x2 gate is `const double m[8]`: row-major U00,U01,U10,U11, each
 *     {re, im} (reference src/q_gates.c:140-141).  It is copied at call time.
 *   - every function returns 0 on success, non-zero on error;
 *     qcs_cuda_last_error() then describes it.  Nothing falls back to the CPU.
 *   - gate calls are deferred into a fusion queue; any call that returns data
 *     (or qcs_cuda_flush) executes the queue first, so results are
 *     observationally identical to eager application.
 *   - not thread-safe per engine; several engines may coexist.
 */
#ifndef QCS_CUDA_H
#define QCS_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qcs_cuda_engine qcs_cuda_engine;

enum {
  QCS_CUDA_OK = 0,
  QCS_CUDA_ERR_INVALID = 1, /* bad qubit / index; state untouched           */
  QCS_CUDA_ERR_CUDA = 2,    /* CUDA runtime / driver failure                */
  QCS_CUDA_ERR_NOMEM = 3,   /* device allocation failed                     */
  QCS_CUDA_ERR_NCCL = 4,    /* communicator failure                         */
  QCS_CUDA_ERR_DRYRUN = 5   /* data requested from a plan-only engine       */
};

/* ---- storage: q_state_init / q_state_free (reference src/q_state.c:32-115) -- */
/* Allocates the device-resident amplitude buffer(s); the state is |0...0> (the
 * reference zeroes the vector and sets amplitude 0 to 1; here the bytes are written
 * lazily -- the first fused pass synthesises its input -- unless "lazy_init" = off).
 * With a communicator installed (qcs_cuda_dist_init) the state is sharded on its
 * top log2(world) qubits and this rank holds one shard. */
int qcs_cuda_state_create(qcs_cuda_engine **out, int n_qubits);
void qcs_cuda_state_destroy(qcs_cuda_engine *e);

/* ---- gates ----------------------------------------------------------------- */
/* q_apply_1q_gate (reference src/q_gates.c:25-149). */
int qcs_cuda_apply_1q(qcs_cuda_engine *e, const double m[8], int target);
/* q_apply_2q_gate, controlled 2x2 (reference src/q_gates.c:158-298). */
int qcs_cuda_apply_c1q(qcs_cuda_engine *e, const double m[8], int control,
                       int target);
/* q_apply_phase_flip (reference src/q_gates.c:305-317). */
int qcs_cuda_phase_flip(qcs_cuda_engine *e, long index);
/* q_apply_diffusion (reference src/q_gates.c:323-356).  The mean is built from the
 * reference's left-to-right sums of the real and of the imaginary parts, replayed
 * exactly on the device (math=exact: every amplitude of a Grover search equals the
 * reference's; math=fast uses a tree sum, which is closer to the exact mean than the
 * reference's drifting sequential sum and therefore NOT within 1e-12 of it on large
 * registers). */
int qcs_cuda_diffusion(qcs_cuda_engine *e);
/* q_state_normalize (reference src/q_utils.c:44-119). */
int qcs_cuda_normalize(qcs_cuda_engine *e);

/* ---- measurement: the loops of qc_measure (reference src/qcs.c:237-284) ---- */
/* prob_0 = sum over basis states with bit `qubit` clear, accumulated in
 * ascending index order (reference src/qcs.c:253-259); the device result is
 * the sequentially rounded sum, bit for bit. */
int qcs_cuda_prob0(qcs_cuda_engine *e, int qubit, double *p0);
/* Zeroes the amplitudes with bit `qubit` != outcome, then normalises
 * (reference src/qcs.c:264-281). */
int qcs_cuda_collapse(qcs_cuda_engine *e, int qubit, int outcome);

/* ---- state access ------------------------------------------------------------ */
/* state->vector[index] (reference src/q_state.c:145-165). */
int qcs_cuda_get_amplitude(qcs_cuda_engine *e, long index, double out[2]);
/* c_norm_sq(state->vector[index]) (reference src/qcs.c:391-395); out-of-range
 * index yields 0.0 and returns QCS_CUDA_OK like the reference. */
int qcs_cuda_probability(qcs_cuda_engine *e, long index, double *p);
/* First strict maximum of |a|^2, all-zero state -> 0 (reference src/qcs.c:464-478). */
int qcs_cuda_argmax(qcs_cuda_engine *e, long *index);

/* ---- sampling: the shot loop of qc_run_shots (reference src/qcs.c:589-605) -- */
/* For each u[k] (= rand()/(double)RAND_MAX drawn by the host in shot order)
 * writes the smallest basis index i with u[k] < p_0 (+) p_1 (+) ... (+) p_i,
 * (+) being the reference's left-to-right rounded accumulation, or -1 when no
 * such i exists (the reference drops that shot). */
int qcs_cuda_sample(qcs_cuda_engine *e, const double *u, int shots,
                    long *indices);

/* ---- queue / bulk transfer ----------------------------------------------------- */
int qcs_cuda_flush(qcs_cuda_engine *e);
/* Copies `count` amplitudes starting at basis index `first` to/from host
 * memory as interleaved {re,im} doubles.  which = 0: live buffer
 * (state->vector), 1: scratch buffer (state->scratch_vector; only maintained
 * under reference semantics).  Used by qc_print_state and by
 * the parity tests. */
int qcs_cuda_read_amplitudes(qcs_cuda_engine *e, int which, long first,
                             long count, double *out);
int qcs_cuda_write_amplitudes(qcs_cuda_engine *e, int which, long first,
                              long count, const double *in);

/* ---- configuration / introspection ------------------------------------------------ */
/* Keys: "semantics" = reference|corrected, "fusion" = on|off,
 * "dryrun" = 0|1, "pass_flops" = <float>, "tile_kernel" = ldg8|ldg|tma|tma16,
 * "exchange" = p2p|nccl (multi-GPU position swaps: in-place peer-memory kernel, or NCCL send/recv),
 * "tile_bits" = 10|11|12 (largest tile of a fused pass, default 11; a pass runs on the smallest tile that holds its pairing qubits),
 * "peephole" = on|off (corrected semantics: drop exactly self-cancelling X/Y/Z/CNOT/CZ pairs from the queue),
 * "fuse_swaps" = on|off (p2p only: a swap rides on the stores of a fused pass instead of a kernel of its own),
 * "math" = exact|fast (default exact: every amplitude bit-identical to the reference's sequential mode.
 *   fast -- corrected semantics, tile_kernel ldg8|ldg -- trades that for throughput: fused multiply-adds,
 *   runs of controlled phases merged into one factor, commuting gates scheduled out of order; amplitudes
 *   then agree with the reference within 1e-12 relative),
 * "reorder" = on|off, "reorder_segments" = 1..12 (math=fast: commutation-aware scheduling and how many
 *   segments a reordered pass may spend),
 * "lazy_init" = on|off, "fuse_argmax" = on|off (qc_find_most_likely_state behind queued gates takes one
 *   candidate per tile from the flush's last pass), "swap_store" = bulk|thread, "remap_max" = 1..3
 *   (position pairs one carrying pass may trade), "victim_policy" = mru|lru (INTEGRATION.md lists all keys).
 * Defaults come from QCS_CUDA_<KEY> in the environment; set_default applies
 * to engines created afterwards. */
int qcs_cuda_set_default(const char *key, const char *value);
/* The process-wide default for `key` set by qcs_cuda_set_default: copies it into buf (when cap allows)
 * and returns its length, or -1 when no default is set (the environment / built-in value applies). */
long qcs_cuda_get_default(const char *key, char *buf, long cap);
/* Freed state buffers are kept for the next qc_create of the same size (cudaMalloc / cudaFree of a
 * 16 GiB state cost 10-350 ms); this returns them to the driver.  Returns the bytes released. */
long qcs_cuda_trim_pool(void);
int qcs_cuda_num_qubits(const qcs_cuda_engine *e);
/* Current layout: perm[q] = physical index position of logical qubit q (identity on one GPU;
 * position swaps change it when the state is sharded).  perm must hold n_qubits ints. */
int qcs_cuda_get_layout(const qcs_cuda_engine *e, int *perm);

typedef struct qcs_cuda_stats {
  long gates_submitted;     /* qcs_cuda_apply_* calls accepted            */
  long gates_executed;         /* gates that reached a kernel                */
  long passes;                 /* fused tile-kernel launches                 */
  long kernel_launches;        /* every kernel this library launched         */
  long segments;               /* in-tile register re-assignments            */
  long remaps;                 /* global<->local qubit exchanges             */
  double algorithmic_bytes;    /* SURVEY 8(d) bytes of all executed passes   */
  double pass_bytes;           /* same, fused tile passes only               */
  double pass_ms;              /* device time inside fused tile passes       */
  double exchange_bytes;       /* bytes sent over NVLink by this rank        */
  double exchange_ms;          /* device time inside stand-alone exchanges   */
  long fused_remaps;           /* remaps folded into the stores of a pass    */
  double fused_remap_pass_ms;  /* device time of the passes that carried one
                                  (also counted in pass_ms)                  */
  double pass_flops_per_amp;   /* separately rounded FP64 operations per
                                  amplitude, summed over executed passes     */
  long gates_cancelled;        /* dropped by the queue peephole (exact pairs) */
  long multi_remaps;           /* carrying passes that traded 2 or 3 position pairs at once (counted in fused_remaps) */
  double fused_remap_bytes;    /* bytes per direction this rank moved over NVLink inside carrying passes (part of exchange_bytes) */
  long out_of_place_remaps;    /* carrying passes that wrote the ranks' second shard buffers (option remap_buffer): no per-tile handshake */
} qcs_cuda_stats;

int qcs_cuda_get_stats(qcs_cuda_engine *e, qcs_cuda_stats *out);
int qcs_cuda_reset_stats(qcs_cuda_engine *e);
/* When enabled (non-zero) every fused pass is bracketed by CUDA events on the
 * engine's stream so pass_ms is filled in (used by bench.py's roofline). */
int qcs_cuda_set_timing(qcs_cuda_engine *e, int enabled);

/* Device-side stopwatch: record a CUDA event on the engine's stream (slot 0..7),
 * later read the elapsed device time between two recorded slots (synchronises). */
int qcs_cuda_marker_record(qcs_cuda_engine *e, int slot);
int qcs_cuda_marker_elapsed_ms(qcs_cuda_engine *e, int from_slot, int to_slot,
                               double *ms);

/* Measures this GPU's issue rate of separately rounded FP64 operations (mul.rn / add.rn, never
 * fused): the arithmetic ceiling of the gate kernels, reported by bench.py beside the HBM one. */
int qcs_cuda_probe_fp64(double *ops_per_second);
/* sizeof the kernel-parameter block that describes one fused pass (what a flush ships to the GPU). */
long qcs_cuda_pass_descriptor_bytes(void);

/* Writes a human-readable description of the passes the last flush executed
 * (tile bits, segments, gates per segment) into buf; returns bytes needed. */
long qcs_cuda_describe_last_plan(qcs_cuda_engine *e, char *buf, long cap);

/* Copies the kernel-parameter block (PassParams, common.h) of pass `pass_index` of the last flush
 * into buf (when cap is large enough); returns the number of passes of the last flush.  Lets a
 * CPU test interpret exactly what the GPU would be handed (tests/test_planner.py). */
long qcs_cuda_last_plan_raw(qcs_cuda_engine *e, long pass_index, void *buf, long cap);
/* math=fast: the thread-table fans that travel next to pass `pass_index` (PassParams::n_thread_tables
 * tables of 2^(tile_bits - reg_bits) complex numbers {re, im}, common.h QCS_OP_TFAN_BASE), copied into
 * buf when cap (in doubles) is large enough; returns their size in doubles. */
long qcs_cuda_last_plan_tables(qcs_cuda_engine *e, long pass_index, double *buf, long cap);
/* Sharded engines: how many (local position, global position) pairs entry `pass_index` of the last
 * flush trades -- all at once on the stores of that pass (a remap of up to 3 pairs: an all-to-all
 * among 2^k ranks), or (plan-only engines list these too, as entries with no segments) as a
 * stand-alone swap between passes -- and which: lpos / gpos must hold 3 ints each.  0: none. */
long qcs_cuda_last_plan_swap(qcs_cuda_engine *e, long pass_index, int *lpos, int *gpos);
/* Per-pass figures of the last flush (timing enabled): device time in ms (-1 when not measured), the
 * planner's FP64 operation count per amplitude, tile bits and segment count of pass `pass_index`.
 * Any out pointer may be NULL.  Returns the number of passes of the last flush. */
long qcs_cuda_last_plan_pass_info(qcs_cuda_engine *e, long pass_index, double *ms, double *flops_per_amp,
                                  int *tile_bits, int *segments);

/* Dry-run engines record what a real engine would execute, in order.  Entry i is
 * written as 12 doubles: out[0] = 1 (gate) or 2 (position swap);
 *   gate: out[1] = kind | flags << 8 (common.h GateKind / GateFlags), out[2] = target position,
 *         out[3] = control position (-1: none), out[4..12) = the 2x2 matrix;
 *   swap: out[1] = local position, out[2] = global position;  out[0] = 3: both positions local.
 * Returns the number of entries; fills `out` when index is in range. */
long qcs_cuda_trace_read(qcs_cuda_engine *e, long index, double out[12]);

const char *qcs_cuda_last_error(void);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) ---------------------------- */
/* Rank 0 obtains an id, the launcher broadcasts the 128 bytes (bench.py uses
 * torch.distributed), then every rank calls qcs_cuda_dist_init.  Engines
 * created afterwards are shards of one logical state. */
int qcs_cuda_dist_unique_id(char id[128]);
int qcs_cuda_dist_init(int rank, int world, const char id[128], int device);
int qcs_cuda_dist_finalize(void);
/* Plan-only sharding (no NCCL, no GPU): engines created afterwards must be
 * dry-run ("dryrun" = 1); they plan passes and position swaps exactly as rank
 * `rank` of `world` would and record them in the trace below.  Used by the CPU
 * tests that replay the schedule with two gloo processes. */
int qcs_cuda_dist_init_plan_only(int rank, int world);
int qcs_cuda_dist_rank(void);
int qcs_cuda_dist_world(void);

/* Exported by the host layer (libqcs.so), not by libqcs_cuda.so: the engine
 * behind a circuit, for callers that want the bulk-transfer / statistics entry
 * points above next to the qcs.h API (tests, bench.py). */
struct t_q_circuit;
qcs_cuda_engine *qc_cuda_engine(struct t_q_circuit *circuit);
/* qc_run_shots without the dense 2^n histogram (reference src/qcs.c:575-607 draws and decides the
 * same way): indices[s] = outcome of shot s, or -1 where the reference drops the shot.  Consumes
 * `shots` values of rand(), like qc_run_shots. */
void qc_run_shots_sparse(struct t_q_circuit *circuit, int shots, long *indices);

#ifdef __cplusplus
}
#endif
#endif /* QCS_CUDA_H */
