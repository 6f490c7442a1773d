/*
 * qsim_b200.h -- C ABI of libqsim_b200.so, the B200 (sm_100a) state-vector path.
 *
 * This is the drop-in boundary for the ONE hot path of
 * RiccardoFiorentini/GPU_quantum_simulator: "apply a gate list to a 2^n complex
 * state vector".  The reference has no FFI layer; its boundary is the set of
 * non-static C functions of quantum_simulator.c (prototypes at :25-30) plus the
 * CLI of main() (:32-79).  Every entry point below names what it replaces.
 *
 * Plain C: pointers, sizes, ints.  No C++/torch types, no exceptions, the
 * library never calls exit().  All calls are blocking and return QSB_OK (0) or
 * a negative qsb_status; qsb_last_error() gives the message (thread-local).
 * A handle is not re-entrant.  There is NO CPU fallback: without a usable
 * CUDA device qsb_create() fails with QSB_ERR_NOGPU.
 *
 * Conventions (all inherited from the reference):
 *   - little-endian qubits: qubit k <-> bit k of the amplitude index
 *     (quantum_simulator.c:83, :96-99)
 *   - a gate is a (multi-)controlled single-qubit unitary; the matrix is
 *     row-major  out0 = m00*a + m01*b, out1 = m10*a + m11*b  where a/b are the
 *     amplitudes with target bit 0/1 (quantum_simulator_naive.cu:82-86; equal to
 *     quantum_simulator.c:88-89 for every matrix the reference builds)
 *   - "rz(theta)" in circuit text is the PHASE gate diag(1, e^{i theta})
 *     (quantum_simulator.c:205-208)
 */
#ifndef QSIM_B200_H
#define QSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum qsb_status {
    QSB_OK = 0,
    QSB_ERR_ARG = -1,     /* bad argument */
    QSB_ERR_CUDA = -2,    /* CUDA runtime error (message has file:line, cf. CHECK, naive.cu:29-47) */
    QSB_ERR_NOMEM = -3,   /* host or device allocation failed ("Malloc error", quantum_simulator.c:170) */
    QSB_ERR_PARSE = -4,   /* circuit text rejected ("Unknown token", :213) */
    QSB_ERR_NOGPU = -5,   /* no CUDA device: the product has no CPU path */
    QSB_ERR_COMM = -6,    /* multi-GPU exchange failed */
    QSB_ERR_IO = -7       /* cannot open circuit file (:129) */
} qsb_status;

enum { QSB_F32 = 32, QSB_F64 = 64 };

/* One gate of the IR: controls-mask + target + 2x2 complex matrix.
 * Replaces the (U[4], target) / (control, target) argument pairs of
 * execute_single_qubit_gate (:81) and execute_cnot (:94), and one row of the
 * four arrays parse_circuit() fills in the CUDA variants (naive.cu:224-402). */
typedef struct qsb_gate {
    uint64_t controls; /* bit k set: qubit k must be 1 for the gate to act   */
    int32_t target;    /* qubit the 2x2 matrix acts on                        */
    int32_t flags;     /* reserved, 0                                         */
    double m[8];       /* m00 m01 m10 m11, each (re, im)                      */
} qsb_gate_t;

/* Execution modes.  SWEEP = one full pass over the state per gate, the
 * reference's own schedule (naive.cu:163-189) kept as a cross-check and as the
 * "unfused" baseline; TILED = the fused tile-pass schedule (the product). */
enum { QSB_MODE_TILED = 0, QSB_MODE_SWEEP = 1,
       /* DENSE = experiment of BASELINE.json configuration 4: gates merged greedily into dense k-qubit unitaries
        * (k = tile_bits, 2..5; the reference's 4x4 accumulator of quantum_simulator_4x4.cu:327-501 generalised), one
        * sweep per block like kernel_gate_4 (:109-146).  Single GPU.  Measured slower than TILED at every k
        * (DESIGN.md section 3.1); kept for the k sweep and the tensor-core check, not as a product path. */
       QSB_MODE_DENSE = 2 };

typedef struct qsb_options {
    int32_t precision;    /* QSB_F32 | QSB_F64 (state dtype on the device)    */
    int32_t device;       /* CUDA ordinal, -1 = current                       */
    int32_t mode;         /* QSB_MODE_*                                       */
    int32_t tile_bits;    /* TILED: reserved, the tile is 2^12 (f32) / 2^11 (f64) amplitudes (build-time, csrc/tiled.h);
                           * DENSE: the fusion width k, 2..5 (0 = 4) */
    int32_t low_bits;     /* contiguous low index bits every tile keeps, 0 = default (4 f32 / 3 f64) */
    int32_t rank;         /* this process' shard, 0..world-1                  */
    int32_t world_size;   /* power of two; state sharded on the top log2(world) qubits */
    int32_t use_graph;    /* 1 = capture the pass launches of a plan into a CUDA graph on its first qsb_execute and
                           * replay it afterwards (single-GPU tiled plans; for plans executed many times on small
                           * registers, where a pass is shorter than its launch).  0 = plain stream launches. */
    int32_t verbose;
    /* planner / exchange tuning knobs, all 0 = default (used by bench.py A/B runs and the tests):
     *   [0] minimum number of local gates a pass must still find before a qubit exchange is scheduled (default 10)
     *   [1] 2 = keep the qubits of phase gates thread-level at any price ("lazy diagonals")
     *   [2] k+1 = trim tail rounds of SM-bound passes that hold fewer than k gates (default k = 2; 1 = off)
     *   [3] fusion-depth cap: stop adding rounds to a pass at this estimated SM cost (unit-form gate units)
     *   [4] 1 = do not defer phase gates that touch a vector bit; 2 = no 2x2 products of consecutive one-qubit gates;
     *       4 = no merged controlled-phase runs (G_DIAGA); 5 = a CX between two h on its target stays a CX (default: h cx h becomes one
     *       controlled phase); 6 = the rewrite also with an h on one side only (fewer passes, more diagonal slots)
     *   [5] exchange flavour: 1 direct fused peer scatter (victims trade places with the rank bits wherever they are),
     *       2 NCCL all-to-all, 3 pipelined copy-engine exchange, 4 round-1 fused scatter (victims moved to the top local
     *       positions first)   (0: chosen by qsb_comm_init -- 1, or 2 if the peer shards cannot be mapped)
     *       With [0] = 0 a sharded plan is built for a few exchange thresholds and the cheapest schedule kept.
     *   [6] 1 = do not sink thread-level phases to later rounds; 2 = first-come tile choice (no hill climbing);
     *       3 = no lane relocation at the end of a pass (an empty round turns the registers instead, as in round 1);
     *       4 = lane relocation only for the qubits in conflict (default: every pass re-picks the qubits that live on
     *       the lane positions).  With [6] = 0 a sharded plan also tries 4 and keeps the cheaper schedule. */
    int32_t reserved[7];
} qsb_options_t;

typedef struct qsb_sim qsb_t;      /* simulator: owns the device state          */
typedef struct qsb_plan qsb_plan_t; /* a fused, scheduled circuit (device-ready) */

/* Counters of the last qsb_execute / qsb_apply_gates on a handle. */
typedef struct qsb_run_stats {
    double device_ms;         /* CUDA-event time, first pass start -> last pass end */
    double plan_ms;           /* host fusion + scheduling time                 */
    uint64_t source_gates;    /* gates handed in                               */
    uint64_t device_ops;      /* ops after canonicalisation / fusion           */
    uint32_t passes;          /* full sweeps over the local state              */
    uint32_t rounds;          /* shared-memory exchange rounds, all passes     */
    uint32_t swaps;           /* global<->local qubit exchanges (multi-GPU)    */
    uint32_t kernel_launches; /* kernels launched in the timed region          */
    uint64_t bytes_moved;     /* algorithmic HBM bytes: sum over passes of 2*N_loc*B */
    uint64_t bytes_exchanged; /* bytes this rank sent over NVLink              */
    double exchange_ms;       /* part of device_ms spent in exchanges          */
} qsb_run_stats_t;

/* ---- lifecycle ---------------------------------------------------------- */
void qsb_options_default(qsb_options_t *opt);
/* Allocates 2^(num_qubits)/world amplitudes on the device and sets |0...0>.
 * Replaces malloc + init at quantum_simulator.c:168-177 and cudaMalloc +
 * init_state_vector at naive.cu:148-158. */
int qsb_create(qsb_t **out, int num_qubits, const qsb_options_t *opt);
void qsb_destroy(qsb_t *s);
int qsb_reset(qsb_t *s); /* back to |0...0> */
int qsb_num_qubits(const qsb_t *s);
int qsb_precision(const qsb_t *s);

/* ---- the hot path --------------------------------------------------------
 * Replaces the per-gate host loop (quantum_simulator.c:145-243,
 * naive.cu:163-189) and the host "preprocessing" of the CUDA variants
 * (preproces.cu:215-269, 4x4.cu:327-501, 4x4_permute.cu:350-434). */
int qsb_apply_gates(qsb_t *s, const qsb_gate_t *gates, size_t n); /* plan + execute + free */
/* A plan is made for the qubit layout the handle has at qsb_plan_create (the identity after qsb_create /
 * qsb_reset; fused passes leave the qubits permuted, qsb_get_layout).  qsb_execute from any other layout is
 * refused with QSB_ERR_ARG instead of computing on the wrong bits: to run a plan again, qsb_reset first (or
 * plan again from the current layout, which is what qsb_apply_gates does). */
int qsb_plan_create(qsb_t *s, const qsb_gate_t *gates, size_t n, qsb_plan_t **out);
int qsb_execute(qsb_t *s, qsb_plan_t *plan);
void qsb_plan_destroy(qsb_plan_t *plan);
int qsb_plan_stats(const qsb_plan_t *plan, qsb_run_stats_t *out); /* static counters, times = 0 */
int qsb_last_run_stats(const qsb_t *s, qsb_run_stats_t *out);
/* Host-only scheduling (no device needed): what qsb_plan_create would build
 * for a (num_qubits, precision, world) machine.  Used by CPU tests. */
int qsb_plan_dry_run(int num_qubits, const qsb_options_t *opt, const qsb_gate_t *gates,
                     size_t n, qsb_run_stats_t *out);

/* ---- readout --------------------------------------------------------------
 * Amplitudes come back in LOGICAL index order as interleaved (re, im) doubles,
 * i.e. the memory layout of the `complex *v` compute_state_vector returns
 * (quantum_simulator.c:125).  first/count address the GLOBAL index space; in a
 * multi-GPU run each rank may only read its own shard. */
int qsb_download(qsb_t *s, double *re_im, uint64_t first, uint64_t count);
int qsb_upload(qsb_t *s, const double *re_im, uint64_t first, uint64_t count);
/* The shard as it lies on the device: amplitudes first..first+count of the LOCAL PHYSICAL index space
 * (fp64 pairs), plus the qubit layout needed to interpret it: perm64[q] = physical index bit of
 * logical qubit q (bits >= *nloc are the rank bits).  This is how a sharded state is gathered:
 * fused passes may leave the qubits permuted, and no rank owns a contiguous logical range. */
int qsb_download_physical(qsb_t *s, double *re_im, uint64_t first, uint64_t count);
int qsb_get_layout(const qsb_t *s, int8_t *perm64, int *nloc);
/* Raw copy in the device dtype (float or double pairs), logical order. */
int qsb_download_native(qsb_t *s, void *dst, uint64_t first, uint64_t count);
/* sum |a|^2 over the local shard, and the local argmax (global index). */
int qsb_norm_argmax(qsb_t *s, double *norm, uint64_t *argmax_idx, double *argmax_p);
/* |a_i|^2 for i in [first, first+count), fp64. */
int qsb_probabilities(qsb_t *s, double *p, uint64_t first, uint64_t count);
/* Inclusive prefix sum of |a|^2 in logical index order -- compute_state_cumulative_distribution
 * (:256-268).  |a|^2 and the prefix sums are computed on the device (block scans chained by serial
 * offsets: monotone, equal to the reference's single fp64 accumulator up to re-association, <= 1e-12). */
int qsb_cdf(qsb_t *s, double *cdf, uint64_t first, uint64_t count);
/* `shots` draws from the distribution; same search rule as measurement() (:270-283: first index
 * with cdf != 0 and cdf >= r, clamped to the last index), r from a seeded generator instead of
 * rand().  Device-side and streaming: no 2^n array is built (segment totals + one re-scanned
 * segment per shot), so it works at any n that fits the device.  On one GPU the result equals a
 * search of qsb_cdf with the same r.  On a sharded state every rank must call it with the same
 * seed and shots (collective: per-rank totals are all-gathered, the shots all-reduced); every rank
 * receives all shots, drawn in rank-major physical order and returned as logical indices. */
int qsb_sample(qsb_t *s, uint64_t seed, int shots, uint64_t *out);
/* The r of shot k for a seed (splitmix64 -> 53-bit uniform in [0,1)): lets a caller replay
 * measurement()'s search on a CDF of its own. */
double qsb_sample_uniform(uint64_t seed, int k);

/* ---- shard dump / reload ----------------------------------------------------
 * The reference keeps the state only in memory (quantum_simulator.c:75 frees it
 * at exit); long sharded runs want a checkpoint.  File = 128-byte header (magic
 * "QSBSHARD", version, qubits, precision, world, rank, local bits, the logical ->
 * physical qubit map) + the shard exactly as it lies in HBM.  One file per rank;
 * loading checks that the handle has the same shape and restores the qubit map. */
int qsb_save_state(qsb_t *s, const char *path);
int qsb_load_state(qsb_t *s, const char *path);

/* ---- multi-GPU (one process per GPU) --------------------------------------
 * nccl_unique_id: the 128 bytes of an ncclUniqueId, identical on every rank
 * (qsb_comm_unique_id makes one on rank 0; the caller broadcasts it, e.g. with
 * torch.distributed).  The reference has no multi-GPU path. */
int qsb_comm_unique_id(void *id128);
int qsb_comm_init(qsb_t *s, const void *nccl_unique_id128);

/* ---- QASM front end -------------------------------------------------------
 * Accepts the reference grammar bit-for-bit (quantum_simulator.c:133-242:
 * two header statements, `qubit[n] q;` or `qubit q[n];`, operands `q[k]` or
 * `$k`, CRLF, gate set cx x sx z s sdg t tdg rz h) plus a superset
 * (y p rx ry u cz cp crz swap ccx ..., `gate` definitions, ctrl @ / negctrl @ / inv @ / pow(k) @
 * modifiers, gphase, several registers, whole-register operands, `pi` expressions and the
 * usual functions in parameters; barrier / measure / reset / classical statements ignored).
 * The CUDA variants' "<num_q> <num_g>" header (naive.cu:239-240) is accepted
 * too.  *gates is malloc'ed; release with qsb_free(). */
int qsb_parse_qasm_file(const char *path, int *num_qubits, qsb_gate_t **gates, size_t *n);
int qsb_parse_qasm_string(const char *text, int *num_qubits, qsb_gate_t **gates, size_t *n);
/* Fill `g` with a named gate of the table at quantum_simulator.c:184-211 (+superset). */
int qsb_gate_from_name(const char *name, const double *params, int nparams,
                       const int *qubits, int nqubits, qsb_gate_t *out, int *nout);
void qsb_free(void *p);

const char *qsb_last_error(void);
const char *qsb_version(void);

/* ---- reference-compatible shims --------------------------------------------
 * Same names, argument meaning, ownership (callee mallocs, caller frees) and
 * stdout behaviour as the functions at quantum_simulator.c:25-30, so the
 * reference's own main() links against libqsim_b200.so unchanged (see
 * INTEGRATION.md).  `v` is a host `double complex *`; every call runs on the
 * GPU.  Declared with void* / double* here so this header needs no <complex.h>.
 */
double *qsb_ref_compute_state_vector(const char *filename, int *num_q);       /* -> compute_state_vector */
void qsb_ref_execute_single_qubit_gate(double *v, int num_q, const double U[8], int target); /* U as complex U[4] */
void qsb_ref_execute_cnot(double *v, int num_q, int control, int target);
double *qsb_ref_compute_state_cumulative_distribution(const double *v, int num_q);
long long qsb_ref_measurement(const double *cumul, int num_q);

#ifdef __cplusplus
}
#endif
#endif /* QSIM_B200_H */
