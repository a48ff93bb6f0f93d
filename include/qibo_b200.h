/*
 * qibo_b200 -- C ABI of the B200-native state-vector engine (drop-in boundary).
 *
 * The reference (qiboteam/qibo 0.3.5) has no FFI: its hot path is Python/NumPy behind the
 * `qibo.backends.Backend` plugin API.  Each entry point below names the reference method whose
 * arithmetic it replaces (paths relative to /root/reference/src/qibo).  The Python host layer
 * (`qibo_b200/backend.py`, a `Backend` subclass loaded through `MetaBackend.load`,
 * backends/__init__.py:325-350) binds these through ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers are caller-owned (torch allocates them in the
 *     Python host; `qb_malloc` exists for non-torch hosts);
 *   - qubit ids are Qibo's: qubit 0 is the MOST significant bit of the flat state index
 *     (tests/test_gates_gates.py:39-43), qubit q <-> bit position nqubits-1-q;
 *   - matrices/diagonals are HOST pointers to interleaved complex128 (re,im doubles), row-major, the
 *     row/column index having targets[0] as its most significant bit (gates/abstract.py:439-442);
 *     they are evaluated in double on the host and cast to the state's dtype before the multiply, as
 *     npmatrices.py:21-24 does;
 *   - every call returns 0 on success or a negative QB_ERR_* code; the message is in qb_last_error()
 *     (thread-local).  Nothing throws or aborts across the ABI;
 *   - work is enqueued on the handle's CUDA stream; calls that return host data synchronise that stream.
 *   - there is NO CPU fallback: without a CUDA device qb_create fails with QB_ERR_CUDA.
 */
#ifndef QIBO_B200_H
#define QIBO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB_VERSION 100 /* 0.1.0 */

/* state dtypes (Backend.dtype, backends/abstract.py:133-179) */
#define QB_C64 0  /* complex64  : float2  amplitudes */
#define QB_C128 1 /* complex128 : double2 amplitudes */
/* real dtypes of probability vectors (real(state).dtype, abstract.py:2751) */
#define QB_F32 0
#define QB_F64 1

#define QB_OK 0
#define QB_ERR_INVALID (-1)     /* bad argument            -> ValueError in the host layer          */
#define QB_ERR_OOM (-2)         /* cudaErrorMemoryAllocation -> Backend.oom_error -> RuntimeError    */
#define QB_ERR_CUDA (-3)        /* any other CUDA failure / no device                                */
#define QB_ERR_UNSUPPORTED (-4) /* -> NotImplementedError                                            */

#define QB_MAX_QUBITS 40
#define QB_MAX_OP_TARGETS 6   /* dense blocks: FusedGate / Unitary up to 6 targets                   */
#define QB_MAX_OP_CONTROLS 32

typedef struct qb_context* qb_handle;

/* ---- library / context -------------------------------------------------------------------------- */
int qb_version(void);
const char* qb_last_error(void);
int qb_device_count(int* count);
/* One context per (device, stream).  stream == NULL: the library creates its own non-blocking stream.
 * Replaces NumpyBackend.__init__/set_device (backends/numpy.py:20-36, 76-87). */
int qb_create(int device, void* cuda_stream, qb_handle* out);
int qb_destroy(qb_handle h);
int qb_set_stream(qb_handle h, void* cuda_stream);
int qb_sync(qb_handle h);
int qb_mem_info(qb_handle h, size_t* free_bytes, size_t* total_bytes);
int qb_malloc(qb_handle h, size_t bytes, void** dptr);
int qb_free(qb_handle h, void* dptr);
/* kind: 0 host->device, 1 device->host, 2 device->device.  Replaces Backend.cast / to_numpy
 * (backends/numpy.py:38-64, 98-107) for the state buffer. */
int qb_memcpy(qb_handle h, void* dst, const void* src, size_t bytes, int kind);
/* Same copy, enqueued and NOT waited for, on `cuda_stream` (a cudaStream_t; NULL = the handle's stream).  kind 2 also
 * copies between a local buffer and a peer-mapped one (qb_ipc_open_handle): the DMA engines move the bytes over NVLink
 * without occupying an SM -- the transport of the chunk-pipelined global<->local exchange (models/distcircuit.py:185-268
 * is the reference's one-swap-at-a-time scheme). */
int qb_memcpy_async(qb_handle h, void* dst, const void* src, size_t bytes, int kind, void* cuda_stream);

/* ---- K6: state construction (abstract.py:2243-2273 zero_state, :2199-2241 plus/minus_state) ------- */
int qb_state_set_basis(qb_handle h, void* state, int nqubits, int dtype, uint64_t index);
int qb_state_fill(qb_handle h, void* state, int nqubits, int dtype, double re, double im);
/* complex64 <-> complex128 conversion of `count` amplitudes (Backend.cast with a dtype change) */
int qb_state_cast(qb_handle h, void* dst, int dst_dtype, const void* src, int src_dtype, uint64_t count);
/* sum |amp|^2 over the buffer (deterministic two-stage reduction) */
int qb_state_norm2(qb_handle h, const void* state, int nqubits, int dtype, double* out_host);

/* ---- K1: one gate, one sweep (abstract.py:2322-2361 apply_gate; :3176-3197 _apply_gate_controlled_by)
 * `matrix` is the 2^nt x 2^nt target matrix; `controls` are the control_qubits of a `controlled_by`
 * gate (may be empty).  Named controlled gates (CNOT, CZ, CU1, TOFFOLI...) may be passed with their
 * full matrix over gate.qubits and no controls: the library detects control / diagonal / swap
 * structure exactly and touches only the amplitudes that can change.  In place. */
int qb_apply_matrix(qb_handle h, void* state, int nqubits, int dtype, const double* matrix, int ntargets,
                    const int* targets, int ncontrols, const int* controls);
/* same with only the diagonal given (2^nt complex entries) */
int qb_apply_diagonal(qb_handle h, void* state, int nqubits, int dtype, const double* diag, int ntargets,
                      const int* targets, int ncontrols, const int* controls);

/* ---- K2: a whole gate queue, several gates per HBM sweep (abstract.py:3321-3322 the gate loop of
 * _execute_circuit; FusedGate blocks from models/circuit.py:954-1003 arrive as dense ops) ----------- */
typedef struct {
  int32_t ntargets;
  int32_t ncontrols;
  int32_t targets[QB_MAX_OP_TARGETS];
  int32_t controls[QB_MAX_OP_CONTROLS];
  int32_t is_diagonal;  /* 1: `data` holds 2^nt diagonal entries instead of a matrix */
  int32_t reserved;
  const double* data;   /* host, interleaved complex128 */
} qb_op;

typedef struct {
  int32_t nops;            /* gates submitted                                   */
  int32_t nsweeps;         /* HBM sweeps (kernel launches) they were packed into */
  int32_t ndense_passes;   /* shared-memory passes over tiles, total             */
  int32_t ndiag_ops;       /* diagonal bundles applied                           */
  double bytes_moved;      /* algorithmic bytes: sum over sweeps of 2*B*2^n      */
  float elapsed_ms;        /* CUDA-event time of the whole program (flags & QB_PROGRAM_TIME) */
  int32_t nstage_sweeps;   /* of nsweeps: sweeps made of straight-line stage passes only (the lean two-team kernel) */
  int32_t perm_fused;      /* *_permuted calls: 1 = the trailing qubit permutation rode on the last sweep (no K8 launch) */
  int32_t reserved;
} qb_program_stats;

#define QB_PROGRAM_TIME 1      /* bracket with CUDA events, synchronise, fill elapsed_ms */
#define QB_PROGRAM_NO_FUSE 2   /* one sweep per gate (gate-by-gate accounting)           */
#define QB_PROGRAM_INPUT_ZERO 8 /* qb_apply_program* / qb_program_run*: the input is |0...0> (zero_state, abstract.py:2243-2273) and `state` is
                                  UNINITIALISED memory: the first sweep makes its tiles instead of reading them, so the 2^n-amplitude fill
                                  and one read pass disappear (a program without a sweep fills `state` first) */
#define QB_PROGRAM_PERM_FUSED_ONLY 4 /* *_permuted calls: QB_ERR_UNSUPPORTED (nothing launched) when the permutation cannot ride on a sweep */
int qb_apply_program(qb_handle h, void* state, int nqubits, int dtype, const qb_op* ops, int nops, int flags,
                     qb_program_stats* stats /* may be NULL */);
/* Compiled programs: the same gate queue planned ONCE, its sweep programs kept resident in device memory, and launched
 * as often as needed without host planning or a program upload -- what repeated execution asks for (one simulation per
 * shot in Backend.execute_circuit_repeated, abstract.py:2532-2636; re-execution of a circuit, models/circuit.py:1071-1110;
 * every step of a distributed run).  `qb_program_run` applies it to any state of the nqubits/dtype it was compiled for,
 * on the stream of `h`; flags: QB_PROGRAM_TIME.  Destroy with the handle it was created on. */
typedef struct qb_program_s* qb_program;
int qb_program_create(qb_handle h, int nqubits, int dtype, const qb_op* ops, int nops, int flags, qb_program* out,
                      qb_program_stats* stats /* may be NULL */);
int qb_program_run(qb_handle h, qb_program program, void* state, int flags, qb_program_stats* stats /* may be NULL */);
/* Parameter slots (Circuit.set_parameters + re-execution, models/circuit.py:788-857; parallel.py:124-192): new matrices
 * for some ops of a compiled program.  The schedule is kept -- only the sweep programs are emitted again and re-uploaded in
 * stream order; if the new values change the gate structure (an angle of 0 turns a rotation into an identity) the program
 * is planned from scratch.  `family` != QB_GATE_MATRIX: the matrix is evaluated here from `theta` as
 * backends/npmatrices.py does (RX :79, RY :84, RZ :89, U1, CU1 :230, CRX, CRY, CRZ), so a variational loop sends angles. */
#define QB_GATE_MATRIX 0
#define QB_GATE_RX 1
#define QB_GATE_RY 2
#define QB_GATE_RZ 3
#define QB_GATE_U1 4
#define QB_GATE_CRX 5 /* two targets: (control, target), the full 4x4 matrix as the reference applies it */
#define QB_GATE_CRY 6
#define QB_GATE_CRZ 7
#define QB_GATE_CU1 8
typedef struct {
  int32_t op_index;     /* index into the ops the program was created from */
  int32_t family;       /* QB_GATE_* */
  int32_t conjugate;    /* 1: complex-conjugate the matrix (the column side of a density-matrix gate) */
  int32_t reserved;
  double theta[3];
  const double* matrix; /* QB_GATE_MATRIX: host, interleaved complex128, the shape of the op's original data */
} qb_param_update;
int qb_program_set_params(qb_handle h, qb_program p, const qb_param_update* updates, int nupdates);
int qb_program_destroy(qb_handle h, qb_program program);
/* A gate queue FOLLOWED BY A QUBIT PERMUTATION (the run of SWAP gates ending models/qft.py:55-57; the closing local
 * permutation of a distributed run), out of place: the ops are applied to `state` and the result, with qubit q moved to
 * qubit dest_of_qubit[q], is written to `dst` (another buffer of the same size; `state` is left holding an intermediate
 * state).  When the permutation's tile can be a sweep's tile, the LAST sweep writes its tiles permuted through a second
 * tensor map -- no separate K8 pass over the state (stats->perm_fused = 1, nsweeps counts only sweep launches); otherwise
 * the ops run in place and K8 follows (perm_fused = 0, nsweeps includes the K8 launch). */
int qb_apply_program_permuted(qb_handle h, void* state, void* dst, int nqubits, int dtype, const qb_op* ops, int nops,
                              const int* dest_of_qubit, int flags, qb_program_stats* stats /* may be NULL */);
int qb_program_create_permuted(qb_handle h, int nqubits, int dtype, const qb_op* ops, int nops, const int* dest_of_qubit,
                               int flags, qb_program* out, qb_program_stats* stats /* may be NULL */);
/* runs a program made by qb_program_create_permuted (qb_program_run refuses those: they need `dst`) */
int qb_program_run_permuted(qb_handle h, qb_program program, void* state, void* dst, int flags,
                            qb_program_stats* stats /* may be NULL */);
/* host-only: run the sweep planner without touching a device (used by the CPU test-suite) */
int qb_plan_program(int nqubits, int dtype, const qb_op* ops, int nops, int flags, qb_program_stats* stats,
                    int32_t* sweep_of_op /* nops entries, may be NULL */);
int qb_plan_program_permuted(int nqubits, int dtype, const qb_op* ops, int nops, const int* dest_of_qubit, int flags,
                             qb_program_stats* stats, int32_t* sweep_of_op /* nops entries, may be NULL */);

/* ---- K8: qubit permutation, out of place, one sweep (a run of SWAP gates such as the bit reversal ending
 * models/qft.py:55-57; gates/gates.py:1669 SWAP).  dst[.. qubit dest_of_qubit[q] ..] = src[.. qubit q ..];
 * src and dst must not overlap. */
int qb_permute_qubits(qb_handle h, const void* src, void* dst, int nqubits, int dtype, const int* dest_of_qubit);

/* ---- K3: probabilities and marginals (abstract.py:2734-2758 + _order_probabilities :3371-3381) ----
 * out[idx] with idx bits ordered as `qubits` (caller order, qubits[0] = MSB); out is a DEVICE buffer of
 * 2^nmeasured reals of the state's real dtype (float for complex64, double for complex128). */
int qb_probabilities(qb_handle h, const void* state, int nqubits, int dtype, const int* qubits, int nmeasured,
                     void* probs_out);

/* ---- K4: shot sampling (abstract.py:2774-2781 sample_shots -> np.random.choice: sequential float64
 * cumsum, divide by last, searchsorted side="right") ---------------------------------------------------
 * mode 0: numpy-exact sequential-order scan (bit-identical CDF to np.cumsum);
 * mode 1: parallel three-phase scan (deterministic, rounding differs from np.cumsum at ~1e-16). */
#define QB_SCAN_EXACT 0
#define QB_SCAN_PARALLEL 1
int qb_cdf(qb_handle h, const void* probs, int rdtype, uint64_t nbins, double* cdf_out, int mode);
/* idx = #{k : cdf[k] <= u}; uniforms and out are DEVICE buffers */
int qb_sample_cdf(qb_handle h, const double* cdf, uint64_t nbins, const double* uniforms, uint64_t nshots,
                  int64_t* out);
/* convenience with HOST uniforms/out (the plugin call): scan + search, scratch CDF owned by the context.
 * total_out (may be NULL) receives sum(probs) before normalisation -- np.random.choice validates it. */
int qb_sample(qb_handle h, const void* probs, int rdtype, uint64_t nbins, const double* uniforms_host,
              uint64_t nshots, int64_t* out_host, int mode, double* total_out);

/* ---- K5: collapse (abstract.py:2424-2440 collapse_state -> _collapse_statevector :3279-3304) ------
 * `qubits` sorted ascending; `outcome` is the decimal of the measured bits, qubits[0] = MSB. */
int qb_collapse(qb_handle h, void* state, int nqubits, int dtype, const int* qubits, int nmeasured,
                uint64_t outcome, int normalize);

/* Density matrices (rho: flat 2^n x 2^n row-major array on the device).  qb_probabilities_dm replaces the density-matrix
 * branch of Backend.calculate_probabilities (backends/abstract.py:2741-2749): |sum of the diagonal over the unmeasured
 * qubits|, bins in the caller's qubit order, real dtype of the state.  qb_collapse_dm replaces
 * Backend._collapse_density_matrix (abstract.py:3249-3277): rows and columns projected on `outcome`, divided by the trace. */
int qb_probabilities_dm(qb_handle h, const void* rho, int nqubits, int dtype, const int* qubits, int nmeasured, void* probs_out);
int qb_collapse_dm(qb_handle h, void* rho, int nqubits, int dtype, const int* qubits, int nmeasured, uint64_t outcome, int normalize);

/* ---- K9: expectation values without leaving the device (abstract.py:2946-3054 exp_value_observable_symbolic: the
 * einsum contraction of one Pauli term with the state; abstract.py:2180-2190 overlap_statevector) --------------------
 * <psi| P |psi> for the Pauli string `paulis` (characters I, X, Y, Z; paulis[i] acts on qubits[i]): one read pass over
 * the state, no copy.  out_host receives (re, im); the imaginary part vanishes up to rounding for a normalised state. */
int qb_expval_pauli(qb_handle h, const void* state, int nqubits, int dtype, const char* paulis, const int* qubits,
                    int nterm_qubits, double* out_host);
/* <a|b> = sum conj(a) b of two buffers of 2^nqubits amplitudes; out_host receives (re, im). */
int qb_state_vdot(qb_handle h, const void* a, const void* b, int nqubits, int dtype, double* out_host);

/* ---- K7: global<->local qubit exchange for the distributed scheme (models/distcircuit.py; the
 * executor Backend.execute_distributed_circuit is NotImplemented in-tree, abstract.py:2638-2647) -----
 * Copies the half of the shard with local qubit `local_qubit` == `bit` into/out of a contiguous
 * staging buffer of 2^(nqubits-1) amplitudes (pack: state -> staging; unpack: staging -> state). */
int qb_pack_half(qb_handle h, const void* state, int nqubits, int dtype, int local_qubit, int bit, void* staging);
int qb_unpack_half(qb_handle h, void* state, int nqubits, int dtype, int local_qubit, int bit, const void* staging);
/* Exchange over NVLink peer memory, no staging and no NCCL on the data path: ONE kernel swaps, element by
 * element in registers, this rank's outgoing half (local qubit == 1 - my_bit) with the partner's outgoing half
 * (local qubit == my_bit), reading and writing the partner's shard through a peer-mapped pointer
 * (`peer_state`, from qb_ipc_open_handle).  The two ranks of a pair split the index range (`part` of `nparts`),
 * so both NVLink directions carry loads and stores at once.  The caller brackets it with a stream-ordered
 * barrier on both ranks. */
int qb_swap_half_p2p(qb_handle h, void* state, void* peer_state, int nqubits, int dtype, int local_qubit, int my_bit,
                     int part, int nparts);
/* Several exchanges at once (k global qubits <-> the k leading local qubits): an all-to-all of contiguous chunks in ONE
 * kernel.  Entry i swaps state[my_offsets[i] + e] with peer_states[i][peer_offsets[i] + e] for e in [begins[i], ends[i])
 * (offsets and ranges in amplitudes; the two ranks of a pair take disjoint ranges).  1..8 entries.  Replaces k calls of
 * qb_swap_half_p2p: (2^k - 1) / 2^k of a shard crosses NVLink instead of k / 2 (distcircuit.py:194-196 swaps one pair of
 * qubits at a time).  Same stream-ordered barriers around it. */
int qb_alltoall_p2p(qb_handle h, void* state, int dtype, int npeers, void* const* peer_states, const uint64_t* my_offsets,
                    const uint64_t* peer_offsets, const uint64_t* begins, const uint64_t* ends);
/* Out-of-place form of the same all-to-all: entry i copies state[my_offsets[i] + e] to dest_buffers[i][dest_offsets[i] + e]
 * for e in [begins[i], ends[i]) -- the destination ranks' SECOND buffers (this rank's own for the chunk that stays).
 * Remote stores only; afterwards every rank continues in its second buffer. */
int qb_alltoall_push_p2p(qb_handle h, const void* state, int dtype, int nentries, void* const* dest_buffers,
                         const uint64_t* my_offsets, const uint64_t* dest_offsets, const uint64_t* begins, const uint64_t* ends);
/* CUDA IPC plumbing so that ranks (one process per GPU) can map each other's buffers */
int qb_ipc_get_handle(qb_handle h, void* dptr, void* handle_out_64bytes);
int qb_ipc_open_handle(qb_handle h, const void* handle_64bytes, void** dptr_out);
int qb_ipc_close_handle(qb_handle h, void* dptr);

#ifdef __cplusplus
}
#endif
#endif /* QIBO_B200_H */
