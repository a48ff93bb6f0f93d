"""ctypes binding of libqibo_b200.so (the C ABI declared in include/qibo_b200.h).

The library is the product: there is no Python/NumPy/torch implementation of the hot path behind it.
If the shared object is missing the import fails loudly (build it with ``python -m qibo_b200.build``).
"""

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QB_LIB_PATH") or os.path.join(HERE, "lib", "libqibo_b200.so")

QB_C64, QB_C128 = 0, 1
QB_F32, QB_F64 = 0, 1
QB_OK, QB_ERR_INVALID, QB_ERR_OOM, QB_ERR_CUDA, QB_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
QB_MAX_OP_TARGETS, QB_MAX_OP_CONTROLS = 6, 32
QB_PROGRAM_TIME, QB_PROGRAM_NO_FUSE, QB_PROGRAM_PERM_FUSED_ONLY, QB_PROGRAM_INPUT_ZERO = 1, 2, 4, 8
QB_SCAN_EXACT, QB_SCAN_PARALLEL = 0, 1


class QbOp(ctypes.Structure):
    _fields_ = [
        ("ntargets", c_int32),
        ("ncontrols", c_int32),
        ("targets", c_int32 * QB_MAX_OP_TARGETS),
        ("controls", c_int32 * QB_MAX_OP_CONTROLS),
        ("is_diagonal", c_int32),
        ("reserved", c_int32),
        ("data", c_void_p),
    ]


class QbProgramStats(ctypes.Structure):
    _fields_ = [
        ("nops", c_int32),
        ("nsweeps", c_int32),
        ("ndense_passes", c_int32),
        ("ndiag_ops", c_int32),
        ("bytes_moved", c_double),
        ("elapsed_ms", c_float),
        ("nstage_sweeps", c_int32),
        ("perm_fused", c_int32),
        ("reserved", c_int32),
    ]


class QbParamUpdate(ctypes.Structure):
    _fields_ = [
        ("op_index", c_int32),
        ("family", c_int32),
        ("conjugate", c_int32),
        ("reserved", c_int32),
        ("theta", c_double * 3),
        ("matrix", c_void_p),
    ]


QB_GATE_MATRIX, QB_GATE_RX, QB_GATE_RY, QB_GATE_RZ, QB_GATE_U1, QB_GATE_CRX, QB_GATE_CRY, QB_GATE_CRZ, QB_GATE_CU1 = range(9)

# name -> (restype, argtypes); every symbol include/qibo_b200.h declares
PROTOTYPES = {
    "qb_version": (c_int, []),
    "qb_last_error": (c_char_p, []),
    "qb_device_count": (c_int, [POINTER(c_int)]),
    "qb_create": (c_int, [c_int, c_void_p, POINTER(c_void_p)]),
    "qb_destroy": (c_int, [c_void_p]),
    "qb_set_stream": (c_int, [c_void_p, c_void_p]),
    "qb_sync": (c_int, [c_void_p]),
    "qb_mem_info": (c_int, [c_void_p, POINTER(c_size_t), POINTER(c_size_t)]),
    "qb_malloc": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "qb_free": (c_int, [c_void_p, c_void_p]),
    "qb_memcpy": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
    "qb_memcpy_async": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "qb_state_set_basis": (c_int, [c_void_p, c_void_p, c_int, c_int, c_uint64]),
    "qb_state_fill": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_double]),
    "qb_state_cast": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_uint64]),
    "qb_state_norm2": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_double)]),
    "qb_apply_matrix": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, POINTER(c_int), c_int, POINTER(c_int)]),
    "qb_apply_diagonal": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, POINTER(c_int), c_int, POINTER(c_int)]),
    "qb_apply_program": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(QbOp), c_int, c_int, POINTER(QbProgramStats)]),
    "qb_expval_pauli": (c_int, [c_void_p, c_void_p, c_int, c_int, c_char_p, POINTER(c_int), c_int, POINTER(c_double)]),
    "qb_state_vdot": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(c_double)]),
    "qb_program_create": (c_int, [c_void_p, c_int, c_int, POINTER(QbOp), c_int, c_int, POINTER(c_void_p), POINTER(QbProgramStats)]),
    "qb_program_run": (c_int, [c_void_p, c_void_p, c_void_p, c_int, POINTER(QbProgramStats)]),
    "qb_program_destroy": (c_int, [c_void_p, c_void_p]),
    "qb_program_set_params": (c_int, [c_void_p, c_void_p, POINTER(QbParamUpdate), c_int]),
    "qb_plan_program": (c_int, [c_int, c_int, POINTER(QbOp), c_int, c_int, POINTER(QbProgramStats), POINTER(c_int32)]),
    "qb_apply_program_permuted": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(QbOp), c_int, POINTER(c_int), c_int, POINTER(QbProgramStats)]),
    "qb_program_create_permuted": (c_int, [c_void_p, c_int, c_int, POINTER(QbOp), c_int, POINTER(c_int), c_int, POINTER(c_void_p), POINTER(QbProgramStats)]),
    "qb_program_run_permuted": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, POINTER(QbProgramStats)]),
    "qb_plan_program_permuted": (c_int, [c_int, c_int, POINTER(QbOp), c_int, POINTER(c_int), c_int, POINTER(QbProgramStats), POINTER(c_int32)]),
    "qb_permute_qubits": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(c_int)]),
    "qb_probabilities": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_int), c_int, c_void_p]),
    "qb_cdf": (c_int, [c_void_p, c_void_p, c_int, c_uint64, c_void_p, c_int]),
    "qb_sample_cdf": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_uint64, c_void_p]),
    "qb_sample": (c_int, [c_void_p, c_void_p, c_int, c_uint64, c_void_p, c_uint64, c_void_p, c_int, POINTER(c_double)]),
    "qb_collapse": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_int), c_int, c_uint64, c_int]),
    "qb_probabilities_dm": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_int), c_int, c_void_p]),
    "qb_collapse_dm": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_int), c_int, c_uint64, c_int]),
    "qb_pack_half": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "qb_unpack_half": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "qb_swap_half_p2p": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int]),
    "qb_alltoall_p2p": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p), POINTER(c_uint64), POINTER(c_uint64), POINTER(c_uint64), POINTER(c_uint64)]),
    "qb_alltoall_push_p2p": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p), POINTER(c_uint64), POINTER(c_uint64), POINTER(c_uint64), POINTER(c_uint64)]),
    "qb_ipc_get_handle": (c_int, [c_void_p, c_void_p, c_void_p]),
    "qb_ipc_open_handle": (c_int, [c_void_p, c_void_p, POINTER(c_void_p)]),
    "qb_ipc_close_handle": (c_int, [c_void_p, c_void_p]),
}


class QiboB200Error(RuntimeError):
    """CUDA / library failure that is not an out-of-memory or an argument error."""


class QiboB200OutOfMemory(MemoryError):
    """cudaErrorMemoryAllocation inside the library (Backend.oom_error)."""


_lib = None


def load():
    """Load the shared library (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"qibo_b200: {LIB_PATH} is missing. The CUDA library is the product (no CPU fallback); "
            "build it with `python -m qibo_b200.build`."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    return load().qb_last_error().decode("utf-8", "replace")


def check(rc):
    """Map a QB_ERR_* status to the exception class the reference's callers expect."""
    if rc == QB_OK:
        return
    msg = last_error()
    if rc == QB_ERR_INVALID:
        raise ValueError(msg)
    if rc == QB_ERR_OOM:
        raise QiboB200OutOfMemory(msg)
    if rc == QB_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise QiboB200Error(msg)
