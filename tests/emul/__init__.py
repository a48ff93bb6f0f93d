"""TEST INFRASTRUCTURE: builds and loads the CPU emulation of the sweep data path (tests/emul/emul.cpp)."""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_emul.so")
SRC = os.path.join(HERE, "emul.cpp")
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "qibo_b200", "csrc")


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    if _stale():
        subprocess.run(
            ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", SO, SRC], check=True
        )
    from qibo_b200 import _lib

    lib = ctypes.CDLL(SO)
    lib.emul_apply_program.restype = ctypes.c_int
    lib.emul_apply_program.argtypes = [
        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_lib.QbOp), ctypes.c_int, ctypes.c_int,
        ctypes.POINTER(_lib.QbProgramStats),
    ]
    lib.emul_last_error.restype = ctypes.c_char_p
    lib.emul_canon_kind.restype = ctypes.c_int
    lib.emul_canon_kind.argtypes = [
        ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
    ]
    return lib


def apply_program(state, nqubits, ops, fuse=True):
    """Run ``ops`` (qibo_b200.ops.Op list) through the emulated sweep path, in place on a copy; -> (state, stats)."""
    from qibo_b200 import _lib
    from qibo_b200.ops import pack_ops

    lib = load()
    state = np.ascontiguousarray(state).copy()
    dtype = _lib.QB_C128 if state.dtype == np.complex128 else _lib.QB_C64
    arr, keep = pack_ops(ops)
    stats = _lib.QbProgramStats()
    rc = lib.emul_apply_program(
        state.ctypes.data, nqubits, dtype, arr, len(ops), 0 if fuse else _lib.QB_PROGRAM_NO_FUSE, ctypes.byref(stats)
    )
    if rc != 0:
        raise RuntimeError(lib.emul_last_error().decode())
    del keep
    return state, stats


def canon_kind(nqubits, op):
    lib = load()
    kind, nt, nc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    t = (ctypes.c_int * max(1, len(op.targets)))(*op.targets)
    c = (ctypes.c_int * max(1, len(op.controls)))(*op.controls)
    rc = lib.emul_canon_kind(
        nqubits, op.data.ctypes.data, int(op.is_diagonal), len(op.targets), t, len(op.controls), c,
        ctypes.byref(kind), ctypes.byref(nt), ctypes.byref(nc),
    )
    if rc != 0:
        raise ValueError(lib.emul_last_error().decode())
    return kind.value, nt.value, nc.value


def permute_qubits(state, nqubits, dest_of_qubit):
    from qibo_b200 import _lib

    lib = load()
    src = np.ascontiguousarray(state)
    dst = np.empty_like(src)
    arr = (ctypes.c_int * nqubits)(*[int(d) for d in dest_of_qubit])
    lib.emul_permute_qubits.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    rc = lib.emul_permute_qubits(src.ctypes.data, dst.ctypes.data, nqubits, _lib.QB_C128 if src.dtype == np.complex128 else _lib.QB_C64, arr)
    assert rc == 0
    return dst


def apply_program_replay(state, nqubits, ops_old, ops_new, fuse=True):
    """Plan ``ops_old``, emit ``ops_new`` on that schedule (qb_program_set_params' path) and run it -> (state, replayed)."""
    from qibo_b200 import _lib
    from qibo_b200.ops import pack_ops

    lib = load()
    state = np.ascontiguousarray(state).copy()
    dtype = _lib.QB_C128 if state.dtype == np.complex128 else _lib.QB_C64
    a, keep_a = pack_ops(ops_old)
    b, keep_b = pack_ops(ops_new)
    replayed = ctypes.c_int()
    lib.emul_apply_program_replay.restype = ctypes.c_int
    lib.emul_apply_program_replay.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_lib.QbOp), ctypes.POINTER(_lib.QbOp),
                                              ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    rc = lib.emul_apply_program_replay(state.ctypes.data, nqubits, dtype, a, b, len(ops_old), 0 if fuse else _lib.QB_PROGRAM_NO_FUSE,
                                       ctypes.byref(replayed))
    if rc != 0:
        raise RuntimeError(lib.emul_last_error().decode())
    del keep_a, keep_b
    return state, bool(replayed.value)


def family_matrix(family, thetas, ntargets, is_diagonal=False):
    lib = load()
    th = (ctypes.c_double * 3)(*(list(thetas) + [0.0, 0.0, 0.0])[:3])
    out = (ctypes.c_double * 64)()
    count = ctypes.c_int()
    lib.emul_family_matrix.restype = ctypes.c_int
    lib.emul_family_matrix.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                       ctypes.POINTER(ctypes.c_int)]
    rc = lib.emul_family_matrix(family, th, ntargets, int(is_diagonal), out, ctypes.byref(count))
    if rc != 0:
        raise ValueError("family does not match")
    v = np.array(out[: count.value]).view(np.complex128)
    return v if is_diagonal else v.reshape(1 << ntargets, 1 << ntargets)


def apply_program_permuted(state, nqubits, ops, dest_of_qubit, fuse=True, replay=False):
    """``ops`` then the qubit permutation (qubit q -> dest_of_qubit[q]) through the emulated sweep path with the
    permutation riding on the last sweep when the planner can fuse it -> (result, stats, fused)."""
    from qibo_b200 import _lib
    from qibo_b200.ops import pack_ops

    lib = load()
    state = np.ascontiguousarray(state).copy()
    dst = np.full_like(state, np.nan)
    dtype = _lib.QB_C128 if state.dtype == np.complex128 else _lib.QB_C64
    arr, keep = pack_ops(ops)
    stats = _lib.QbProgramStats()
    fused = ctypes.c_int()
    dest = (ctypes.c_int * nqubits)(*[int(d) for d in dest_of_qubit])
    lib.emul_apply_program_permuted.restype = ctypes.c_int
    lib.emul_apply_program_permuted.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_lib.QbOp), ctypes.c_int,
        ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(_lib.QbProgramStats), ctypes.POINTER(ctypes.c_int), ctypes.c_int,
    ]
    rc = lib.emul_apply_program_permuted(
        state.ctypes.data, dst.ctypes.data, nqubits, dtype, arr, len(ops), dest, 0 if fuse else _lib.QB_PROGRAM_NO_FUSE,
        ctypes.byref(stats), ctypes.byref(fused), int(replay),
    )
    if rc != 0:
        raise RuntimeError(lib.emul_last_error().decode())
    del keep
    return dst, stats, bool(fused.value)
