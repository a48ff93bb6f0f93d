"""Gate descriptions handed to the C ABI (`qb_op`), independent of qibo.

An :class:`Op` is "matrix (or diagonal) on ``targets`` where all ``controls`` are 1", in Qibo's
conventions: qubit 0 is the most significant bit of the state index, ``targets[0]`` is the most
significant bit of the matrix index (gates/abstract.py:439-442).
"""

import ctypes
from dataclasses import dataclass, field
from typing import Sequence, Tuple

import numpy as np

from qibo_b200 import _lib


def is_wide(op) -> bool:
    """More targets than a sweep tile pass takes (Unitary / FusedGate / I on 7+ qubits, gates/gates.py:2774)."""
    return len(op.targets) > _lib.QB_MAX_OP_TARGETS


@dataclass
class Op:
    data: np.ndarray  # (2^k, 2^k) matrix or (2^k,) diagonal
    targets: Tuple[int, ...]
    controls: Tuple[int, ...] = ()
    is_diagonal: bool = False
    name: str = ""

    def __post_init__(self):
        k = len(self.targets)
        self.data = np.ascontiguousarray(self.data, dtype=np.complex128)
        want = (2**k,) if self.is_diagonal else (2**k, 2**k)
        if self.data.shape != want:
            raise ValueError(f"gate data has shape {self.data.shape}, expected {want} for {k} target qubits")
        if len(self.controls) > _lib.QB_MAX_OP_CONTROLS:
            raise NotImplementedError("too many control qubits")
        self.targets = tuple(int(q) for q in self.targets)
        self.controls = tuple(int(q) for q in self.controls)


def pack_ops(ops: Sequence[Op]):
    """-> (ctypes array of qb_op, keep-alive list).  Matrices are passed by host pointer."""
    arr = (_lib.QbOp * max(len(ops), 1))()
    for i, op in enumerate(ops):
        if len(op.targets) > _lib.QB_MAX_OP_TARGETS:  # (Engine routes wider blocks through permute + GEMM, see apply_wide)
            raise NotImplementedError(f"a sweep program takes gates on at most {_lib.QB_MAX_OP_TARGETS} target qubits")
        c = arr[i]
        c.ntargets = len(op.targets)
        c.ncontrols = len(op.controls)
        for j, q in enumerate(op.targets):
            c.targets[j] = q
        for j, q in enumerate(op.controls):
            c.controls[j] = q
        c.is_diagonal = 1 if op.is_diagonal else 0
        c.data = op.data.__array_interface__["data"][0]  # (the address; `.ctypes.data` builds a ctypes object per call)
    return arr, [op.data for op in ops]
