"""qibo-free descriptions of the BASELINE.json circuits as :class:`qibo_b200.ops.Op` lists.

Used by bench.py / smoke() / the GPU tests on the box, where the reference package may be absent.
Gate matrices follow backends/npmatrices.py (H :27, RX :79, RY :84, RZ :89, CNOT :144, CZ :162,
CU1 :230, SWAP :265); circuit structure follows models/qft.py:47-58,
examples/benchmarks/circuits.py:7-22 and tests/test_models_circuit_fuse.py:124-138.
"""

import math

import numpy as np

from qibo_b200.ops import Op

_S2 = math.sqrt(2.0)


def matrix(name, *params):
    if name == "H":
        return np.array([[1, 1], [1, -1]], dtype=np.complex128) / _S2
    if name == "X":
        return np.array([[0, 1], [1, 0]], dtype=np.complex128)
    if name == "Z":
        return np.array([[1, 0], [0, -1]], dtype=np.complex128)
    if name == "RX":
        c, s = math.cos(params[0] / 2.0), math.sin(params[0] / 2.0)
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)
    if name == "RY":
        c, s = math.cos(params[0] / 2.0), math.sin(params[0] / 2.0)
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    if name == "RZ":
        ph = np.exp(0.5j * params[0])
        return np.array([[np.conj(ph), 0], [0, ph]], dtype=np.complex128)
    if name == "CNOT":
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)
    if name == "CZ":
        return np.diag([1, 1, 1, -1]).astype(np.complex128)
    if name == "CU1":
        return np.diag([1, 1, 1, np.exp(1j * params[0])]).astype(np.complex128)
    if name == "SWAP":
        return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
    raise KeyError(name)


def op(name, qubits, *params):
    """Named gate over ``gate.qubits`` (sorted controls + targets) with its full matrix, as the reference applies it."""
    return Op(matrix(name, *params), tuple(qubits), name=name)


def qft(nqubits, with_swaps=True):
    ops = []
    for i1 in range(nqubits):
        ops.append(op("H", (i1,)))
        for i2 in range(i1 + 1, nqubits):
            ops.append(op("CU1", (i2, i1), math.pi / 2 ** (i2 - i1)))
    if with_swaps:
        for q in range(nqubits // 2):
            ops.append(op("SWAP", (q, nqubits - q - 1)))
    return ops


def variational(nqubits, nlayers, thetas):
    theta = iter(thetas)
    ops = []
    for _ in range(nlayers):
        for i in range(nqubits):
            ops.append(op("RY", (i,), float(next(theta))))
        for i in range(0, nqubits - 1, 2):
            ops.append(op("CZ", (i, i + 1)))
        for i in range(nqubits):
            ops.append(op("RY", (i,), float(next(theta))))
        for i in range(1, nqubits - 2, 2):
            ops.append(op("CZ", (i, i + 1)))
        ops.append(op("CZ", (0, nqubits - 1)))
    return ops


def random_circuit(nqubits, ngates, seed):
    np.random.seed(seed)
    one, two = ["RX", "RY", "RZ"], ["CNOT", "CZ", "SWAP"]
    thetas = np.pi * np.random.random((ngates,))
    ops = []
    for i in range(ngates):
        g = one[int(np.random.randint(0, 3))]
        q0 = int(np.random.randint(0, nqubits))
        ops.append(op(g, (q0,), float(thetas[i])))
        g = two[int(np.random.randint(0, 3))]
        q0, q1 = np.random.randint(0, nqubits, (2,))
        while q0 == q1:
            q0, q1 = np.random.randint(0, nqubits, (2,))
        ops.append(op(g, (int(q0), int(q1))))
    return ops
