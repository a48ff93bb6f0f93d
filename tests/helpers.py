"""Shared helpers for the parity tests (oracle side + product op construction)."""

import numpy as np

from oracle import numpy_oracle as orc
from qibo_b200.ops import Op


def rand_state(n, seed, dtype="complex128"):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    x /= np.linalg.norm(x)
    return x.astype(dtype)


def rand_unitary(k, rng):
    a = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def oracle_apply(state, op: Op, nqubits):
    """Apply a product Op with the oracle (reference algorithm)."""
    dtype = state.dtype
    data = op.data.astype(dtype)
    mat = np.diag(data) if op.is_diagonal else data
    if op.controls:
        return orc.apply_gate_controlled_by(state, mat, sorted(op.controls), list(op.targets), nqubits)
    return orc.apply_gate(state, mat, list(op.targets), nqubits)


def oracle_run(state, ops, nqubits):
    for op in ops:
        state = oracle_apply(state, op, nqubits)
    return state


def ops_from_named(named):
    """oracle op tuples (name, qubits, params) -> product Ops with the oracle's matrices."""
    return [Op(orc.gate_matrix(name, *params), tuple(qubits), name=name) for name, qubits, params in named]


def random_zoo(nqubits, ngates, seed, max_dense=5):
    """A random mix of everything the planner distinguishes: dense k<=max_dense with/without controls, diagonals,
    phases, swaps, named controlled gates."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(ngates):
        kind = rng.integers(0, 9)
        if kind == 0:  # dense unitary, user-ordered targets, maybe controls
            k = int(rng.integers(1, max_dense + 1))
            nc = int(rng.integers(0, 3))
            qs = rng.permutation(nqubits)[: k + nc].tolist()
            ops.append(Op(rand_unitary(k, rng), tuple(qs[:k]), tuple(qs[k:])))
        elif kind == 1:  # general diagonal on k qubits
            k = int(rng.integers(1, 5))
            qs = rng.permutation(nqubits)[:k].tolist()
            d = np.exp(1j * rng.uniform(0, 2 * np.pi, size=2**k))
            ops.append(Op(d, tuple(qs), is_diagonal=True))
        elif kind == 2:  # controlled phase (CU1-like), full matrix form
            a, b = rng.permutation(nqubits)[:2].tolist()
            ops.append(Op(orc.gate_matrix("CU1", float(rng.uniform(0, 6))), (a, b)))
        elif kind == 3:
            a, b = rng.permutation(nqubits)[:2].tolist()
            nc = int(rng.integers(0, 2))
            cs = [q for q in rng.permutation(nqubits).tolist() if q not in (a, b)][:nc]
            ops.append(Op(orc.gate_matrix("SWAP"), (a, b), tuple(cs)))
        elif kind == 4:
            a, b, c = rng.permutation(nqubits)[:3].tolist()
            ops.append(Op(orc.gate_matrix("TOFFOLI"), (a, b, c)))
        elif kind == 5:
            a, b = rng.permutation(nqubits)[:2].tolist()
            name = ["CNOT", "CZ", "RZZ", "fSim", "RXX", "iSWAP"][int(rng.integers(0, 6))]
            params = {"RZZ": (0.3,), "fSim": (0.4, 0.9), "RXX": (1.1,)}.get(name, ())
            ops.append(Op(orc.gate_matrix(name, *params), (a, b)))
        elif kind == 6:
            q = int(rng.integers(0, nqubits))
            name = ["H", "X", "Y", "Z", "S", "T", "RX", "RY", "RZ", "U1"][int(rng.integers(0, 10))]
            params = (float(rng.uniform(0, 6)),) if name in ("RX", "RY", "RZ", "U1") else ()
            ops.append(Op(orc.gate_matrix(name, *params), (q,)))
        elif kind == 7:  # multi-controlled phase / Z
            k = int(rng.integers(2, 5))
            qs = rng.permutation(nqubits)[:k].tolist()
            ops.append(Op(orc.gate_matrix("U1", float(rng.uniform(0, 6))), (qs[0],), tuple(qs[1:])))
        else:  # controlled RY
            k = int(rng.integers(1, 4))
            qs = rng.permutation(nqubits)[: k + 1].tolist()
            ops.append(Op(orc.gate_matrix("RY", float(rng.uniform(0, 6))), (qs[0],), tuple(qs[1:])))
    return ops
