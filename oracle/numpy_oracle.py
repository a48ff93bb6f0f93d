"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference NumpyBackend hot path.

Every function cites the reference lines it follows (paths relative to /root/reference/src/qibo).
The formulation deliberately mirrors the reference's *algorithm* (axis transposition + reshape +
matmul, sequential cumsum + searchsorted, ...), so that rounding behaviour is the same; it is the
checker for the CUDA path, never part of it.
"""

import cmath
import math
from collections import Counter

import numpy as np

SHOT_BATCH_SIZE = 2**18  # config.py:41

# --------------------------------------------------------------------------------------
# G3: gate matrices (backends/npmatrices.py).  Evaluated in Python double, then cast to the
# backend dtype *before* the multiply (npmatrices.py:21-24).
# --------------------------------------------------------------------------------------


def gate_matrix(name, *params, dtype="complex128"):
    """Matrix over ``gate.qubits = sorted(controls) + targets`` (gates/abstract.py:439-442)."""
    s2 = math.sqrt(2)
    if name == "H":  # npmatrices.py:27-28
        m = np.array([[1, 1], [1, -1]], dtype=dtype) / s2
        return m
    if name == "X":  # :31
        m = [[0, 1], [1, 0]]
    elif name == "Y":  # :35
        m = [[0j, -1j], [1j, 0j]]
    elif name == "Z":  # :39
        m = [[1, 0], [0, -1]]
    elif name == "S":  # :51
        m = [[1 + 0j, 0j], [0j, 1j]]
    elif name == "SDG":
        m = [[1 + 0j, 0j], [0j, -1j]]
    elif name == "T":  # :59
        m = [[1 + 0j, 0], [0, cmath.exp(1j * math.pi / 4.0)]]
    elif name == "TDG":
        m = [[1 + 0j, 0], [0, cmath.exp(-1j * math.pi / 4.0)]]
    elif name == "RX":  # :79-82
        (theta,) = params
        cos = np.cos(theta / 2.0) + 0j
        isin = -1j * np.sin(theta / 2.0)
        m = [[cos, isin], [isin, cos]]
    elif name == "RY":  # :84-87
        (theta,) = params
        cos = np.cos(theta / 2.0) + 0j
        sin = np.sin(theta / 2.0) + 0j
        m = [[cos, -sin], [sin, cos]]
    elif name == "RZ":  # :89-91
        (theta,) = params
        phase = np.exp(0.5j * theta)
        m = [[np.conj(phase), 0], [0, phase]]
    elif name == "U1":  # :116-118
        (theta,) = params
        m = [[1, 0], [0, np.exp(1j * theta)]]
    elif name == "U3":  # :125-141
        theta, phi, lam = params
        cost = np.cos(theta / 2)
        sint = np.sin(theta / 2)
        eplus = np.exp(1j * (phi + lam) / 2.0)
        eminus = np.exp(1j * (phi - lam) / 2.0)
        m = [
            [np.conj(eplus) * cost, -np.conj(eminus) * sint],
            [eminus * sint, eplus * cost],
        ]
    elif name == "CNOT":  # :144
        m = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]
    elif name == "CZ":  # :162
        m = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, -1]]
    elif name == "CU1":  # :230-238
        (theta,) = params
        m = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, np.exp(1j * theta)]]
    elif name == "CRX":  # :203-212
        (theta,) = params
        cos = np.cos(theta / 2.0) + 0j
        isin = -1j * np.sin(theta / 2.0)
        m = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, cos, isin], [0, 0, isin, cos]]
    elif name == "CRY":  # :214-218
        (theta,) = params
        cos = np.cos(theta / 2.0) + 0j
        sin = np.sin(theta / 2.0) + 0j
        m = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, cos, -sin], [0, 0, sin, cos]]
    elif name == "CRZ":  # :220-228
        (theta,) = params
        phase = np.exp(0.5j * theta)
        m = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, np.conj(phase), 0], [0, 0, 0, phase]]
    elif name == "SWAP":  # :265-268
        m = [[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]
    elif name == "iSWAP":  # :271
        m = [[1 + 0j, 0j, 0j, 0j], [0j, 0j, 1j, 0j], [0j, 1j, 0j, 0j], [0j, 0j, 0j, 1 + 0j]]
    elif name == "fSim":  # :312-323
        theta, phi = params
        cost = np.cos(theta) + 0j
        isint = -1j * np.sin(theta)
        phase = np.exp(-1j * phi)
        m = [
            [1 + 0j, 0j, 0j, 0j],
            [0j, cost, isint, 0j],
            [0j, isint, cost, 0j],
            [0j, 0j, 0j, phase],
        ]
    elif name == "RZZ":  # :379-390
        (theta,) = params
        phase = np.exp(0.5j * theta)
        m = np.diag([np.conj(phase), phase, phase, np.conj(phase)])
        return m.astype(dtype)
    elif name == "RXX":  # :353-364
        (theta,) = params
        cos = np.cos(theta / 2.0) + 0j
        isin = -1j * np.sin(theta / 2.0)
        m = [
            [cos, 0, 0, isin],
            [0, cos, isin, 0],
            [0, isin, cos, 0],
            [isin, 0, 0, cos],
        ]
    elif name == "TOFFOLI":  # :457-470
        m = np.eye(8)
        m[-2:, -2:] = [[0, 1], [1, 0]]
        return m.astype(dtype)
    elif name == "CCZ":  # :473
        m = np.eye(8)
        m[-1, -1] = -1
        return m.astype(dtype)
    elif name == "Unitary":  # :555-561
        (u,) = params
        return np.asarray(u).astype(dtype)
    else:  # pragma: no cover
        raise KeyError(name)
    return np.array(m, dtype=dtype)


# --------------------------------------------------------------------------------------
# G1 / G2: gate application
# --------------------------------------------------------------------------------------


def permutations(qubits, nqubits):
    """einsum_utils.py:94-110 -- move the gate's qubits to the front, and the inverse."""
    fwd = list(qubits) + [q for q in range(nqubits) if q not in qubits]
    inv = [0] * nqubits
    for pos, axis in enumerate(fwd):
        inv[axis] = pos
    return fwd, inv


def apply_gate(state, matrix, qubits, nqubits):
    """abstract.py:2322-2361, state-vector branch: transpose -> (2^k, -1) -> matmul -> back.

    ``qubits`` is ``gate.qubits`` (sorted controls, then targets); qubit 0 is the MSB of the flat index;
    ``qubits[0]`` is the MSB of the matrix row/column index.
    """
    shape = nqubits * (2,)
    state = np.reshape(state, shape)
    fwd, inv = permutations(tuple(qubits), nqubits)
    state = np.transpose(state, fwd)
    state = np.reshape(state, (2 ** len(qubits), -1))
    state = matrix @ state
    state = np.reshape(state, shape)
    state = np.transpose(state, inv)
    return np.reshape(state, (2**nqubits,))


def control_order(controls, targets, nqubits):
    """einsum_utils.py:58-71 (controls must be sorted ascending, as gate.control_qubits is)."""
    loop_start = 0
    order = list(controls)
    targets = list(targets)
    orig_targets = list(targets)
    for control in controls:
        order.extend(range(loop_start, control))
        loop_start = control + 1
        for i, t in enumerate(orig_targets):
            if t > control:
                targets[i] -= 1
    order.extend(range(loop_start, nqubits))
    return order, targets


def apply_gate_controlled_by(state, matrix, controls, targets, nqubits):
    """abstract.py:3176-3197: ``matrix`` acts on ``targets`` only where all ``controls`` are 1."""
    controls = sorted(controls)
    ncontrol = len(controls)
    nactive = nqubits - ncontrol
    state = np.reshape(state, nqubits * (2,))
    order, red_targets = control_order(controls, targets, nqubits)
    state = np.transpose(state, order)
    state = np.reshape(state, (2**ncontrol,) + nactive * (2,))
    # the einsum at abstract.py:3190 is "apply matrix on axes red_targets of state[-1]"
    updates = apply_gate(state[-1], matrix, red_targets, nactive).reshape(nactive * (2,))
    state = np.concatenate([state[:-1], updates[None]], axis=0)
    state = np.reshape(state, nqubits * (2,))
    rorder = [0] * nqubits
    for i, r in enumerate(order):
        rorder[r] = i
    state = np.transpose(state, rorder)
    return np.reshape(state, (2**nqubits,))


def apply_gate_density_matrix(rho, matrix, qubits, nqubits):
    """abstract.py:2341-2348: rho' = U rho U^dagger, two contractions (X1 in SURVEY 8a)."""
    dim = 2**nqubits
    # left: U acts on the row index; right: conj(U) acts on the column index
    vec = np.reshape(rho, (dim * dim,))
    vec = apply_gate(vec, np.conj(matrix), [q + nqubits for q in qubits], 2 * nqubits)
    vec = apply_gate(vec, matrix, list(qubits), 2 * nqubits)
    return np.reshape(vec, (dim, dim))


def matrix_fused(members, fused_qubits, dtype="complex128"):
    """abstract.py:2680-2717 -- dense matrix of a FusedGate over its sorted target set.

    ``members``: list of ``(matrix, qubits, ncontrols)`` where ``matrix`` is what ``gate.matrix(backend)``
    returns (target-only matrix when the gate came from ``controlled_by``) and ``qubits = gate.qubits``.
    """
    fused_qubits = list(fused_qubits)
    rank = len(fused_qubits)
    total = np.eye(2**rank, dtype=dtype)
    for gmatrix, qubits, ncontrols in members:
        gmatrix = np.asarray(gmatrix, dtype=dtype)
        k = len(qubits)
        if ncontrols > 0 and gmatrix.shape[0] < 2**k:  # :2691-2694 block_diag(identity, gmatrix)
            full = np.eye(2**k, dtype=dtype)
            full[-gmatrix.shape[0] :, -gmatrix.shape[0] :] = gmatrix
            gmatrix = full
        gmatrix = np.kron(gmatrix, np.eye(2 ** (rank - k), dtype=dtype))  # :2697-2698
        gmatrix = np.reshape(gmatrix, 2 * rank * (2,))
        indices = list(qubits) + [q for q in fused_qubits if q not in qubits]
        indices = [int(i) for i in np.argsort(indices)]
        tr = indices + [i + rank for i in indices]
        gmatrix = np.transpose(gmatrix, tr).reshape(2**rank, 2**rank)  # :2711-2712
        total = gmatrix @ total  # :2715
    return total


# --------------------------------------------------------------------------------------
# State constructors (abstract.py:2243-2273 zero_state; plus_state :2199-2221)
# --------------------------------------------------------------------------------------


def zero_state(nqubits, dtype="complex128"):
    state = np.zeros(2**nqubits, dtype=dtype)
    state[0] = 1
    return state


def plus_state(nqubits, dtype="complex128"):
    state = np.ones(2**nqubits, dtype=dtype)
    return state / math.sqrt(2**nqubits)


# --------------------------------------------------------------------------------------
# P1: probabilities / marginals
# --------------------------------------------------------------------------------------


def calculate_probabilities(state, qubits, nqubits):
    """abstract.py:2734-2758 (state-vector branch) + _order_probabilities :3371-3381."""
    qubits = list(qubits)
    rtype = np.real(state).dtype
    unmeasured = tuple(q for q in range(nqubits) if q not in qubits)
    probs = np.reshape(np.abs(state) ** 2, nqubits * (2,)).astype(rtype)
    if unmeasured:
        probs = np.sum(probs, axis=unmeasured)
    # remaining axes are the measured qubits in ascending order; permute to the caller's order
    rank = {q: i for i, q in enumerate(sorted(qubits))}
    probs = np.transpose(probs, [rank[q] for q in qubits])
    return probs.ravel()


# --------------------------------------------------------------------------------------
# S1 / S2 / S3: sampling
# --------------------------------------------------------------------------------------


def choice_from_uniforms(probs, uniforms):
    """What ``np.random.choice(len(p), size, p=p)`` computes (abstract.py:1265-1309 -> numpy legacy
    ``RandomState.choice``): sequential float64 cumsum, normalise by the last entry, right-bisect."""
    cdf = np.cumsum(np.asarray(probs, dtype=np.float64))
    cdf /= cdf[-1]
    return np.searchsorted(cdf, uniforms, side="right").astype(np.int64)


def sample_shots(probs, nshots, seed=None):
    """abstract.py:2774-2781. ``seed`` -> ``np.random.seed`` (Backend.set_seed, :181-187)."""
    if seed is not None:
        np.random.seed(seed)
    return choice_from_uniforms(probs, np.random.random_sample(nshots))


def sample_frequencies(probs, nshots, seed=None):
    """abstract.py:2760-2772 + update_frequencies :2795-2801 (renormalise first, 2^18-shot batches)."""
    if seed is not None:
        np.random.seed(seed)
    probs = np.asarray(probs)
    nprobs = probs / np.sum(probs)
    freqs = np.zeros(len(nprobs), dtype=np.int64)
    batches = (nshots // SHOT_BATCH_SIZE) * [SHOT_BATCH_SIZE] + [nshots % SHOT_BATCH_SIZE]
    for b in batches:
        samples = choice_from_uniforms(nprobs, np.random.random_sample(b))
        res, counts = np.unique(samples, return_counts=True)
        freqs[res] += counts
    return Counter({i: int(f) for i, f in enumerate(freqs) if f > 0})


def samples_to_binary(samples, nqubits):
    """abstract.py:2783-2786 (MSB first)."""
    qrange = np.arange(nqubits - 1, -1, -1, dtype=np.int32)
    return np.right_shift(np.asarray(samples)[:, None], qrange) % 2


def samples_to_decimal(samples, nqubits):
    """abstract.py:2788-2793."""
    qrange = np.arange(nqubits - 1, -1, -1, dtype=np.int32)
    qrange = (2**qrange)[:, None]
    return (np.asarray(samples, dtype=np.int32) @ qrange)[:, 0]


def calculate_frequencies(samples):
    """abstract.py:2727-2732."""
    res, counts = np.unique(samples, return_counts=True)
    return Counter(dict(zip(res.tolist(), counts.tolist())))


# --------------------------------------------------------------------------------------
# C1: collapse
# --------------------------------------------------------------------------------------


def collapse_statevector(state, qubits, shot, nqubits, normalize=True):
    """abstract.py:3279-3304 + _append_zeros :3236-3247. ``qubits`` sorted, ``shot`` decimal (MSB first)."""
    qubits = list(qubits)
    shot = int(np.asarray(shot).ravel()[0])
    binshot = [(shot >> (len(qubits) - 1 - i)) & 1 for i in range(len(qubits))]
    shape = state.shape
    psi = np.reshape(state, nqubits * (2,))
    order = qubits + [q for q in range(nqubits) if q not in qubits]
    psi = np.transpose(psi, order)
    psi = np.reshape(psi, (2 ** len(qubits),) + (nqubits - len(qubits)) * (2,))[shot]
    if normalize:
        norm = np.sqrt(np.sum(np.abs(psi) ** 2))
        psi = psi / norm
    for q, r in zip(qubits, binshot):
        psi = np.expand_dims(psi, q)
        zeros = np.zeros_like(psi)
        psi = np.concatenate([zeros, psi], axis=q) if r == 1 else np.concatenate([psi, zeros], axis=q)
    return np.reshape(psi, shape)


# --------------------------------------------------------------------------------------
# Circuit descriptions used by the BASELINE configs.  An op is (name, qubits, params) with
# qubits == gate.qubits (sorted controls + targets).
# --------------------------------------------------------------------------------------


def qft_ops(nqubits, with_swaps=True):
    """models/qft.py:47-58."""
    ops = []
    for i1 in range(nqubits):
        ops.append(("H", (i1,), ()))
        for i2 in range(i1 + 1, nqubits):
            theta = math.pi / 2 ** (i2 - i1)
            ops.append(("CU1", (i2, i1), (theta,)))
    if with_swaps:
        for q in range(nqubits // 2):
            ops.append(("SWAP", (q, nqubits - q - 1), ()))
    return ops


def variational_ops(nqubits, nlayers, thetas):
    """examples/benchmarks/circuits.py:7-22 (== tests/test_models_circuit_fuse.py:109-115)."""
    theta = iter(thetas)
    ops = []
    for _ in range(nlayers):
        for i in range(nqubits):
            ops.append(("RY", (i,), (float(next(theta)),)))
        for i in range(0, nqubits - 1, 2):
            ops.append(("CZ", (i, i + 1), ()))
        for i in range(nqubits):
            ops.append(("RY", (i,), (float(next(theta)),)))
        for i in range(1, nqubits - 2, 2):
            ops.append(("CZ", (i, i + 1), ()))
        ops.append(("CZ", (0, nqubits - 1), ()))
    return ops


def random_ops(nqubits, ngates, seed):
    """tests/test_models_circuit_fuse.py:124-138 generator (same legacy-RNG consumption order)."""
    np.random.seed(seed)
    one = ["RX", "RY", "RZ"]
    two = ["CNOT", "CZ", "SWAP"]
    thetas = np.pi * np.random.random((ngates,))
    ops = []
    for i in range(ngates):
        g = one[int(np.random.randint(0, 3))]
        q0 = int(np.random.randint(0, nqubits))
        ops.append((g, (q0,), (float(thetas[i]),)))
        g = two[int(np.random.randint(0, 3))]
        q0, q1 = np.random.randint(0, nqubits, (2,))
        while q0 == q1:
            q0, q1 = np.random.randint(0, nqubits, (2,))
        ops.append((g, (int(q0), int(q1)), ()))
    return ops


def run_ops(state, ops, nqubits, dtype=None):
    """The gate loop of _execute_circuit (abstract.py:3321-3322) over an op list."""
    dtype = dtype or state.dtype
    for name, qubits, params in ops:
        state = apply_gate(state, gate_matrix(name, *params, dtype=dtype), qubits, nqubits)
    return state


def pauli_expectation(state, paulis, qubits, nqubits):
    """<state| P |state> for a Pauli string: the contraction backends/abstract.py:3030-3053 performs for one term of
    ``exp_value_observable_symbolic`` ("abc,ad,cf,dbf->": conj(state), the 2x2 matrices, state), restated as
    einsum over the reshaped state."""
    mats = {"I": np.eye(2), "X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1.0, -1.0])}
    psi = np.asarray(state, dtype=np.complex128).reshape(nqubits * (2,))
    phi = psi
    for f, q in zip(paulis, qubits):
        if f == "I":
            continue
        phi = np.moveaxis(np.tensordot(mats[f].astype(np.complex128), phi, axes=([1], [q])), 0, q)
    return complex(np.vdot(psi, phi))
