"""Generates tests/golden/*.npz by running the UNMODIFIED reference (qibo 0.3.5 NumpyBackend).

Run in the build container (the reference cannot travel to the GPU box):

    ./baseline/install_ref.sh          # or: PYTHONPATH=/root/reference/src:<shim>
    PYTHONPATH=baseline/_ref python tests/golden/make_golden.py

Every fixture stores the inputs and the reference's outputs, so the tests never need qibo.
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

import qibo  # noqa: E402
from qibo import Circuit, gates  # noqa: E402
from qibo.backends import NumpyBackend  # noqa: E402
from qibo.models import QFT  # noqa: E402


def rand_state(n, rng, dtype="complex128"):
    x = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    x /= np.linalg.norm(x)
    return x.astype(dtype)


def rand_unitary(k, rng):
    a = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def main():
    qibo.set_backend("numpy")
    rng = np.random.default_rng(20261017)
    out = {}

    # ---------------- G1/G2: single gates on random states -------------------------------
    gate_cases = []
    n = 6

    def add_case(tag, make_gate, nq=n, dtype="complex128"):
        be = NumpyBackend()
        be.set_dtype(dtype)
        psi = rand_state(nq, rng, dtype)
        gate = make_gate()
        res = be.apply_gate(gate, np.copy(psi), nq)
        idx = len(gate_cases)
        matrix = np.asarray(gate.matrix(be))
        gate_cases.append(
            dict(
                tag=tag,
                nqubits=nq,
                dtype=dtype,
                name=gate.__class__.__name__,
                qubits=[int(q) for q in gate.qubits],
                controls=[int(q) for q in gate.control_qubits],
                targets=[int(q) for q in gate.target_qubits],
                is_controlled_by=bool(gate.is_controlled_by),
                params=[float(p) for p in getattr(gate, "parameters", ()) if np.isscalar(p)],
            )
        )
        out[f"gate{idx}_in"] = psi
        out[f"gate{idx}_out"] = np.asarray(res)
        out[f"gate{idx}_matrix"] = matrix

    for dtype in ("complex128", "complex64"):
        for q in (0, 2, 5):
            add_case(f"H({q})", lambda q=q: gates.H(q), dtype=dtype)
            add_case(f"X({q})", lambda q=q: gates.X(q), dtype=dtype)
            add_case(f"Y({q})", lambda q=q: gates.Y(q), dtype=dtype)
            add_case(f"Z({q})", lambda q=q: gates.Z(q), dtype=dtype)
            add_case(f"RX({q})", lambda q=q: gates.RX(q, 0.1234 + q), dtype=dtype)
            add_case(f"RY({q})", lambda q=q: gates.RY(q, 0.4321 + q), dtype=dtype)
            add_case(f"RZ({q})", lambda q=q: gates.RZ(q, 1.234 + q), dtype=dtype)
            add_case(f"U1({q})", lambda q=q: gates.U1(q, 0.77 + q), dtype=dtype)
            add_case(f"U3({q})", lambda q=q: gates.U3(q, 0.1, 0.2 + q, 0.3), dtype=dtype)
            add_case(f"S({q})", lambda q=q: gates.S(q), dtype=dtype)
            add_case(f"T({q})", lambda q=q: gates.T(q), dtype=dtype)
        for a, b in ((0, 1), (1, 0), (0, 5), (5, 0), (2, 4), (4, 3)):
            add_case(f"CNOT({a},{b})", lambda a=a, b=b: gates.CNOT(a, b), dtype=dtype)
            add_case(f"CZ({a},{b})", lambda a=a, b=b: gates.CZ(a, b), dtype=dtype)
            add_case(f"CU1({a},{b})", lambda a=a, b=b: gates.CU1(a, b, 0.3 + a), dtype=dtype)
            add_case(f"CRX({a},{b})", lambda a=a, b=b: gates.CRX(a, b, 0.3 + a), dtype=dtype)
            add_case(f"SWAP({a},{b})", lambda a=a, b=b: gates.SWAP(a, b), dtype=dtype)
            add_case(f"iSWAP({a},{b})", lambda a=a, b=b: gates.iSWAP(a, b), dtype=dtype)
            add_case(f"fSim({a},{b})", lambda a=a, b=b: gates.fSim(a, b, 0.5, 0.25), dtype=dtype)
            add_case(f"RZZ({a},{b})", lambda a=a, b=b: gates.RZZ(a, b, 0.65), dtype=dtype)
            add_case(f"RXX({a},{b})", lambda a=a, b=b: gates.RXX(a, b, 0.65), dtype=dtype)
        for a, b, c in ((0, 1, 2), (5, 3, 0), (2, 0, 4)):
            add_case(f"TOFFOLI({a},{b},{c})", lambda a=a, b=b, c=c: gates.TOFFOLI(a, b, c), dtype=dtype)
            add_case(f"CCZ({a},{b},{c})", lambda a=a, b=b, c=c: gates.CCZ(a, b, c), dtype=dtype)
        # controlled_by (G2 path; test_gates_gates.py:1738-1940)
        add_case("X.cby(0,1,2)->5", lambda: gates.X(5).controlled_by(0, 1, 2), dtype=dtype)
        add_case("RY.cby(4,1)->0", lambda: gates.RY(0, 0.9).controlled_by(4, 1), dtype=dtype)
        add_case("SWAP.cby(0)", lambda: gates.SWAP(2, 4).controlled_by(0), dtype=dtype)
        add_case("SWAP.cby(5,1)", lambda: gates.SWAP(3, 0).controlled_by(5, 1), dtype=dtype)
        add_case("fSim.cby(3)", lambda: gates.fSim(5, 1, 0.4, 0.6).controlled_by(3), dtype=dtype)
        add_case("Z.cby(1,2,3,4)", lambda: gates.Z(0).controlled_by(1, 2, 3, 4), dtype=dtype)
        # dense k-qubit unitaries (test_gates_gates.py:1658-1735), user-ordered targets
        for k, targets in ((1, (3,)), (2, (4, 1)), (3, (5, 0, 2)), (4, (1, 3, 0, 5)), (5, (4, 0, 5, 2, 1))):
            u = rand_unitary(k, rng)
            add_case(f"Unitary{k}{targets}", lambda u=u, t=targets: gates.Unitary(u, *t), dtype=dtype)
        u = rand_unitary(2, rng)
        add_case("Unitary2.cby(0,3)", lambda u=u: gates.Unitary(u, 4, 1).controlled_by(0, 3), dtype=dtype)

    out["gate_cases"] = np.array(json.dumps(gate_cases))

    # ---------------- circuits: QFT / variational / random / fused --------------------------
    circ_cases = []

    def add_circuit(tag, circuit, nq, dtype="complex128", zero=False, store_queue=False):
        be = NumpyBackend()
        be.set_dtype(dtype)
        psi = None if zero else rand_state(nq, rng, dtype)
        res = be.execute_circuit(circuit, None if zero else np.copy(psi)).state()
        idx = len(circ_cases)
        queue = None
        if store_queue:  # the reference fuser's output: (gate.qubits, gate.matrix) per queue entry
            queue = []
            for j, g in enumerate(circuit.queue):
                queue.append([int(q) for q in g.qubits])
                out[f"circ{idx}_q{j}"] = np.asarray(g.matrix(be))
        circ_cases.append(dict(tag=tag, nqubits=nq, dtype=dtype, zero=zero, queue=queue))
        if not zero:
            out[f"circ{idx}_in"] = psi
        out[f"circ{idx}_out"] = np.asarray(res)

    for nq in (3, 5, 8, 11):
        for dtype in ("complex128", "complex64"):
            add_circuit(f"qft{nq}", QFT(nq), nq, dtype)
            add_circuit(f"qft{nq}_zero", QFT(nq), nq, dtype, zero=True)
    add_circuit("qft7_noswap", QFT(7, with_swaps=False), 7)

    from oracle import numpy_oracle as orc  # op-list generators are part of what is pinned

    var_thetas = 2 * np.pi * np.random.default_rng(7).random(2 * 3 * 10)
    out["var_thetas"] = var_thetas
    c = Circuit(10)
    for name, qubits, params in orc.variational_ops(10, 3, var_thetas):
        c.add(getattr(gates, name)(*qubits, *params))
    for dtype in ("complex128", "complex64"):
        add_circuit("var10x3", c, 10, dtype)
        for mq in (2, 3, 4, 5):
            add_circuit(f"var10x3_fuse{mq}", c.fuse(max_qubits=mq), 10, dtype, store_queue=True)

    c = Circuit(9)
    for name, qubits, params in orc.random_ops(9, 40, seed=11):
        c.add(getattr(gates, name)(*qubits, *params))
    for dtype in ("complex128", "complex64"):
        add_circuit("rand9x40", c, 9, dtype)
        for mq in (2, 3, 4):
            add_circuit(f"rand9x40_fuse{mq}", c.fuse(max_qubits=mq), 9, dtype, store_queue=True)
    out["circ_cases"] = np.array(json.dumps(circ_cases))

    # matrix_fused (G4): members + fused matrix for a few FusedGates of the random circuit
    fused = c.fuse(max_qubits=4)
    be = NumpyBackend()
    fcases = []
    for g in fused.queue:
        if g.__class__.__name__ != "FusedGate" or len(fcases) >= 6:
            continue
        i = len(fcases)
        members = []
        for j, m in enumerate(g.gates):
            out[f"fused{i}_m{j}"] = np.asarray(m.matrix(be))
            members.append(dict(qubits=[int(q) for q in m.qubits], ncontrols=len(m.control_qubits)))
        out[f"fused{i}_matrix"] = np.asarray(g.matrix(be))
        fcases.append(dict(targets=[int(q) for q in g.target_qubits], members=members))
    out["fused_cases"] = np.array(json.dumps(fcases))

    # ---------------- P1: probabilities / marginals with caller-ordered qubits ----------------
    pcases = []
    for dtype in ("complex128", "complex64"):
        psi = rand_state(7, rng, dtype)
        be = NumpyBackend()
        be.set_dtype(dtype)
        for qubits in ([0], [6], [0, 5, 3], list(range(7)), [1, 5, 2, 0], [6, 0], [3, 2, 1], [6, 5, 4, 3, 2, 1, 0]):
            i = len(pcases)
            out[f"prob{i}_in"] = psi
            out[f"prob{i}_out"] = np.asarray(be.calculate_probabilities(psi, qubits, 7))
            pcases.append(dict(nqubits=7, qubits=qubits, dtype=dtype))
    out["prob_cases"] = np.array(json.dumps(pcases))

    # ---------------- S1/S2: sampling goldens --------------------------------------------
    scases = []
    be = NumpyBackend()
    for nb, nshots, seed in ((4, 1000, 1234), (32, 5000, 1234), (1024, 20000, 7), (2**16, 30000, 99)):
        p = rng.random(nb)
        p /= p.sum()
        i = len(scases)
        be.set_seed(seed)
        out[f"samp{i}_probs"] = p
        out[f"samp{i}_shots"] = np.asarray(be.sample_shots(p, nshots))
        be.set_seed(seed)
        freq = be.sample_frequencies(p * 0.999, nshots)  # un-normalised input: S2 renormalises
        keys = np.array(sorted(freq), dtype=np.int64)
        out[f"samp{i}_freq_keys"] = keys
        out[f"samp{i}_freq_vals"] = np.array([freq[k] for k in keys], dtype=np.int64)
        scases.append(dict(nbins=nb, nshots=nshots, seed=seed))
    out["samp_cases"] = np.array(json.dumps(scases))

    # golden of tests/test_measurements_probabilistic.py:11-34 -- H(0),H(1) + M(0,1), seed 1234, 1000 shots
    be = NumpyBackend()
    be.set_seed(1234)
    c = Circuit(2)
    c.add(gates.H(0))
    c.add(gates.H(1))
    c.add(gates.M(0, 1))
    freq = be.execute_circuit(c, nshots=1000).frequencies(False)
    out["probabilistic_golden"] = np.array([freq[k] for k in range(4)], dtype=np.int64)
    assert dict(freq) == {0: 249, 1: 231, 2: 253, 3: 267}, freq

    # C1 README config: QFT(15) + M(all), 100 shots, seed 1234
    be = NumpyBackend()
    be.set_seed(1234)
    c = QFT(15)
    c.add(gates.M(*range(15)))
    res = be.execute_circuit(c, nshots=100)
    out["c1_samples"] = np.asarray(res.samples(binary=False))
    st = np.asarray(res.state())
    out["c1_state_head"] = st[:64]
    out["c1_state_absmax"] = np.array(np.abs(st).max())

    # ---------------- C1: collapse -------------------------------------------------------
    ccases = []
    be = NumpyBackend()
    for qubits, shot in (([0], 1), ([2, 4], 2), ([0, 3, 5], 5), ([5], 0), ([1, 2, 3, 4], 9)):
        for normalize in (True, False):
            psi = rand_state(6, rng)
            i = len(ccases)
            out[f"coll{i}_in"] = psi
            out[f"coll{i}_out"] = np.asarray(
                be.collapse_state(np.copy(psi), qubits, np.array([shot]), 6, normalize=normalize)
            )
            ccases.append(dict(nqubits=6, qubits=qubits, shot=shot, normalize=normalize))
    out["coll_cases"] = np.array(json.dumps(ccases))

    # ---------------- X1: density-matrix gate application -------------------------------------
    dcases = []
    be = NumpyBackend()
    for make in (
        lambda: gates.H(1),
        lambda: gates.CNOT(0, 2),
        lambda: gates.RY(2, 0.3),
        lambda: gates.CU1(2, 0, 0.7),
        lambda: gates.fSim(1, 3, 0.3, 0.2),
        lambda: gates.X(3).controlled_by(0, 1),
    ):
        g = make()
        psi = rand_state(4, rng)
        phi = rand_state(4, rng)
        rho = 0.6 * np.outer(psi, psi.conj()) + 0.4 * np.outer(phi, phi.conj())
        i = len(dcases)
        out[f"dm{i}_in"] = rho
        out[f"dm{i}_out"] = np.asarray(be.apply_gate(g, np.copy(rho), 4))
        out[f"dm{i}_matrix"] = np.asarray(g.matrix(be))
        dcases.append(
            dict(
                nqubits=4,
                qubits=[int(q) for q in g.qubits],
                controls=[int(q) for q in g.control_qubits],
                targets=[int(q) for q in g.target_qubits],
                is_controlled_by=bool(g.is_controlled_by),
            )
        )
    out["dm_cases"] = np.array(json.dumps(dcases))

    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
