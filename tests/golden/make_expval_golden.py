"""Generates tests/golden/expval_golden.npz by importing the UNMODIFIED reference (run in the build container only):

    PYTHONPATH=baseline/_ref python tests/golden/make_expval_golden.py

Pins oracle.numpy_oracle.pauli_expectation (and, through it, the K9 kernels) to the reference's
Backend.exp_value_observable_symbolic (backends/abstract.py:2946-3054) and overlap_statevector (:2180-2190) on the
NumpyBackend.
"""

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    from qibo import Circuit, gates
    from qibo.backends import NumpyBackend

    be = NumpyBackend()
    rng = np.random.default_rng(77)
    out, cases = {}, []
    for n in (3, 6, 9):
        c = Circuit(n)
        for q in range(n):
            c.add(gates.RY(q, theta=float(rng.random() * 3)))
            c.add(gates.RX(q, theta=float(rng.random() * 3)))
        for q in range(n - 1):
            c.add(gates.CNOT(q, q + 1))
        for q in range(n):
            c.add(gates.RZ(q, theta=float(rng.random() * 3)))
        res = be.execute_circuit(c)
        state = np.asarray(res.state())
        terms, term_qubits, coeffs = [], [], []
        for _ in range(12):
            k = int(rng.integers(1, min(n, 4) + 1))
            qs = [int(q) for q in rng.choice(n, size=k, replace=False)]
            term = "".join(rng.choice(list("XYZ"), size=k))
            terms.append(term)
            term_qubits.append(tuple(qs))
            coeffs.append(float(rng.normal()))
        per_term = [float(be.exp_value_observable_symbolic(c, [t], [q], [1.0], n)) for t, q in zip(terms, term_qubits)]
        total = float(be.exp_value_observable_symbolic(c, terms, term_qubits, coeffs, n))
        other = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
        other /= np.linalg.norm(other)
        i = len(cases)
        out[f"ev{i}_state"] = state
        out[f"ev{i}_other"] = other
        out[f"ev{i}_per_term"] = np.array(per_term)
        out[f"ev{i}_overlap"] = np.array(complex(be.overlap_statevector(state, other)))
        cases.append(dict(nqubits=n, terms=terms, term_qubits=[list(q) for q in term_qubits], coefficients=coeffs, total=total))
    out["cases"] = np.array(json.dumps(cases))
    path = os.path.join(HERE, "expval_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
