"""Pins the oracle (oracle/numpy_oracle.py) to the reference's own outputs (tests/golden)."""

from collections import Counter

import numpy as np
import pytest

from conftest import tol
from oracle import numpy_oracle as orc


def test_gates_match_reference(golden):
    cases = golden.cases("gate_cases")
    assert len(cases) > 150
    for i, c in enumerate(cases):
        psi, ref, mat = golden[f"gate{i}_in"], golden[f"gate{i}_out"], golden[f"gate{i}_matrix"]
        n = c["nqubits"]
        if c["is_controlled_by"]:
            out = orc.apply_gate_controlled_by(psi, mat, c["controls"], c["targets"], n)
        else:
            out = orc.apply_gate(psi, mat, c["qubits"], n)
        assert out.dtype == ref.dtype, c["tag"]
        if c["is_controlled_by"]:  # reference uses einsum here (abstract.py:3190): last-bit rounding differs
            eps = np.finfo(ref.real.dtype).eps
            np.testing.assert_allclose(out, ref, atol=4 * eps, rtol=0, err_msg=c["tag"])
        else:  # same algorithm (transpose + matmul) -> bit-identical
            np.testing.assert_array_equal(out, ref, err_msg=c["tag"])
        # restated matrix table (G3) for the named gates
        if c["name"] not in ("Unitary",) and not c["is_controlled_by"] and c["name"] != "FusedGate":
            np.testing.assert_array_equal(orc.gate_matrix(c["name"], *c["params"], dtype=c["dtype"]), mat, err_msg=c["tag"])


def _ops_for(tag, golden):
    if tag.startswith("qft"):
        n = int(tag[3:].split("_")[0])
        return orc.qft_ops(n, with_swaps="noswap" not in tag)
    if tag.startswith("var10x3"):
        return orc.variational_ops(10, 3, golden["var_thetas"])
    if tag.startswith("rand9x40"):
        return orc.random_ops(9, 40, seed=11)
    raise KeyError(tag)


def test_circuits_match_reference(golden):
    for i, c in enumerate(golden.cases("circ_cases")):
        n, dtype = c["nqubits"], c["dtype"]
        psi = orc.zero_state(n, dtype) if c["zero"] else golden[f"circ{i}_in"]
        ref = golden[f"circ{i}_out"]
        if c["queue"] is None:
            out = orc.run_ops(psi, _ops_for(c["tag"], golden), n, dtype=dtype)
            np.testing.assert_array_equal(out, ref, err_msg=c["tag"])
        else:  # reference-fused queue: (qubits, dense matrix) entries through G1
            out = psi
            for j, qubits in enumerate(c["queue"]):
                out = orc.apply_gate(out, golden[f"circ{i}_q{j}"], qubits, n)
            np.testing.assert_array_equal(out, ref, err_msg=c["tag"])
            # and the fused circuit equals the unfused one within the north-star tolerance
            plain = orc.run_ops(psi, _ops_for(c["tag"], golden), n, dtype=dtype)
            assert np.abs(plain - ref).max() < 10 * tol(dtype)


def test_matrix_fused_matches_reference(golden):
    cases = golden.cases("fused_cases")
    assert cases
    for i, c in enumerate(cases):
        members = [
            (golden[f"fused{i}_m{j}"], m["qubits"], m["ncontrols"]) for j, m in enumerate(c["members"])
        ]
        out = orc.matrix_fused(members, c["targets"])
        np.testing.assert_allclose(out, golden[f"fused{i}_matrix"], atol=1e-15, rtol=0)


def test_probabilities_match_reference(golden):
    for i, c in enumerate(golden.cases("prob_cases")):
        out = orc.calculate_probabilities(golden[f"prob{i}_in"], c["qubits"], c["nqubits"])
        ref = golden[f"prob{i}_out"]
        assert out.dtype == ref.dtype
        np.testing.assert_array_equal(out, ref)


def test_sampling_matches_reference(golden):
    for i, c in enumerate(golden.cases("samp_cases")):
        p = golden[f"samp{i}_probs"]
        shots = orc.sample_shots(p, c["nshots"], seed=c["seed"])
        np.testing.assert_array_equal(shots, golden[f"samp{i}_shots"])
        freq = orc.sample_frequencies(p * 0.999, c["nshots"], seed=c["seed"])
        ref = Counter(dict(zip(golden[f"samp{i}_freq_keys"].tolist(), golden[f"samp{i}_freq_vals"].tolist())))
        assert freq == ref
    # tests/test_measurements_probabilistic.py:28-32 golden
    psi = orc.run_ops(orc.zero_state(2), [("H", (0,), ()), ("H", (1,), ())], 2)
    probs = orc.calculate_probabilities(psi, [0, 1], 2)
    freq = orc.sample_frequencies(probs, 1000, seed=1234)
    assert dict(freq) == {0: 249, 1: 231, 2: 253, 3: 267}
    assert golden["probabilistic_golden"].tolist() == [249, 231, 253, 267]


def test_config1_qft15_shots(golden):
    """BASELINE config 1: QFT(15) complex128, final state + nshots=100 measurement."""
    psi = orc.run_ops(orc.zero_state(15), orc.qft_ops(15), 15)
    np.testing.assert_array_equal(psi[:64], golden["c1_state_head"])
    probs = orc.calculate_probabilities(psi, list(range(15)), 15)
    shots = orc.sample_shots(probs, 100, seed=1234)
    np.testing.assert_array_equal(shots, golden["c1_samples"])


def test_collapse_matches_reference(golden):
    for i, c in enumerate(golden.cases("coll_cases")):
        out = orc.collapse_statevector(golden[f"coll{i}_in"], c["qubits"], [c["shot"]], c["nqubits"], c["normalize"])
        np.testing.assert_array_equal(out, golden[f"coll{i}_out"])


def test_density_matrix_matches_reference(golden):
    for i, c in enumerate(golden.cases("dm_cases")):
        rho, mat, n = golden[f"dm{i}_in"], golden[f"dm{i}_matrix"], c["nqubits"]
        if c["is_controlled_by"]:
            k = len(c["controls"]) + len(c["targets"])
            full = np.eye(2**k, dtype=mat.dtype)
            full[-mat.shape[0] :, -mat.shape[0] :] = mat
            mat, qubits = full, sorted(c["controls"]) + c["targets"]
        else:
            qubits = c["qubits"]
        out = orc.apply_gate_density_matrix(rho, mat, qubits, n)
        np.testing.assert_allclose(out, golden[f"dm{i}_out"], atol=1e-15, rtol=0)


def test_samples_binary_decimal_roundtrip():
    s = np.array([0, 5, 1023, 77], dtype=np.int64)
    b = orc.samples_to_binary(s, 10)
    assert b.shape == (4, 10) and b[1].tolist() == [0, 0, 0, 0, 0, 0, 0, 1, 0, 1]
    np.testing.assert_array_equal(orc.samples_to_decimal(b, 10), s)
    assert orc.calculate_frequencies(np.array([1, 1, 3])) == Counter({1: 2, 3: 1})


def test_pauli_expectation_matches_reference():
    """f1: the oracle's one-term contraction against Backend.exp_value_observable_symbolic / overlap_statevector of the
    reference NumpyBackend (tests/golden/make_expval_golden.py)."""
    import json
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "expval_golden.npz"))
    for i, c in enumerate(json.loads(str(z["cases"]))):
        n, state = c["nqubits"], z[f"ev{i}_state"]
        got = [orc.pauli_expectation(state, t, q, n) for t, q in zip(c["terms"], c["term_qubits"])]
        assert np.abs(np.imag(got)).max() < 1e-14
        assert np.abs(np.real(got) - z[f"ev{i}_per_term"]).max() < 1e-13
        total = sum(co * g.real for co, g in zip(c["coefficients"], got))
        assert abs(total - c["total"]) < 1e-12
        assert abs(np.vdot(state, z[f"ev{i}_other"]) - complex(z[f"ev{i}_overlap"])) < 1e-14
