"""CPU checks of the product's planner + canonicaliser + shared-memory passes through tests/emul
(the same C++ the CUDA kernel runs, compiled for the host), against the oracle."""

import os

import numpy as np
import pytest

from conftest import tol
from helpers import ops_from_named, oracle_run, rand_state, random_zoo
from oracle import numpy_oracle as orc
from qibo_b200.ops import Op
import emul


@pytest.fixture(params=[None, "2", "3"], ids=["L=default", "L=2", "L=3"])
def low_bits(request, monkeypatch):
    if request.param is not None:
        monkeypatch.setenv("QB_SWEEP_LOW_BITS", request.param)
    return request.param


def test_canonical_forms():
    CK = dict(NOOP=0, DENSE=1, DIAG=2, PHASE=3, SWAP=4)
    g = orc.gate_matrix
    assert emul.canon_kind(4, Op(g("CU1", 0.3), (2, 0))) == (CK["PHASE"], 0, 2)
    assert emul.canon_kind(4, Op(g("CZ"), (1, 3))) == (CK["PHASE"], 0, 2)
    assert emul.canon_kind(4, Op(g("Z"), (1,))) == (CK["PHASE"], 0, 1)
    assert emul.canon_kind(4, Op(g("CNOT"), (1, 3))) == (CK["DENSE"], 1, 1)
    assert emul.canon_kind(4, Op(g("TOFFOLI"), (1, 3, 0))) == (CK["DENSE"], 1, 2)
    assert emul.canon_kind(4, Op(g("CCZ"), (1, 3, 0))) == (CK["PHASE"], 0, 3)
    assert emul.canon_kind(4, Op(g("SWAP"), (1, 3))) == (CK["SWAP"], 2, 0)
    assert emul.canon_kind(4, Op(g("SWAP"), (1, 3), (0,))) == (CK["SWAP"], 2, 1)
    assert emul.canon_kind(4, Op(g("RZ", 0.2), (1,))) == (CK["DIAG"], 1, 0)
    assert emul.canon_kind(4, Op(g("RZZ", 0.2), (1, 2))) == (CK["DIAG"], 2, 0)
    assert emul.canon_kind(4, Op(g("H"), (1,))) == (CK["DENSE"], 1, 0)
    assert emul.canon_kind(4, Op(g("fSim", 0.1, 0.2), (1, 0))) == (CK["DENSE"], 2, 0)
    assert emul.canon_kind(4, Op(g("CRX", 0.1), (1, 0))) == (CK["DENSE"], 1, 1)
    assert emul.canon_kind(4, Op(np.eye(4), (1, 0))) == (CK["NOOP"], 0, 0)
    with pytest.raises(ValueError):
        emul.canon_kind(4, Op(g("CZ"), (1, 1)))
    with pytest.raises(ValueError):
        emul.canon_kind(4, Op(g("H"), (4,)))


def test_golden_gates_through_emulator(golden, low_bits):
    for i, c in enumerate(golden.cases("gate_cases")):
        psi, ref, mat = golden[f"gate{i}_in"], golden[f"gate{i}_out"], golden[f"gate{i}_matrix"]
        if c["is_controlled_by"]:
            op = Op(mat, tuple(c["targets"]), tuple(c["controls"]))
        else:
            op = Op(mat, tuple(c["qubits"]))
        for fuse in (True, False):
            out, _ = emul.apply_program(psi, c["nqubits"], [op], fuse=fuse)
            assert np.abs(out - ref).max() < tol(c["dtype"]), c["tag"]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [1, 2, 3, 7, 12, 13, 14, 15])
def test_qft(n, dtype, low_bits):
    psi = rand_state(n, 100 + n, dtype)
    ops = ops_from_named(orc.qft_ops(n))
    ref = orc.run_ops(psi, orc.qft_ops(n), n, dtype=dtype)
    out, stats = emul.apply_program(psi, n, ops)
    assert np.abs(out - ref).max() < tol(dtype)
    assert stats.nops == len(ops) and stats.nsweeps >= 1
    if n >= 8:
        assert stats.nsweeps < len(ops) / 4  # several gates per sweep
    out1, stats1 = emul.apply_program(psi, n, ops, fuse=False)
    assert np.abs(out1 - ref).max() < tol(dtype)
    assert stats1.nsweeps == len(ops)
    # analytic identity: QFT == ifft with ortho norm
    if dtype == "complex128":
        assert np.abs(out - np.fft.ifft(psi, norm="ortho")).max() < 1e-12


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_variational_and_random(dtype, low_bits):
    n = 14
    thetas = 2 * np.pi * np.random.default_rng(7).random(2 * 2 * n)
    named = orc.variational_ops(n, 2, thetas)
    psi = rand_state(n, 5, dtype)
    out, stats = emul.apply_program(psi, n, ops_from_named(named))
    assert np.abs(out - orc.run_ops(psi, named, n, dtype=dtype)).max() < tol(dtype)
    named = orc.random_ops(n, 60, seed=11)
    out, stats = emul.apply_program(psi, n, ops_from_named(named))
    assert np.abs(out - orc.run_ops(psi, named, n, dtype=dtype)).max() < tol(dtype)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_zoo(dtype, seed, low_bits):
    n = 14 if dtype == "complex128" else 15
    ops = random_zoo(n, 40, seed)
    psi = rand_state(n, seed, dtype)
    ref = oracle_run(psi, ops, n)
    out, stats = emul.apply_program(psi, n, ops)
    assert np.abs(out - ref).max() < tol(dtype)
    out, stats = emul.apply_program(psi, n, ops, fuse=False)
    assert np.abs(out - ref).max() < tol(dtype)


def test_fused_reference_queue(golden):
    """circuit.fuse() output of the reference (dense FusedGate matrices) through the sweep path."""
    for i, c in enumerate(golden.cases("circ_cases")):
        if c["queue"] is None:
            continue
        psi = golden[f"circ{i}_in"]
        ops = [Op(golden[f"circ{i}_q{j}"], tuple(q)) for j, q in enumerate(c["queue"])]
        out, _ = emul.apply_program(psi, c["nqubits"], ops)
        assert np.abs(out - golden[f"circ{i}_out"]).max() < tol(c["dtype"]), c["tag"]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [1, 3, 6, 9, 13, 16])
def test_permute_qubits(n, dtype):
    """K8 index math: dst[.. qubit dest[q] ..] = src[.. qubit q ..] equals the transpose NumPy does."""
    rng = np.random.default_rng(n)
    psi = rand_state(n, n, dtype)
    perms = [list(range(n)), list(range(n - 1, -1, -1))] + [rng.permutation(n).tolist() for _ in range(4)]
    for dest in perms:
        out = emul.permute_qubits(psi, n, dest)
        # axis q of the source becomes axis dest[q] of the destination
        ref = np.moveaxis(psi.reshape(n * (2,)), list(range(n)), dest).reshape(-1)
        np.testing.assert_array_equal(out, ref)
    # a run of SWAP gates is such a permutation
    named = [("SWAP", (q, n - 1 - q), ()) for q in range(n // 2)]
    if named:
        dest = list(range(n))
        for _, (a, b), _ in named:
            dest = [b if d == a else a if d == b else d for d in dest]
        np.testing.assert_array_equal(emul.permute_qubits(psi, n, dest), orc.run_ops(psi, named, n, dtype=dtype))


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_six_qubit_dense_block(dtype):
    """A 64x64 block does not fit the shared-memory program buffer: its matrix stays in the global blob."""
    from helpers import rand_unitary

    n = 14
    rng = np.random.default_rng(6)
    ops = [Op(rand_unitary(6, rng), (13, 2, 7, 0, 9, 4)), Op(orc.gate_matrix("H"), (3,)), Op(rand_unitary(6, rng), (1, 5, 3, 8, 12, 6), (10,))]
    psi = rand_state(n, 3, dtype)
    ref = oracle_run(psi, ops, n)
    out, _ = emul.apply_program(psi, n, ops)
    assert np.abs(out - ref).max() < tol(dtype)


def _phase_heavy_program(rng, n, ngates):
    """H / RY next to phases on their own qubits, CU1 / CZ / multi-controlled phases sharing qubits: the shapes the
    planner's fan merging, re-rooting and stage / layer fusion rewrite."""
    ops = []
    for _ in range(ngates):
        kind = rng.integers(0, 7)
        if kind == 0:
            ops.append(Op(orc.gate_matrix("H"), (int(rng.integers(0, n)),)))
        elif kind == 1:
            a, b = rng.permutation(n)[:2].tolist()
            ops.append(Op(orc.gate_matrix("CU1", float(rng.uniform(0, 6))), (a, b)))
        elif kind == 2:
            ops.append(Op(np.array([1, np.exp(1j * rng.uniform(0, 6))]), (int(rng.integers(0, n)),), is_diagonal=True))
        elif kind == 3:
            a, b = rng.permutation(n)[:2].tolist()
            ops.append(Op(orc.gate_matrix("CZ"), (a, b)))
        elif kind == 4:
            ops.append(Op(np.exp(1j * rng.uniform(0, 6, size=2)), (int(rng.integers(0, n)),), is_diagonal=True))
        elif kind == 5:
            qs = rng.permutation(n)[: int(rng.integers(2, 4))].tolist()
            ops.append(Op(orc.gate_matrix("U1", float(rng.uniform(0, 6))), (qs[0],), tuple(qs[1:])))
        else:
            ops.append(Op(orc.gate_matrix("RY", float(rng.uniform(0, 6))), (int(rng.integers(0, n)),)))
    return ops


@pytest.mark.parametrize("seed", range(6))
def test_phase_heavy_programs(seed):
    rng = np.random.default_rng(seed)
    for _ in range(12):
        n = int(rng.integers(4, 15))
        ops = _phase_heavy_program(rng, n, int(rng.integers(5, 60)))
        for dtype in ("complex128", "complex64"):
            psi = rand_state(n, seed, dtype)
            ref = oracle_run(psi.astype(np.complex128), ops, n)
            out, _ = emul.apply_program(psi, n, ops)
            assert np.abs(out - ref).max() < (1e-12 if dtype == "complex128" else 2e-5), (seed, n, dtype)


def test_planner_regressions():
    """Found by the fuzz above.  (1) A phase on exactly a fan's control set is a scalar on the CONTROL slice: a one-entry
    fan with such a scalar is no longer a lone phase on all of its bits.  (2) A pass with a fused stage op AND a fused
    layer op: payloads are re-assigned by owner (positions counted from the end of the payload list were off by the
    number of fused pairs -> wrong tables, out-of-range offsets).  (3) complex64 sweeps whose program outgrows the
    shared-memory buffer (one 1 KiB group table per distinct register set) are planned again with fewer ops."""
    g = orc.gate_matrix
    ph = lambda t: np.array([1, np.exp(1j * t)])  # noqa: E731
    cases = [
        (11, [Op(g("CZ"), (6, 5)), Op(ph(0.24), (6,), is_diagonal=True)]),
        (6, [Op(g("RY", 3.0), (2,)), Op(g("RY", 1.9), (3,)), Op(ph(3.3), (2,), is_diagonal=True), Op(g("H"), (2,)),
             Op(g("U1", -0.6), (2,), (3,)), Op(g("RY", 3.6), (4,)), Op(ph(1.2), (4,), is_diagonal=True)]),
        (5, [Op(g("H"), (1,)), Op(ph(0.7), (1,), is_diagonal=True), Op(g("CU1", 0.4), (1, 3)), Op(ph(1.1), (3,), is_diagonal=True),
             Op(ph(2.1), (1,), is_diagonal=True), Op(g("CZ"), (3, 1)), Op(g("H"), (3,)), Op(ph(0.3), (3,), is_diagonal=True)]),
    ]
    for n, ops in cases:
        for dtype in ("complex128", "complex64"):
            psi = rand_state(n, 3, dtype)
            out, _ = emul.apply_program(psi, n, ops)
            assert np.abs(out - oracle_run(psi.astype(np.complex128), ops, n)).max() < (1e-12 if dtype == "complex128" else 2e-5)
    for seed in (21, 33, 79, 84):  # "internal: sweep program too large" before the retry
        rng = np.random.default_rng(1000 + seed)
        n = int(rng.integers(5, 15))
        ops = random_zoo(n, int(rng.integers(10, 70)), 5000 + seed, max_dense=min(5, n))
        psi = rand_state(n, seed, "complex64")
        out, _ = emul.apply_program(psi, n, ops)
        assert np.abs(out - oracle_run(psi.astype(np.complex128), ops, n)).max() < 1e-4


# ------------------------------------------------------------------------------------------ parameter slots (f2)
def test_gate_families_match_reference_matrices():
    """qb_program_set_params evaluates RX/RY/RZ/U1/CU1/CRX/CRY/CRZ from angles (csrc/qb_families.hpp): same matrices as
    the oracle's restatement of backends/npmatrices.py."""
    from qibo_b200 import _lib

    fam = {"RX": _lib.QB_GATE_RX, "RY": _lib.QB_GATE_RY, "RZ": _lib.QB_GATE_RZ, "U1": _lib.QB_GATE_U1,
           "CU1": _lib.QB_GATE_CU1, "CRX": _lib.QB_GATE_CRX, "CRY": _lib.QB_GATE_CRY, "CRZ": _lib.QB_GATE_CRZ}
    for name, code in fam.items():
        for theta in (0.0, 0.3, -1.7, np.pi, 5.9):
            nt = 2 if name.startswith("C") else 1
            got = emul.family_matrix(code, [theta], nt)
            want = orc.gate_matrix(name, theta)
            assert np.abs(got - want).max() < 1e-15, (name, theta)
    assert np.abs(emul.family_matrix(fam["CU1"], [0.4], 2, is_diagonal=True) - np.diag(orc.gate_matrix("CU1", 0.4))).max() < 1e-15
    with pytest.raises(ValueError):
        emul.family_matrix(fam["RX"], [0.4], 1, is_diagonal=True)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("case", ["variational", "qft_angles", "random"])
def test_schedule_replay_with_new_parameters(case, dtype):
    """Circuit.set_parameters + re-execution (models/circuit.py:788-857): the ops of a planned program get new matrices,
    the schedule is reused, only the passes are emitted again -- result equal to the oracle on the NEW ops.  A parameter
    that changes the gate structure (angle 0 -> identity) falls back to a fresh plan."""
    n = 14
    rng = np.random.default_rng(3)

    def build(thetas):
        th = iter(thetas)
        if case == "variational":
            return ops_from_named(orc.variational_ops(n, 3, list(thetas)))
        if case == "qft_angles":
            named = []
            for i in range(n):
                named.append(("H", (i,), ()))
                for j in range(i + 1, n):
                    named.append(("CU1", (j, i), (float(next(th)),)))
            return ops_from_named(named)
        named = []
        for k in range(60):
            q = k % n
            named.append((["RX", "RY", "RZ"][k % 3], (q,), (float(next(th)),)))
            named.append((["CNOT", "CZ"][k % 2], (q, (q + 1 + k % 5) % n), ()))
        return ops_from_named(named)

    nparams = {"variational": 2 * 3 * n, "qft_angles": n * (n - 1) // 2, "random": 60}[case]
    old = build(rng.uniform(0.1, 6.0, nparams))
    new_thetas = rng.uniform(0.1, 6.0, nparams)
    new = build(new_thetas)
    psi = rand_state(n, 5, dtype)
    out, replayed = emul.apply_program_replay(psi, n, old, new)
    assert replayed
    assert np.abs(out - oracle_run(psi, new, n)).max() < tol(dtype)
    # structure change: one angle becomes 0 (RY(0) = identity, CU1(0) = identity)
    new_thetas[3] = 0.0
    new0 = build(new_thetas)
    out, replayed = emul.apply_program_replay(psi, n, old, new0)
    assert np.abs(out - oracle_run(psi, new0, n)).max() < tol(dtype)


def _permuted_reference(state, n, dest):
    return np.moveaxis(state.reshape(n * (2,)), list(range(n)), list(dest)).reshape(-1)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [13, 14, 16, 17])
def test_qft_bit_reversal_rides_on_last_sweep(n, dtype):
    """Permuting sweep (qb_apply_program_permuted): the QFT's closing SWAP run written by the last sweep, out of place,
    through the destination layout -- same state as gates + SWAPs in the oracle; also when the program is emitted again
    on the kept schedule (qb_program_set_params)."""
    from qibo_b200.engine import split_segments

    ops = ops_from_named(orc.qft_ops(n))
    segs = split_segments(ops, n, True, True)
    assert [k for k, _ in segs] == ["opsperm"]
    gops, dest = segs[0][1]
    psi = rand_state(n, 300 + n, dtype)
    ref = orc.run_ops(psi, orc.qft_ops(n), n, dtype=dtype)
    tile_bits = 12 if dtype == "complex128" else 13
    for replay in (False, True):
        out, stats, fused = emul.apply_program_permuted(psi, n, gops, dest, replay=replay)
        assert fused == (n > tile_bits)
        assert not np.isnan(out).any()  # every destination amplitude written
        assert np.abs(out - ref).max() < tol(dtype)
        if fused:
            assert stats.perm_fused == 1 and stats.nstage_sweeps == stats.nsweeps or dtype == "complex64"


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("seed", range(6))
def test_random_programs_with_trailing_permutation(dtype, seed):
    """Zoo programs (dense blocks, controls, fans, swaps) followed by random permutations: fused when the geometry allows,
    otherwise the plain plan + permutation -- the result is the same either way; the pure permutation (no gate) too."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(14, 17))
    psi = rand_state(n, seed, dtype)
    perms = [list(range(n - 1, -1, -1)), rng.permutation(n).tolist()]
    # partial reversals (what a rank of a sharded QFT ends with: the leading qubits stay)
    keep = int(rng.integers(1, 4))
    perms.append(list(range(keep)) + list(range(n - 1, keep - 1, -1)))
    nfused = 0
    for dest in perms:
        for ngates in (0, 25):
            ops = random_zoo(n, ngates, seed) if ngates else []
            ref = _permuted_reference(oracle_run(psi, ops, n) if ops else psi, n, dest)
            out, stats, fused = emul.apply_program_permuted(psi, n, ops, dest)
            nfused += int(fused)
            assert not np.isnan(out).any()
            assert np.abs(out - ref).max() < tol(dtype), (n, dest, ngates, fused)
    assert nfused >= 2


def test_sharded_qft_tail_fuses_into_one_sweep():
    """The last local segment of every rank of a sharded QFT -- three stages on the leading local qubits, then the reversal
    of the others -- is ONE permuting sweep (tile = 5 + 4 row bits + the three stage bits)."""
    from qibo_b200.engine import split_segments

    n = 16
    named = []
    for q in range(3):
        named.append(("H", (q,), ()))
        for j in range(q + 1, 3):
            named.append(("CU1", (j, q), (np.pi / 2 ** (j - q),)))
    named += [("SWAP", (3 + q, n - 1 - q), ()) for q in range((n - 3) // 2)]
    ops = ops_from_named(named)
    segs = split_segments(ops, n, True, True)
    assert [k for k, _ in segs] == ["opsperm"]
    gops, dest = segs[0][1]
    psi = rand_state(n, 5, "complex128")
    out, stats, fused = emul.apply_program_permuted(psi, n, gops, dest)
    assert fused and stats.nsweeps == 1
    assert np.abs(out - orc.run_ops(psi, named, n)).max() < 1e-12


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("seed", range(8))
def test_merged_sign_gates(dtype, seed, low_bits, monkeypatch):
    """MH_SIGNS (opt-in, QB_SIGNS=1): runs of +-1 diagonal gates (CZ, Z, CZ fans sharing a control, CCZ left alone) between
    layers of one-qubit gates become one micro-op per pass; register x register, register x thread, register x
    outside-the-tile and thread x thread pairs, lone Z terms on every kind of bit, a global -1."""
    monkeypatch.setenv("QB_SIGNS", "1")
    rng = np.random.default_rng(4000 + seed)
    n = int(rng.integers(14, 18))
    named = []
    for layer in range(int(rng.integers(2, 5))):
        for q in range(n):
            r = rng.random()
            if r < 0.6:
                named.append(("RY", (q,), (float(rng.uniform(0, 6)),)))
            elif r < 0.75:
                named.append(("H", (q,), ()))
        style = int(rng.integers(0, 3))
        if style == 0:  # the ansatz's ladder
            for q in range(0, n - 1, 2):
                named.append(("CZ", (q, q + 1), ()))
            for q in range(1, n - 1, 2):
                named.append(("CZ", (q, q + 1), ()))
            named.append(("CZ", (0, n - 1), ()))
        elif style == 1:  # random pairs, some twice (they cancel), Z gates, a global sign through Z X Z X
            for _ in range(int(rng.integers(5, 40))):
                a, b = rng.permutation(n)[:2].tolist()
                named.append(("CZ", (a, b), ()))
                if rng.random() < 0.2:
                    named.append(("Z", (int(rng.integers(0, n)),), ()))
        else:  # fans: one control, many partners; three-qubit phases stay separate ops
            c = int(rng.integers(0, n))
            for q in rng.permutation(n)[: int(rng.integers(2, n))].tolist():
                if q != c:
                    named.append(("CZ", (c, q), ()))
            a, b, c3 = rng.permutation(n)[:3].tolist()
            named.append(("CCZ", (a, b, c3), ()))
            named.append(("CU1", (a, b), (float(rng.uniform(0.1, 3)),)))
    psi = rand_state(n, seed, dtype)
    ref = orc.run_ops(psi, named, n, dtype=dtype)
    out, stats = emul.apply_program(psi, n, ops_from_named(named))
    assert np.abs(out - ref).max() < tol(dtype)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_out_of_place_program_identity_permutation(dtype):
    """Engine.compile_copying: a program whose last sweep writes into ANOTHER buffer (the identity permutation riding on a
    permuting sweep) -- what the chunk of a pipelined exchange that stays on its rank runs.  QFT stages without the
    closing swaps (the chunk programs of a sharded QFT) and a zoo program; the source buffer may be left in an
    intermediate state, the destination must be complete."""
    for n, named in ((15, [g for g in orc.qft_ops(15) if g[0] != "SWAP"]), (16, None)):
        ops = ops_from_named(named) if named is not None else random_zoo(n, 30, 3)
        psi = rand_state(n, 40 + n, dtype)
        ref = oracle_run(psi, ops, n)
        out, stats, fused = emul.apply_program_permuted(psi, n, ops, list(range(n)))
        assert fused and stats.perm_fused == 1
        assert not np.isnan(out).any()
        assert np.abs(out - ref).max() < tol(dtype)
