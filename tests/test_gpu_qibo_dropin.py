"""Drop-in tests: the backend plugged into the UNMODIFIED reference package through Qibo's own loader
(`qibo.set_backend("qibo_b200")`), compared with the reference NumpyBackend on the same circuits.
They mirror the reference's own tests (tests/test_models_qft.py, test_models_circuit_fuse.py,
test_measurements*.py, test_models_circuit_execution.py).  Skipped when qibo is not importable."""

import numpy as np
import pytest

from conftest import have_qibo, tol

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_qibo(), reason="reference package not importable (baseline/_ref)")]


@pytest.fixture()
def backends():
    import qibo
    from qibo.backends import NumpyBackend, construct_backend

    ours = construct_backend("qibo_b200")
    ref = NumpyBackend()
    yield ours, ref
    qibo.backends._Global._backend = None


def rand_state(n, seed, dtype="complex128"):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    return (x / np.linalg.norm(x)).astype(dtype)


def test_loader_and_attributes(backends):
    import qibo
    from qibo.backends import list_available_backends

    ours, _ = backends
    assert ours.name == "qibo_b200" and ours.platform == "cuda-sm100a" and ours.device == "/GPU:0"
    assert ours.supports_multigpu and ours.qubits is None and ours.connectivity is None and ours.natives is None
    assert list_available_backends("qibo-b200")["qibo-b200"] == {"cuda-sm100a": True}
    qibo.set_backend("qibo-b200")
    assert qibo.get_backend().name == "qibo_b200"
    qibo.set_backend("qibo_b200", platform="cuda-sm100a")
    with pytest.raises(ValueError):
        ours.set_device("/CPU:0")


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [3, 8, 15])
def test_qft_matches_numpy_backend(backends, n, dtype):
    from qibo.models import QFT

    ours, ref = backends
    ours.set_dtype(dtype)
    ref.set_dtype(dtype)
    psi = rand_state(n, n, dtype)
    a = ours.execute_circuit(QFT(n), np.copy(psi)).state()
    b = ref.execute_circuit(QFT(n), np.copy(psi)).state()
    assert a.dtype == b.dtype == np.dtype(dtype)
    assert np.abs(ours.to_numpy(a) - b).max() < tol(dtype)
    # zero initial state, and Circuit.__call__ through the global backend
    import qibo

    qibo.set_backend("qibo_b200")
    qibo.set_dtype(dtype)
    c = QFT(n)
    res = c()
    assert np.abs(np.asarray(res.state()) - ref.execute_circuit(QFT(n)).state()).max() < tol(dtype)
    qibo.set_dtype("complex128")


def test_config1_qft15_with_shots(backends, golden):
    """BASELINE config 1 (README example): QFT(15) + M(all), nshots=100, seed 1234: samples bit-exact."""
    from qibo import gates
    from qibo.models import QFT

    ours, ref = backends
    c = QFT(15)
    c.add(gates.M(*range(15)))
    ours.set_seed(1234)
    res = ours.execute_circuit(c, nshots=100)
    np.testing.assert_array_equal(np.asarray(res.samples(binary=False)), golden["c1_samples"])
    assert np.abs(np.asarray(res.state())[:64] - golden["c1_state_head"]).max() < 1e-12
    ours.set_seed(1234)
    c2 = QFT(15)
    c2.add(gates.M(*range(15)))
    freq = ours.execute_circuit(c2, nshots=100).frequencies(binary=False)
    ref.set_seed(1234)
    c3 = QFT(15)
    c3.add(gates.M(*range(15)))
    assert freq == ref.execute_circuit(c3, nshots=100).frequencies(binary=False)


def test_probabilistic_measurement_golden(backends):
    """tests/test_measurements_probabilistic.py:11-34 (NumPy golden frequencies for seed 1234)."""
    from qibo import Circuit, gates

    ours, _ = backends
    ours.set_seed(1234)
    c = Circuit(2)
    c.add(gates.H(0))
    c.add(gates.H(1))
    c.add(gates.M(0, 1))
    result = ours.execute_circuit(c, nshots=1000)
    assert dict(result.frequencies(False)) == {0: 249, 1: 231, 2: 253, 3: 267}


@pytest.mark.parametrize("max_qubits", [2, 3, 4, 5])
@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_variational_layer_fusion(backends, max_qubits, dtype):
    """tests/test_models_circuit_fuse.py:101-118 on our backend, and against NumpyBackend."""
    from qibo import Circuit, gates

    ours, ref = backends
    ours.set_dtype(dtype)
    ref.set_dtype(dtype)
    n, nlayers = 11, 2
    theta = iter(2 * np.pi * np.random.default_rng(3).random(2 * nlayers * n))
    c = Circuit(n)
    for _ in range(nlayers):
        c.add(gates.RY(i, next(theta)) for i in range(n))
        c.add(gates.CZ(i, i + 1) for i in range(0, n - 1, 2))
        c.add(gates.RY(i, next(theta)) for i in range(n))
        c.add(gates.CZ(i, i + 1) for i in range(1, n - 1, 2))
        c.add(gates.CZ(0, n - 1))
    fused = c.fuse(max_qubits=max_qubits)
    a = ours.to_numpy(ours.execute_circuit(fused).state())
    b = ours.to_numpy(ours.execute_circuit(c).state())
    r = ref.execute_circuit(c).state()
    assert np.abs(a - r).max() < tol(dtype) and np.abs(b - r).max() < tol(dtype)
    ours.assert_circuitclose(fused, c, atol=10 * tol(dtype))


def test_gate_zoo_apply_gate(backends):
    """tests/test_gates_gates.py style: backend.apply_gate on host arrays, incl. controlled_by and Unitary."""
    from qibo import gates

    ours, ref = backends
    n = 6
    rng = np.random.default_rng(0)
    u3 = np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))[0]
    zoo = [
        gates.H(0), gates.X(5), gates.Y(2), gates.Z(3), gates.S(1), gates.T(4), gates.SX(0), gates.RX(1, 0.3), gates.RY(2, 0.4),
        gates.RZ(3, 0.5), gates.U1(0, 0.1), gates.U2(1, 0.1, 0.2), gates.U3(2, 0.1, 0.2, 0.3), gates.GPI2(4, 0.3),
        gates.CNOT(0, 5), gates.CY(5, 1), gates.CZ(2, 3), gates.CSX(1, 2), gates.CRX(0, 1, 0.2), gates.CRY(4, 2, 0.3), gates.CRZ(1, 5, 0.4),
        gates.CU1(2, 0, 0.5), gates.CU2(3, 1, 0.1, 0.2), gates.CU3(0, 4, 0.1, 0.2, 0.3), gates.SWAP(0, 5), gates.iSWAP(1, 2),
        gates.FSWAP(3, 4), gates.fSim(0, 3, 0.2, 0.4), gates.RXX(1, 4, 0.3), gates.RYY(2, 5, 0.3), gates.RZZ(0, 2, 0.3), gates.RZX(3, 5, 0.3),
        gates.GIVENS(1, 3, 0.4), gates.RBS(0, 4, 0.4), gates.ECR(2, 4), gates.TOFFOLI(0, 1, 2), gates.CCZ(3, 5, 1), gates.DEUTSCH(0, 2, 4, 0.3),
        gates.Unitary(u3, 4, 0, 2), gates.X(5).controlled_by(0, 1, 2), gates.RY(0, 0.9).controlled_by(4, 1),
        gates.SWAP(2, 4).controlled_by(0), gates.fSim(5, 1, 0.4, 0.6).controlled_by(3), gates.Unitary(u3, 1, 3, 5).controlled_by(0, 2),
        gates.GeneralizedfSim(1, 2, np.linalg.qr(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))[0], 0.3),
    ]
    for g in zoo:
        psi = rand_state(n, 7)
        a = ours.apply_gate(g, np.copy(psi), n)
        b = ref.apply_gate(g, np.copy(psi), n)
        assert len(a) == 2**n and a.dtype == b.dtype
        assert np.abs(ours.to_numpy(a) - b).max() < 1e-12, g.name


def test_density_matrix_circuit(backends):
    from qibo import Circuit, gates

    ours, ref = backends
    c = Circuit(4, density_matrix=True)
    c.add([gates.H(0), gates.CNOT(0, 1), gates.RY(2, 0.3), gates.CU1(3, 1, 0.5), gates.X(3).controlled_by(0, 2), gates.fSim(1, 2, 0.3, 0.1)])
    a = ours.execute_circuit(c).state()
    b = ref.execute_circuit(c).state()
    assert a.shape == b.shape == (16, 16)
    assert np.abs(ours.to_numpy(a) - b).max() < 1e-12


def test_measurement_ordering_registers_and_collapse(backends):
    """tests/test_measurements.py:137-156 (qubit order M(1,5,2,0)) and a collapse circuit."""
    from qibo import Circuit, gates

    ours, ref = backends
    c = Circuit(6)
    c.add(gates.X(0))
    c.add(gates.X(1))
    c.add(gates.M(1, 5, 2, 0))
    res = ours.execute_circuit(c, nshots=100)
    assert res.frequencies() == {"1001": 100}
    np.testing.assert_array_equal(res.samples(binary=False), 100 * [9])
    probs_a = np.asarray(res.probabilities([1, 5, 2, 0]))
    c2 = Circuit(6)
    c2.add([gates.X(0), gates.X(1)])
    c2.add(gates.M(1, 5, 2, 0))
    np.testing.assert_allclose(probs_a, ref.execute_circuit(c2, nshots=100).probabilities([1, 5, 2, 0]), atol=1e-14)

    # mid-circuit collapse: identical RNG stream -> identical outcomes and final frequencies
    def collapse_circuit():
        c = Circuit(3)
        c.add(gates.H(0))
        c.add(gates.H(1))
        out = c.add(gates.M(0, collapse=True))
        c.add(gates.CNOT(0, 2))
        c.add(gates.M(1, 2))
        return c

    ours.set_seed(123)
    fa = ours.execute_circuit(collapse_circuit(), nshots=40).frequencies()
    ref.set_seed(123)
    fb = ref.execute_circuit(collapse_circuit(), nshots=40).frequencies()
    assert fa == fb


def test_initial_state_not_clobbered_and_errors(backends):
    from qibo import Circuit, gates
    from qibo.models import QFT

    ours, _ = backends
    psi = rand_state(5, 1)
    keep = psi.copy()
    dev = ours.cast(psi)  # host array
    res = ours.execute_circuit(QFT(5), psi).state()
    np.testing.assert_array_equal(psi, keep)
    dstate = ours._to_device(psi)
    ours.execute_circuit(QFT(5), dstate)
    np.testing.assert_array_equal(dstate.numpy(), keep)  # device inputs are cloned too
    with pytest.raises(ValueError):
        ours.execute_circuit(QFT(5), np.ones(7))
    assert res.tolist() is not None and len(res) == 32 and res.shape == (32,)
    # tests/test_models_circuit_execution.py:54-60
    c = Circuit(40)
    c.add(gates.H(0))
    with pytest.raises(RuntimeError):
        ours.execute_circuit(c)


def test_symbolic_hamiltonian_expectation_on_device(backends):
    """f1: `SymbolicHamiltonian.expectation(circuit)` -> Backend.exp_value_observable_symbolic
    (abstract.py:2946-3054) runs term by term on the device state (K9) and matches the NumpyBackend; the overlap of
    two device states goes through K9 as well."""
    from qibo import Circuit, gates
    from qibo.hamiltonians import SymbolicHamiltonian
    from qibo.symbols import X, Y, Z

    ours, ref = backends
    n = 6

    def circuit():
        c = Circuit(n)
        for q in range(n):
            c.add(gates.RY(q, theta=0.2 + 0.37 * q))
            c.add(gates.RX(q, theta=1.1 - 0.21 * q))
        for q in range(n - 1):
            c.add(gates.CNOT(q, q + 1))
        return c

    def form():
        return 0.7 * X(0) * Z(3) - 1.3 * Y(1) * Y(2) + 0.25 * Z(5) + 2.0 * X(2) * Y(4) * Z(0) + 0.4

    a = SymbolicHamiltonian(form(), nqubits=n, backend=ours).expectation(circuit())
    b = SymbolicHamiltonian(form(), nqubits=n, backend=ref).expectation(circuit())
    assert abs(float(a) - float(b)) < 1e-12
    s1 = ours.execute_circuit(circuit()).state()
    s2 = ours.execute_circuit(Circuit(n)).state()
    r1 = ref.execute_circuit(circuit()).state()
    r2 = ref.execute_circuit(Circuit(n)).state()
    assert abs(complex(ours.overlap_statevector(s1, s2)) - complex(ref.overlap_statevector(r1, r2))) < 1e-12


def rand_density_matrix(n, seed, dtype="complex128"):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(2**n, 2**n)) + 1j * rng.normal(size=(2**n, 2**n))
    rho = a @ a.conj().T
    return (rho / np.trace(rho)).astype(dtype)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_density_matrix_probabilities_collapse_on_device(backends, dtype):
    """X1 without a host detour: diag(rho) marginals in the caller's qubit order (abstract.py:2741-2749), collapse of rows
    and columns with trace renormalisation (:3249-3277), `controlled_by` gates (:3199-3234), minus_state -- device kernels
    against the NumpyBackend on the same matrices."""
    from qibo import gates
    from qibo_b200.array import DeviceArray

    ours, ref = backends
    ours.set_dtype(dtype)
    ref.set_dtype(dtype)
    n = 5
    rho = rand_density_matrix(n, 3, dtype)
    t = 1e-12 if dtype == "complex128" else 1e-5
    for qubits in ([0], [4], [1, 3], [3, 1], [4, 0, 2], list(range(n)), [2, 4, 1, 0, 3]):
        a = ours.calculate_probabilities(np.copy(rho), qubits, n, density_matrix=True)
        b = ref.calculate_probabilities(np.copy(rho), qubits, n, density_matrix=True)
        assert isinstance(a, DeviceArray) and a.dtype == b.dtype and a.shape == b.shape
        assert np.abs(ours.to_numpy(a) - b).max() < t, qubits
    for qubits, shot in (([0], 1), ([1, 3], 2), ([0, 2, 4], 5), (list(range(n)), 19)):
        for normalize in (True, False):
            a = ours.collapse_state(np.copy(rho), qubits, np.array([shot]), n, normalize=normalize, density_matrix=True)
            b = ref.collapse_state(np.copy(rho), qubits, np.array([shot]), n, normalize=normalize, density_matrix=True)
            assert isinstance(a, DeviceArray) and a.shape == b.shape
            assert np.abs(ours.to_numpy(a) - b).max() < t, (qubits, shot, normalize)
    for gate in (gates.X(2).controlled_by(0, 4), gates.RY(0, 0.7).controlled_by(3), gates.SWAP(1, 2).controlled_by(0),
                 gates.Unitary(np.linalg.qr(np.random.default_rng(1).normal(size=(4, 4)))[0], 3, 1).controlled_by(4)):
        a = ours.apply_gate(gate, np.copy(rho), n)
        b = ref.apply_gate(gate, np.copy(rho), n)
        assert np.abs(ours.to_numpy(a) - b).max() < t, gate.name
    for fn in ("zero_state", "plus_state", "minus_state"):
        a = getattr(ours, fn)(3, density_matrix=True)
        b = getattr(ref, fn)(3, density_matrix=True)
        assert isinstance(a, DeviceArray) and a.shape == b.shape
        assert np.abs(ours.to_numpy(a) - b).max() < t, fn
    assert np.abs(ours.to_numpy(ours.minus_state(4)) - ref.minus_state(4)).max() < t


def test_density_matrix_collapse_circuit(backends):
    """A density-matrix circuit with a collapsing measurement, end to end through Circuit() (tests/test_measurements_collapse.py)."""
    from qibo import Circuit, gates

    ours, ref = backends
    outs = []
    for be in (ours, ref):
        c = Circuit(4, density_matrix=True)
        c.add(gates.H(q) for q in range(4))
        c.add(gates.CNOT(0, 2))
        m = c.add(gates.M(1, 2, collapse=True))
        c.add(gates.RX(3, 0.4))
        c.add(gates.M(0, 3))
        be.set_seed(7)
        res = be.execute_circuit(c, nshots=50)
        outs.append((be.to_numpy(res.state()), m.samples()[0], res.frequencies()))
    assert np.abs(outs[0][0] - outs[1][0]).max() < 1e-12
    assert list(outs[0][1]) == list(outs[1][1]) and outs[0][2] == outs[1][2]


def test_measurement_registers_in_add_order_sharded_path(backends):
    """The joint measurement keeps the qubits in the order they were added (result.py:444-458): M(3, 1) then M(0)."""
    from qibo import Circuit, gates

    ours, ref = backends
    outs = []
    for be in (ours, ref):
        c = Circuit(5)
        c.add(gates.RY(q, theta=0.3 + 0.2 * q) for q in range(5))
        c.add(gates.CNOT(0, 3))
        c.add(gates.M(3, 1, register_name="a"))
        c.add(gates.M(0, register_name="b"))
        be.set_seed(11)
        res = be.execute_circuit(c, nshots=300)
        outs.append((res.frequencies(), res.frequencies(registers=True)))
    assert outs[0] == outs[1]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_set_parameters_reexecution_uses_parameter_slots(backends, dtype):
    """Circuit.set_parameters + re-execution (models/circuit.py:788-857; the loop of models/variational.py:45-120): the
    backend keeps ONE compiled program per Circuit object and patches the angles in place (qb_program_set_params).  Every
    execution must equal the NumpyBackend's on the same angles -- plain, fused, density-matrix, gates outside the
    angle families (U3, controlled_by), and an angle of 0 that changes the gate structure."""
    from qibo import Circuit, gates

    ours, ref = backends
    ours.set_dtype(dtype)
    ref.set_dtype(dtype)
    n = 10
    rng = np.random.default_rng(5)

    def ansatz(density_matrix=False):
        c = Circuit(n, density_matrix=density_matrix)
        for layer in range(2):
            c.add(gates.RY(q, theta=0.1) for q in range(n))
            c.add(gates.CZ(q, q + 1) for q in range(0, n - 1, 2))
            c.add(gates.RX(q, theta=0.1) for q in range(n))
            c.add(gates.CU1(q, q + 1, theta=0.1) for q in range(1, n - 1, 2))
            c.add(gates.RZ(q, theta=0.1) for q in range(0, n, 3))
            c.add(gates.CRY(0, n - 1, theta=0.1))
        c.add(gates.U3(2, 0.1, 0.2, 0.3))
        c.add(gates.RY(4, theta=0.1).controlled_by(1, 7))
        c.add(gates.M(0, 3))
        return c

    for density_matrix, nq in ((False, n), (True, 5)):
        c = ansatz(density_matrix) if not density_matrix else None
        if density_matrix:
            c = Circuit(nq, density_matrix=True)
            c.add(gates.RY(q, theta=0.1) for q in range(nq))
            c.add(gates.CRZ(0, 1, theta=0.2))
            c.add(gates.U3(2, 0.1, 0.2, 0.3))
            c.add(gates.CZ(1, 2))
            c.add(gates.RX(3, theta=0.3))
        nparams = len(c.get_parameters("flatlist"))
        program_ids = set()
        for step in range(4):
            theta = rng.uniform(0.05, 6.2, nparams)
            if step == 3:
                theta[0] = 0.0  # RY(0): the gate becomes an identity -> the library plans afresh
            c.set_parameters(theta)
            a = ours.execute_circuit(c).state()
            cached = getattr(c, "_qb200_program", None)
            assert cached is not None
            program_ids.add(id(cached["program"]))
            c2 = c.copy(deep=True)
            b = ref.execute_circuit(c2).state()
            assert np.abs(ours.to_numpy(a) - b).max() < tol(dtype), (density_matrix, step)
        assert len(program_ids) == 1  # compiled once, patched afterwards
    # through circuit.fuse(): the members of the fused blocks keep their slots
    c = ansatz()
    fc = c.fuse(max_qubits=2)
    for step in range(3):
        # (one entry per gate: the shallow copy behind fuse() loses the flat-list bookkeeping, models/circuit.py:410-411)
        theta = [tuple(rng.uniform(0.05, 6.2, g.nparams)) if g.nparams > 1 else float(rng.uniform(0.05, 6.2)) for g in c.trainable_gates]
        fc.set_parameters(theta)
        a = ours.execute_circuit(fc).state()
        c.set_parameters(theta)
        b = ref.execute_circuit(c.copy(deep=True)).state()
        assert np.abs(ours.to_numpy(a) - b).max() < tol(dtype), step


def test_dense_hamiltonian_expectation_on_device(backends):
    """Backend.expectation_value (abstract.py:2807-2825) through Hamiltonian.expectation on a device-resident final state
    and on a density matrix, against the NumpyBackend."""
    from qibo import Circuit, gates, hamiltonians

    ours, ref = backends
    n = 6
    vals = []
    for be in (ours, ref):
        h = hamiltonians.XXZ(n, delta=0.7, dense=True, backend=be)
        c = Circuit(n)
        c.add(gates.RY(q, theta=0.4 + 0.3 * q) for q in range(n))
        c.add(gates.CNOT(q, q + 1) for q in range(n - 1))
        st = be.execute_circuit(c).state()
        cd = Circuit(n, density_matrix=True)
        cd.add(gates.RX(q, theta=0.2 + 0.1 * q) for q in range(n))
        cd.add(gates.CZ(0, 3))
        rho = be.execute_circuit(cd).state()
        vals.append((float(h.expectation_from_state(st)), float(h.expectation_from_state(rho)), float(h.expectation_from_state(st, normalize=True))))
    assert np.abs(np.array(vals[0]) - np.array(vals[1])).max() < 1e-12


def test_state_host_io_streaming_and_dump_load(backends, tmp_path, monkeypatch):
    """f3: `state(numpy=True)`, `QuantumState.dump` / `load` (result.py:69-88, 116-163) on device-resident states; the
    chunked pinned device->host path (qibo_b200/array.py) gives the same bytes as a plain copy."""
    from qibo import Circuit, gates
    from qibo.result import QuantumState, load_result

    import qibo_b200.array as qarr

    ours, ref = backends
    n = 18
    c = Circuit(n)
    c.add(gates.RY(q, theta=0.1 + 0.05 * q) for q in range(n))
    c.add(gates.CNOT(q, q + 1) for q in range(n - 1))
    res = ours.execute_circuit(c)
    plain = res.state(numpy=True)
    monkeypatch.setattr(qarr, "STREAM_D2H_MIN_BYTES", 1 << 16)
    monkeypatch.setattr(qarr, "STREAM_D2H_CHUNK_BYTES", (1 << 18) + 4096)  # ragged last chunk
    qarr._staging.clear()
    streamed = res.state().numpy()
    np.testing.assert_array_equal(streamed, plain)
    assert np.abs(plain - ref.execute_circuit(c.copy(deep=True)).state()).max() < 1e-12
    path = tmp_path / "state.npy"
    res.dump(str(path))
    loaded = load_result(str(path))
    assert isinstance(loaded, QuantumState)
    np.testing.assert_array_equal(np.asarray(ours.to_numpy(loaded.state())), plain)


def test_pickling_backend_and_circuits_with_compiled_programs(backends):
    """qibo/parallel.py ships the backend and the circuits to joblib worker PROCESSES (tests/test_parallel.py:46-62), and
    MeasurementResult symbols are pickled with their backend (tests/test_measurements.py:470-480).  A circuit that has run
    carries its compiled program (device-resident, this process only): it must pickle to "no program" and be compiled again
    where it lands, and an Engine must come back as a fresh context, not as a copied library handle."""
    import pickle

    from qibo import gates
    from qibo.models import QFT

    ours, ref = backends
    c = QFT(9)
    c.add(gates.M(0, 3))
    first = ours.execute_circuit(c, nshots=10)
    assert getattr(c, "_qb200_program", None) is not None and c._qb200_program["program"] is not None
    c2 = pickle.loads(pickle.dumps(c))
    assert c2._qb200_program["program"] is None
    be2 = pickle.loads(pickle.dumps(ours))
    assert be2.engine_gpu is not ours.engine_gpu and be2.engine_gpu.handle.value != ours.engine_gpu.handle.value
    target = ref.execute_circuit(QFT(9)).state()
    for backend, circuit in ((ours, c2), (be2, c2), (be2, c)):
        out = backend.execute_circuit(circuit, nshots=10).state(numpy=True)
        assert np.abs(out - target).max() < 1e-12
    # the reference's own entry point: worker processes
    circuits = [QFT(n) for n in range(1, 9)]
    serial = [ours.execute_circuit(x) for x in circuits]
    parallel = ours.execute_circuits(circuits, processes=2)
    for a, b in zip(serial, parallel):
        assert np.abs(a.state(numpy=True) - b.state(numpy=True)).max() < 1e-12


@pytest.mark.parametrize("accelerators", [None, {"/GPU:0": 2}])
def test_trotter_state_evolution(backends, accelerators):
    """f4: Trotter evolution circuits (models/evolution.py:77-111, hamiltonians.circuit(dt)) run on the same sweep kernels --
    against the reference backend, step by step through a callback, with and without `accelerators` (one process: the
    logical devices collapse onto this GPU)."""
    from qibo import callbacks, hamiltonians, models

    ours, ref = backends
    n, dt = 5, 0.05
    finals = []
    for backend, acc in ((ours, accelerators), (ref, None)):
        ham = hamiltonians.TFIM(n, h=1.0, dense=False, backend=backend)
        energy = callbacks.Energy(hamiltonians.TFIM(n, h=1.0, backend=backend))
        evolution = models.StateEvolution(ham, dt, callbacks=[energy], accelerators=acc)
        psi0 = np.ones(2**n, dtype=np.complex128) / np.sqrt(2**n)
        final = evolution(final_time=0.5, initial_state=psi0.copy())
        finals.append((backend.to_numpy(final), [float(np.real(backend.to_numpy(x))) for x in energy[:]]))
    (psi_a, e_a), (psi_b, e_b) = finals
    assert np.abs(psi_a - psi_b).max() < 1e-12
    assert np.allclose(e_a, e_b, atol=1e-12)
