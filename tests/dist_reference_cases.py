"""The reference's own distributed-circuit tests, replayed on one process per GPU (TEST SCRIPT, run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist_reference_cases.py

Cases follow /root/reference/tests/test_callbacks.py:130-193 (entropy callbacks inside distributed circuits),
tests/test_measurements.py:108-137 and :173-190 (measurements on distributed circuits) and
tests/test_measurements_collapse.py (collapsing M), each against the NumpyBackend on the same circuit.  Rank 0 prints one
JSON line; every rank asserts."""

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, HERE]
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(os.path.join(_REF, "qibo")):
    sys.path.append(_REF)
os.environ.setdefault("QIBO_LOG_LEVEL", "3")


def main():
    import torch.distributed as dist

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from qibo import Circuit, gates
    from qibo.backends import NumpyBackend, construct_backend
    from qibo.callbacks import EntanglementEntropy, Norm, Overlap

    ours = construct_backend("qibo_b200")
    ours.set_device(f"/GPU:{local}")
    ref = NumpyBackend()
    acc = {f"/GPU:{i}": 1 for i in range(world)}
    done = []

    def close(a, b, atol=1e-12):
        np.testing.assert_allclose(np.asarray(ours.to_numpy(a)), np.asarray(b), atol=atol, rtol=0)

    # ---- test_callbacks.py:119-151 test_entropy_in_distributed_circuit
    target = ref.execute_circuit(_bell(Circuit, gates)).state()
    for conf, want in ((["H", "CNOT", "entropy"], [1.0]), (["H", "entropy", "CNOT"], [0.0]), (["entropy", "H", "CNOT"], [0.0]),
                       (["entropy", "H", "CNOT", "entropy"], [0.0, 1.0]), (["H", "entropy", "CNOT", "entropy"], [0.0, 1.0]),
                       (["entropy", "H", "entropy", "CNOT"], [0.0, 0.0])):
        entropy = EntanglementEntropy([0])
        c = Circuit(4, acc)
        for g in conf:
            c.add({"H": lambda: gates.H(0), "CNOT": lambda: gates.CNOT(0, 1), "entropy": lambda: gates.CallbackGate(entropy)}[g]())
        close(ours.execute_circuit(c).state(), target)
        np.testing.assert_allclose([float(ours.to_numpy(x)) for x in entropy[:]], want, atol=1e-7)
    done.append("entropy_in_distributed_circuit x6")

    # ---- test_callbacks.py:154-193 test_entropy_multiple_executions
    entropy = EntanglementEntropy([0])
    for theta in (0.1234, 0.4321):
        t = Circuit(4)
        t.add([gates.RY(0, theta), gates.CNOT(0, 1)])
        c = Circuit(4, acc)
        c.add(gates.RY(0, theta))
        c.add(gates.CallbackGate(entropy))
        c.add(gates.CNOT(0, 1))
        c.add(gates.CallbackGate(entropy))
        close(ours.execute_circuit(c).state(), ref.execute_circuit(t).state())

    def tent(t):
        cos, sin = np.cos(t / 2.0) ** 2, np.sin(t / 2.0) ** 2
        return -cos * np.log2(cos) - sin * np.log2(sin)

    np.testing.assert_allclose([float(ours.to_numpy(x)) for x in entropy[:]], [0, tent(0.1234), 0, tent(0.4321)], atol=1e-6)
    c = Circuit(8, acc)
    c.add(gates.CallbackGate(entropy))
    try:
        ours.execute_circuit(c)
        raise AssertionError("changing the callback's nqubits must raise")
    except RuntimeError:
        pass
    done.append("entropy_multiple_executions")

    # ---- Norm / Overlap callbacks as reductions over the ranks (no gather), 14 qubits
    n = 14
    rng = np.random.default_rng(3)
    phi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    phi /= np.linalg.norm(phi)
    results = []
    for be, kw in ((ours, acc), (ref, None)):
        norm, ov = Norm(), Overlap(phi)
        c = Circuit(n, kw) if kw else Circuit(n)
        c.add(gates.RY(q, theta=0.2 + 0.1 * q) for q in range(n))
        c.add(gates.CallbackGate(norm))
        c.add(gates.CNOT(q, q + 1) for q in range(n - 1))
        c.add(gates.CallbackGate(ov))
        c.add(gates.H(n - 1))
        c.add(gates.CallbackGate(ov))
        st = be.execute_circuit(c).state()
        results.append((be.to_numpy(st), [float(np.real(be.to_numpy(x))) for x in norm[:]], [complex(be.to_numpy(x)) for x in ov[:]]))
    close(results[0][0], results[1][0])
    np.testing.assert_allclose(results[0][1], results[1][1], atol=1e-12)
    np.testing.assert_allclose(results[0][2], results[1][2], atol=1e-12)
    done.append("norm_overlap_callbacks_sharded")

    # ---- test_measurements.py:108-116 test_measurement_circuit, :134-160 test_measurement_qubit_order
    c = Circuit(4, acc)
    c.add(gates.X(0))
    c.add(gates.M(0))
    res = ours.execute_circuit(c, nshots=100)
    assert res.frequencies(binary=False) == {1: 100} and res.frequencies(binary=True) == {"1": 100}
    np.testing.assert_array_equal(ours.to_numpy(res.samples(binary=True)), np.ones((100, 1)))
    for nshots in (100, 100000):
        c = Circuit(6, acc)
        c.add(gates.X(0))
        c.add(gates.X(1))
        c.add(gates.M(1, 5, 2, 0))
        res = ours.execute_circuit(c, nshots=nshots)
        assert res.frequencies(binary=True) == {"1001": nshots}
    done.append("measurement_circuit, measurement_qubit_order")

    # ---- the sharded measurement path keeps the order in which qubits were added (M(3, 1) then M(0)); same seed, same shots
    for gather_max in ("30", "0"):
        os.environ["QB_GATHER_MAX_QUBITS"] = gather_max
        outs = []
        for be, kw in ((ours, acc), (ref, None)):
            c = Circuit(8, kw) if kw else Circuit(8)
            c.add(gates.RY(q, theta=0.3 + 0.2 * q) for q in range(8))
            c.add(gates.CNOT(0, 3))
            c.add(gates.M(3, 1, register_name="a"))
            c.add(gates.M(0, register_name="b"))
            be.set_seed(11)
            res = be.execute_circuit(c, nshots=400)
            outs.append((dict(res.frequencies()), {k: dict(v) for k, v in res.frequencies(registers=True).items()}))
        if gather_max == "30":
            assert outs[0] == outs[1], outs
        else:  # (CDF built from per-rank partial marginals: a uniform within rounding of an edge may move one shot)
            diff = sum(abs(outs[0][0].get(k, 0) - outs[1][0].get(k, 0)) for k in set(outs[0][0]) | set(outs[1][0]))
            assert diff <= 4, outs
    os.environ.pop("QB_GATHER_MAX_QUBITS")
    done.append("registers_in_add_order (gathered and sharded)")

    # ---- collapsing measurement inside a distributed circuit (distcircuit.py:278-284): ONE distributed execution against
    # the same gates applied one by one on the NumpyBackend with the same seed (same outcome, same collapsed state) ...
    def collapse_circuit(kw):
        c = Circuit(6, kw) if kw else Circuit(6)
        c.add(gates.H(q) for q in range(6))
        c.add(gates.CNOT(0, 4))
        m = c.add(gates.M(0, 3, collapse=True))
        c.add(gates.RY(2, theta=0.4))
        c.add(gates.CNOT(5, 1))
        return c, m

    c, m = collapse_circuit(acc)
    ours.set_seed(123)
    got = ours.execute_distributed_circuit(c).state()
    t, mt = collapse_circuit(None)
    ref.set_seed(123)
    st = ref.zero_state(6)
    for gate in t.queue:
        st = gate.apply(ref, st, 6)
    close(got, st)
    assert [int(x) for x in m.samples()[0]] == [int(x) for x in mt.samples()[0]]  # (Circuit.add returns the M gate's result)
    # ... and through Circuit execution, which re-executes per shot (abstract.py:2532-2636, :2579-2582): same seed, same
    # frequencies of the final measurement
    outs = []
    for be, kw in ((ours, acc), (ref, None)):
        c, m = collapse_circuit(kw)
        c.add(gates.M(1, 5))
        be.set_seed(321)
        res = be.execute_circuit(c, nshots=12)
        outs.append(dict(res.frequencies()))
    assert outs[0] == outs[1], outs
    done.append("collapsing_measurement_in_distributed_circuit (one execution + per-shot re-execution)")

    # ---- a register too large to gather without measurements: a sharded state handle instead of an error
    os.environ["QB_GATHER_MAX_QUBITS"] = "0"
    c = Circuit(10, acc)
    c.add(gates.RY(q, theta=0.3 + 0.2 * q) for q in range(10))
    c.add(gates.CZ(q, q + 1) for q in range(9))
    handle = ours.execute_circuit(c)
    full = ref.execute_circuit(_same(c, Circuit, gates)).state()
    nl = 10 - int(np.log2(world))
    close(handle.tensor.cpu().numpy(), full[rank << nl : (rank + 1) << nl])
    close(handle.probabilities([9, 0, 4]), ref.calculate_probabilities(full, [9, 0, 4], 10))
    assert abs(handle.norm() - 1.0) < 1e-12 and handle.samples(50, qubits=[0, 1]).shape == (50, 2)
    os.environ.pop("QB_GATHER_MAX_QUBITS")
    done.append("sharded_state_handle")

    if rank == 0:
        print(json.dumps({"world": world, "passed": done}), flush=True)
    dist.destroy_process_group()


def _bell(Circuit, gates):
    c = Circuit(4)
    c.add([gates.H(0), gates.CNOT(0, 1)])
    return c


def _same(c, Circuit, gates):
    """The same gates on a plain (non-distributed) circuit for the NumpyBackend."""
    t = Circuit(c.nqubits)
    for g in c.queue:
        t.add(g.__class__(*g.init_args, **g.init_kwargs))
    return t


if __name__ == "__main__":
    main()
