"""GPU parity tests: the CUDA path (through the C ABI, via qibo_b200.engine) against the oracle and the
reference's golden vectors.  Tolerances are the north star's: 1e-12 max-abs complex128, 1e-5 complex64;
samples bit-exact for the same uniforms."""

import os
from collections import Counter

import numpy as np
import pytest
import torch

from conftest import tol
from helpers import ops_from_named, oracle_run, rand_state, random_zoo
from oracle import numpy_oracle as orc
from qibo_b200.ops import Op

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from qibo_b200.engine import Engine

    return Engine(0)


def run_k1(eng, psi, ops, n):
    st = eng.upload(psi)
    for op in ops:
        eng.apply_op(st, n, op)
    return st.numpy()


def run_k2(eng, psi, ops, n, fuse=True):
    st = eng.upload(psi)
    eng.apply_program(st, n, ops, fuse=fuse)
    return st.numpy()


# ------------------------------------------------------------------------------------------ G1 / G2
def test_golden_gates(eng, golden):
    cases = golden.cases("gate_cases")
    for i, c in enumerate(cases):
        psi, ref, mat = golden[f"gate{i}_in"], golden[f"gate{i}_out"], golden[f"gate{i}_matrix"]
        op = Op(mat, tuple(c["targets"]), tuple(c["controls"])) if c["is_controlled_by"] else Op(mat, tuple(c["qubits"]))
        n = c["nqubits"]
        out = run_k1(eng, psi, [op], n)
        assert out.dtype == ref.dtype
        assert np.abs(out - ref).max() < tol(c["dtype"]), ("k1", c["tag"])
        for fuse in (True, False):
            out = run_k2(eng, psi, [op], n, fuse)
            assert np.abs(out - ref).max() < tol(c["dtype"]), ("k2", fuse, c["tag"])


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_zoo_every_bit_position(eng, dtype, seed):
    """Every kernel family (dense k<=5, controls, diagonals, phases, swaps) on targets spanning low (vectorised /
    in-run) and high (cross-tile) bit positions, K1 gate-by-gate and K2 fused."""
    n = 16
    ops = random_zoo(n, 60, seed)
    psi = rand_state(n, seed, dtype)
    ref = oracle_run(psi, ops, n)
    # north-star bounds as they are (profiles/r2a_tol_probe.json: the CUDA path and the reference algorithm in complex64 both
    # sit at ~5e-9 from the complex128 result on these programs)
    assert np.abs(run_k1(eng, psi, ops, n) - ref).max() < tol(dtype)
    assert np.abs(run_k2(eng, psi, ops, n) - ref).max() < tol(dtype)
    assert np.abs(run_k2(eng, psi, ops, n, fuse=False) - ref).max() < tol(dtype)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_single_qubit_gate_on_every_qubit(eng, dtype):
    n = 18
    psi = rand_state(n, 42, dtype)
    rng = np.random.default_rng(5)
    from helpers import rand_unitary

    for q in range(n):
        op = Op(rand_unitary(1, rng), (q,))
        ref = oracle_run(psi, [op], n)
        assert np.abs(run_k1(eng, psi, [op], n) - ref).max() < tol(dtype), q
        assert np.abs(run_k2(eng, psi, [op], n) - ref).max() < tol(dtype), q
    for a, b in [(0, 17), (17, 0), (16, 17), (3, 9), (0, 1), (8, 17)]:
        op = Op(rand_unitary(2, rng), (a, b))
        ref = oracle_run(psi, [op], n)
        assert np.abs(run_k1(eng, psi, [op], n) - ref).max() < tol(dtype), (a, b)
        assert np.abs(run_k2(eng, psi, [op], n) - ref).max() < tol(dtype), (a, b)


# ------------------------------------------------------------------------------------------ circuits
def _ops_for(tag, golden):
    if tag.startswith("qft"):
        return orc.qft_ops(int(tag[3:].split("_")[0]), with_swaps="noswap" not in tag)
    if tag.startswith("var10x3"):
        return orc.variational_ops(10, 3, golden["var_thetas"])
    return orc.random_ops(9, 40, seed=11)


def test_golden_circuits(eng, golden):
    for i, c in enumerate(golden.cases("circ_cases")):
        n, dtype = c["nqubits"], c["dtype"]
        psi = orc.zero_state(n, dtype) if c["zero"] else golden[f"circ{i}_in"]
        if c["queue"] is None:
            ops = ops_from_named(_ops_for(c["tag"], golden))
        else:  # the reference fuser's FusedGate matrices (k <= 5 dense blocks)
            ops = [Op(golden[f"circ{i}_q{j}"], tuple(q)) for j, q in enumerate(c["queue"])]
        ref = golden[f"circ{i}_out"]
        assert np.abs(run_k2(eng, psi, ops, n) - ref).max() < tol(dtype), c["tag"]
        assert np.abs(run_k1(eng, psi, ops, n) - ref).max() < tol(dtype), c["tag"]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [1, 2, 4, 11, 12, 13, 14, 17, 20])
def test_qft_vs_oracle(eng, n, dtype):
    psi = rand_state(n, 100 + n, dtype)
    named = orc.qft_ops(n)
    ref = orc.run_ops(psi, named, n, dtype=dtype)
    ops = ops_from_named(named)
    assert np.abs(run_k2(eng, psi, ops, n) - ref).max() < tol(dtype)
    if n <= 17:
        assert np.abs(run_k1(eng, psi, ops, n) - ref).max() < tol(dtype)
        assert np.abs(run_k2(eng, psi, ops, n, fuse=False) - ref).max() < tol(dtype)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_variational_and_random_vs_oracle(eng, dtype):
    n = 18
    thetas = 2 * np.pi * np.random.default_rng(7).random(2 * 3 * n)
    named = orc.variational_ops(n, 3, thetas)
    psi = rand_state(n, 5, dtype)
    ref = orc.run_ops(psi, named, n, dtype=dtype)
    assert np.abs(run_k2(eng, psi, ops_from_named(named), n) - ref).max() < tol(dtype)
    named = orc.random_ops(n, 80, seed=11)
    ref = orc.run_ops(psi, named, n, dtype=dtype)
    assert np.abs(run_k2(eng, psi, ops_from_named(named), n) - ref).max() < tol(dtype)
    assert np.abs(run_k1(eng, psi, ops_from_named(named), n) - ref).max() < tol(dtype)


@pytest.mark.parametrize("n,dtype", [(24, "complex128"), (26, "complex128"), (27, "complex64"), (30, "complex128")])
def test_qft_full_size_properties(eng, n, dtype):
    """BASELINE sizes where the oracle is too slow: size-independent properties.
    QFT == inverse DFT (checked against torch.fft on the same GPU), QFT|0> is uniform, the norm is kept."""
    from qibo_b200 import circuits

    ops = circuits.qft(n)
    st = eng.basis_state(n, dtype)
    eng.apply_program(st, n, ops)
    expected = 2.0 ** (-n / 2)
    t = st.tensor
    assert float((t.real - expected).abs().max()) < tol(dtype) and float(t.imag.abs().max()) < tol(dtype)
    assert abs(eng.norm2(st) - 1.0) < (1e-9 if dtype == "complex128" else 1e-4)
    if n <= 27:
        g = torch.Generator(device="cuda").manual_seed(n)
        x = torch.randn(2**n, dtype=torch.float64 if dtype == "complex128" else torch.float32, device="cuda", generator=g)
        y = torch.randn(2**n, dtype=x.dtype, device="cuda", generator=g)
        psi = torch.complex(x, y)
        psi /= torch.linalg.vector_norm(psi)
        from qibo_b200.array import DeviceArray

        st = DeviceArray(psi.clone())
        eng.apply_program(st, n, ops)
        ref = torch.fft.ifft(psi, norm="ortho")
        assert float((st.tensor - ref).abs().max()) < tol(dtype)
        # linearity: QFT(a*psi) == a*QFT(psi)
        st2 = DeviceArray(psi * (0.3 - 0.4j))
        eng.apply_program(st2, n, ops)
        assert float((st2.tensor - (0.3 - 0.4j) * st.tensor).abs().max()) < tol(dtype)


@pytest.mark.parametrize("n,dtype", [(30, "complex128"), (32, "complex128"), (31, "complex64")])
@pytest.mark.parametrize("path", ["apply_program", "compiled"])
def test_qft_basis_state_closed_form_at_bench_size(eng, n, dtype, path):
    """BASELINE config 2 (QFT(30)) and the benchmarked QFT(32), phase-sensitive and on EVERY amplitude: QFT of a generic
    basis state |x> against exp(2 pi i x k / 2^n) / sqrt(2^n), evaluated on the device in int64 / float64 chunks
    (qibo_b200/checks.py, itself pinned to the oracle on the CPU).  From |0...0> no CU1 ever fires; from |x> every one of
    the n(n+1)/2 + n/2 gates acts (models/qft.py:47-58).  Both the host-matrix path (qb_apply_program) and the compiled
    program (qb_program_create / qb_program_run) that bench.py times."""
    from qibo_b200 import circuits
    from qibo_b200.checks import generic_basis_state, qft_basis_state_error

    free, _ = eng.mem_info()
    itemsize = 16 if dtype == "complex128" else 8
    if free < 2.2 * itemsize * 2**n:
        pytest.skip("needs the state plus the permutation buffer in device memory")
    ops = circuits.qft(n)
    x = generic_basis_state(n)
    st = eng.basis_state(n, dtype, x)
    if path == "compiled":
        prog = eng.compile(n, dtype, ops)
        eng.run_program(prog, st)
    else:
        eng.apply_program(st, n, ops)
    err = qft_basis_state_error(st.tensor, n, x)
    # relative to the amplitude size 2^(-n/2): 1e-10 is ~1e-15 absolute at n = 30, far inside the 1e-12 north-star bound
    assert err < (1e-10 if dtype == "complex128" else 2e-3), err
    assert abs(eng.norm2(st) - 1.0) < (1e-9 if dtype == "complex128" else 1e-4)
    del st
    torch.cuda.empty_cache()


@pytest.mark.parametrize("case,n,dtype", [("variational", 27, "complex64"), ("random", 26, "complex128"), ("variational", 25, "complex128"),
                                          ("random", 27, "complex64")])
def test_scheduled_sweeps_match_gate_by_gate_at_scale(eng, case, n, dtype):
    """BASELINE configs 3 and 4 at sizes the oracle cannot reach: the sweep path (list-scheduled light-cone sweeps,
    reordered passes, tensor-map copies, swizzled tiles) against the one-gate-per-sweep K1 kernels, which are checked
    against the oracle gate by gate at small n.  Same device, same input, every amplitude compared."""
    from qibo_b200 import circuits
    from qibo_b200.array import DeviceArray

    if case == "variational":
        ops = circuits.variational(n, 4, 2 * np.pi * np.random.default_rng(7).random(2 * 4 * n))
    else:
        ops = circuits.random_circuit(n, 150, seed=11)
    g = torch.Generator(device="cuda").manual_seed(n)
    rdt = torch.float64 if dtype == "complex128" else torch.float32
    psi = torch.complex(torch.randn(2**n, dtype=rdt, device="cuda", generator=g), torch.randn(2**n, dtype=rdt, device="cuda", generator=g))
    psi /= torch.linalg.vector_norm(psi)
    a = DeviceArray(psi.clone())
    stats = eng.apply_program(a, n, ops)
    assert stats.nsweeps < len(ops) / 5
    b = DeviceArray(psi.clone())
    for op in ops:
        eng.apply_op(b, n, op)
    assert float((a.tensor - b.tensor).abs().max()) < tol(dtype)
    assert abs(eng.norm2(a) - 1.0) < (1e-9 if dtype == "complex128" else 1e-4)


@pytest.mark.parametrize("n,world,rank,dtype", [(29, 8, 0, "complex128"), (29, 8, 5, "complex128"), (30, 4, 0, "complex64"),
                                                 (31, 8, 0, "complex128")])
def test_rank_specialised_segments_at_scale(eng, n, world, rank, dtype):
    """What rank `rank` of `world` runs in a sharded QFT(n), on one GPU: the local segments of the distributed plan (fans
    with global qubits folded in, one-stage sweeps whose compute phase is far shorter than a tile load -- the case that
    exposed a stale-parity race between the two compute teams) against the gate-by-gate K1 kernels."""
    from qibo_b200 import circuits, distributed as D
    from qibo_b200.array import DeviceArray

    g = world.bit_length() - 1
    nlocal = n - g
    plan = D.Plan(n, g, circuits.qft(n))
    gen = torch.Generator(device="cuda").manual_seed(n)
    rdt = torch.float64 if dtype == "complex128" else torch.float32
    psi = torch.complex(torch.randn(2**nlocal, dtype=rdt, device="cuda", generator=gen), torch.randn(2**nlocal, dtype=rdt, device="cuda", generator=gen))
    psi /= torch.linalg.vector_norm(psi)
    a, b = DeviceArray(psi.clone()), DeviceArray(psi.clone())
    del psi
    nseg = 0
    for seg in plan.segments:
        if seg.kind != "local":
            continue
        local = [o for o in (D.specialise(p, nlocal, rank) for p in seg.ops) if o is not None]
        if not local:
            continue
        for _ in range(3 if len(local) < 40 else 1):  # short sweeps: repeat (U^3 on both sides) to give a race its chance
            eng.apply_program(a, nlocal, local)
            for op in local:
                eng.apply_op(b, nlocal, op)
        nseg += 1
        assert float((a.tensor - b.tensor).abs().max()) < tol(dtype)
    assert nseg >= 4


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [2, 3, 9, 16, 21])
def test_compiled_program_equals_apply_program(eng, n, dtype):
    """qb_program_create / qb_program_run: the same queue planned once and launched repeatedly gives bit-identical
    amplitudes to qb_apply_program on every launch, on different states, and next to other programs."""
    from helpers import random_zoo
    from qibo_b200 import circuits

    for ops in (circuits.qft(n),) + ((random_zoo(n, 30, seed=n),) if n >= 9 else ()):
        prog = eng.compile(n, dtype, ops)
        other = eng.compile(n, dtype, list(reversed(ops)))
        for seed in (1, 2):
            psi = rand_state(n, seed, dtype)
            a, b = eng.upload(psi), eng.upload(psi)
            s1 = eng.apply_program(a, n, ops)
            s2 = eng.run_program(prog, b, timed=True)
            assert np.array_equal(a.numpy(), b.numpy())
            assert s2.nsweeps == s1.nsweeps == prog.nsweeps and s2.nops == len(ops)
            eng.run_program(other, b)
            eng.apply_program(a, n, list(reversed(ops)))
            assert np.array_equal(a.numpy(), b.numpy())
        prog.close()
        other.close()
    with pytest.raises(ValueError):
        eng.run_program(eng.compile(n, dtype, circuits.qft(n)), eng.upload(rand_state(n + 1, 0, dtype)))


# ------------------------------------------------------------------------------------------ P1
def test_probabilities_golden(eng, golden):
    for i, c in enumerate(golden.cases("prob_cases")):
        st = eng.upload(golden[f"prob{i}_in"])
        out = eng.probabilities(st, c["qubits"], c["nqubits"]).numpy()
        ref = golden[f"prob{i}_out"]
        assert out.dtype == ref.dtype and out.shape == ref.shape
        assert np.abs(out - ref).max() < (1e-14 if c["dtype"] == "complex128" else 1e-6), c


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [3, 5, 9, 16, 20])
def test_probabilities_vs_oracle(eng, n, dtype):
    psi = rand_state(n, n, dtype)
    st = eng.upload(psi)
    rng = np.random.default_rng(n)
    subsets = [list(range(n)), [0], [n - 1], list(range(n - 1, -1, -1))]
    for _ in range(6):
        m = int(rng.integers(1, n + 1))
        subsets.append(rng.permutation(n)[:m].tolist())
    for qubits in subsets:
        ref = orc.calculate_probabilities(psi, qubits, n)
        out = eng.probabilities(st, qubits, n).numpy()
        assert out.dtype == ref.dtype
        if dtype == "complex128":
            assert np.abs(out - ref).max() < 1e-14, qubits
        else:
            # the reference accumulates 2^(n-m) float32 terms in float32 (error ~1e-5 at n=20); the kernel accumulates
            # in double and rounds once, so it is compared with the exactly-accumulated value, and loosely with float32
            exact = orc.calculate_probabilities(psi.astype(np.complex128), qubits, n)
            assert np.abs(out - exact).max() < 2e-7, qubits
            assert np.abs(out - ref).max() < 1e-4, qubits
        assert abs(out.sum() - 1) < 1e-5


# ------------------------------------------------------------------------------------------ S1 / S2
def test_sampling_golden_bit_exact(eng, golden):
    from qibo_b200 import _lib

    for i, c in enumerate(golden.cases("samp_cases")):
        p = golden[f"samp{i}_probs"]
        np.random.seed(c["seed"])
        u = np.random.random_sample(c["nshots"])
        out = eng.sample(eng.upload(p), u, mode=_lib.QB_SCAN_EXACT)
        np.testing.assert_array_equal(out, golden[f"samp{i}_shots"])
        assert out.dtype == np.int64
        # float32 probabilities (complex64 states) follow the same contract: promoted to double, then scanned
        p32 = p.astype(np.float32)
        out32 = eng.sample(eng.upload(p32), u, mode=_lib.QB_SCAN_EXACT)
        np.testing.assert_array_equal(out32, orc.choice_from_uniforms(p32, u))


def test_sampling_search_is_exact_for_any_cdf(eng):
    """Contract (i) of SURVEY 8a hazard 1: idx = #{k: cdf[k] <= u} bit-exactly, for both scan modes."""
    from qibo_b200 import _lib

    rng = np.random.default_rng(1)
    for nbins in (1, 2, 5, 4096, 4097, 2**20 + 3):
        p = rng.random(nbins)
        p /= p.sum()
        u = rng.random(20000)
        u[:3] = [0.0, 0.5, np.nextafter(1.0, 0.0)]
        for mode in (_lib.QB_SCAN_EXACT, _lib.QB_SCAN_PARALLEL):
            dp = eng.upload(p)
            cdf = eng.cdf(dp, mode).numpy()
            out = eng.sample(dp, u, mode=mode)
            np.testing.assert_array_equal(out, np.searchsorted(cdf, u, side="right"))
            assert cdf[-1] == 1.0 and np.all(np.diff(cdf) >= -1e-15)
            if mode == _lib.QB_SCAN_EXACT:
                ref = np.cumsum(p)
                ref /= ref[-1]
                np.testing.assert_array_equal(cdf, ref)  # numpy-exact scan
            else:
                assert np.abs(cdf - np.cumsum(p) / p.sum()).max() < 1e-13
    # edge of the distribution: zero-probability bins are never drawn
    p = np.array([0.0, 0.5, 0.0, 0.5, 0.0])
    out = eng.sample(eng.upload(p), rng.random(1000))
    assert set(out.tolist()) <= {1, 3}


# ------------------------------------------------------------------------------------------ C1
def test_collapse_golden(eng, golden):
    for i, c in enumerate(golden.cases("coll_cases")):
        st = eng.upload(golden[f"coll{i}_in"])
        eng.collapse(st, c["nqubits"], c["qubits"], c["shot"], c["normalize"])
        assert np.abs(st.numpy() - golden[f"coll{i}_out"]).max() < 1e-14, c


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_collapse_vs_oracle_and_idempotence(eng, dtype):
    n = 17
    psi = rand_state(n, 3, dtype)
    for qubits, shot in (([0], 1), ([16], 0), ([3, 9, 16], 5), (list(range(0, 17, 2)), 137)):
        st = eng.upload(psi)
        eng.collapse(st, n, qubits, shot, True)
        ref = orc.collapse_statevector(psi, qubits, [shot], n, True)
        assert np.abs(st.numpy() - ref).max() < (1e-13 if dtype == "complex128" else 1e-5)
        assert abs(eng.norm2(st) - 1.0) < (1e-12 if dtype == "complex128" else 1e-5)
        once = st.numpy().copy()
        eng.collapse(st, n, qubits, shot, True)  # projecting twice changes nothing
        assert np.abs(st.numpy() - once).max() < (1e-15 if dtype == "complex128" else 1e-6)


# ------------------------------------------------------------------------------------------ K6 / misc
def test_state_constructors_and_cast(eng):
    for dtype in ("complex128", "complex64"):
        st = eng.basis_state(10, dtype, 5)
        ref = np.zeros(1024, dtype=dtype)
        ref[5] = 1
        np.testing.assert_array_equal(st.numpy(), ref)
        st = eng.filled_state(10, 1 / 32.0, dtype)
        np.testing.assert_array_equal(st.numpy(), np.full(1024, 1 / 32.0, dtype=dtype))
    psi = rand_state(12, 1)
    d = eng.upload(psi)
    np.testing.assert_array_equal(eng.cast(d, "complex64").numpy(), psi.astype("complex64"))
    np.testing.assert_array_equal(eng.cast(eng.cast(d, "complex64"), "complex128").numpy(), psi.astype("complex64").astype("complex128"))
    assert abs(eng.norm2(d) - 1.0) < 1e-13
    assert isinstance(d.tensor, torch.Tensor) and d.tensor.is_cuda  # the state buffer is a torch tensor too


def test_error_mapping(eng):
    st = eng.basis_state(4)
    with pytest.raises(ValueError):
        eng.apply_op(st, 4, Op(np.eye(2), (7,)))
    with pytest.raises(ValueError):
        eng.probabilities(st, [0, 0], 4)
    with pytest.raises(ValueError):
        eng.collapse(st, 4, [1], 2)


# ------------------------------------------------------------------------------------------ K8
@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [1, 2, 5, 12, 13, 18, 21])
def test_permute_qubits(eng, n, dtype):
    """Out-of-place permutation sweep == the axis transposition NumPy does; a run of SWAP gates is one such sweep."""
    rng = np.random.default_rng(n)
    psi = rand_state(n, n, dtype)
    perms = [list(range(n - 1, -1, -1))] + [rng.permutation(n).tolist() for _ in range(3)]
    for dest in perms:
        st = eng.upload(psi)
        eng.permute_qubits(st, n, dest)
        ref = np.moveaxis(psi.reshape(n * (2,)), list(range(n)), dest).reshape(-1)
        np.testing.assert_array_equal(st.numpy(), ref)
    if n >= 6:
        named = [("SWAP", (q, n - 1 - q), ()) for q in range(n // 2)]
        st = eng.upload(psi)
        stats = eng.apply_program(st, n, ops_from_named(named))
        assert stats.nsweeps == 1
        np.testing.assert_array_equal(st.numpy(), orc.run_ops(psi, named, n, dtype=dtype))
        eng.permute_swap_runs = False
        try:
            st = eng.upload(psi)
            eng.apply_program(st, n, ops_from_named(named))
            np.testing.assert_array_equal(st.numpy(), orc.run_ops(psi, named, n, dtype=dtype))
        finally:
            eng.permute_swap_runs = True


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_six_qubit_dense_block(eng, dtype):
    from helpers import rand_unitary

    n = 15
    rng = np.random.default_rng(6)
    ops = [Op(rand_unitary(6, rng), (13, 2, 7, 0, 9, 4)), Op(orc.gate_matrix("H"), (3,)), Op(rand_unitary(6, rng), (1, 5, 3, 8, 12, 6), (10,))]
    psi = rand_state(n, 3, dtype)
    ref = oracle_run(psi, ops, n)
    assert np.abs(run_k2(eng, psi, ops, n) - ref).max() < tol(dtype)
    assert np.abs(run_k1(eng, psi, ops, n) - ref).max() < tol(dtype)  # apply_op routes k = 6 to the sweep kernel


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_pauli_expectation_and_vdot(eng, dtype):
    """K9 against the reference's golden expectation values and against the oracle on random strings (incl. strings
    that flip the highest and the lowest bit, identity factors, and many Y factors for the i^nY bookkeeping)."""
    import json

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "expval_golden.npz"))
    t = 1e-12 if dtype == "complex128" else 2e-6
    for i, c in enumerate(json.loads(str(z["cases"]))):
        n = c["nqubits"]
        st = eng.upload(z[f"ev{i}_state"].astype(dtype))
        got = [eng.expval_pauli(st, n, term, q) for term, q in zip(c["terms"], c["term_qubits"])]
        assert np.abs(np.real(got) - z[f"ev{i}_per_term"]).max() < t
        assert np.abs(np.imag(got)).max() < t
        other = eng.upload(z[f"ev{i}_other"].astype(dtype))
        assert abs(eng.vdot(st, other, n) - complex(z[f"ev{i}_overlap"])) < t
    n = 18
    psi = rand_state(n, 5, dtype)
    st = eng.upload(psi)
    rng = np.random.default_rng(9)
    strings = [("X", [0]), ("Y", [n - 1]), ("Z", [3]), ("XYZ", [0, n - 1, 7]), ("YYYY", [1, 2, 3, 4]), ("IZI", [5, 6, 7]), ("", [])]
    for _ in range(10):
        k = int(rng.integers(1, 7))
        strings.append(("".join(rng.choice(list("IXYZ"), size=k)), [int(q) for q in rng.choice(n, size=k, replace=False)]))
    for term, qubits in strings:
        want = orc.pauli_expectation(psi.astype(np.complex128), term, qubits, n)
        assert abs(eng.expval_pauli(st, n, term, qubits) - want) < t, (term, qubits)
    with pytest.raises(ValueError):
        eng.expval_pauli(st, n, "XQ", [0, 1])
    with pytest.raises(ValueError):
        eng.expval_pauli(st, n, "XX", [1, 1])


def test_one_backend_entered_from_two_threads(eng):
    """parallel.py:53 runs one backend object from joblib THREADS: two threads drive the same engine (one context, one
    stream, per-context mutex in the library) on their own states at once -- every result must equal the single-threaded
    one bit for bit."""
    import threading

    from qibo_b200 import circuits

    n = 18
    progs = [circuits.qft(n), random_zoo(n, 40, 3), circuits.variational(n, 2, np.random.default_rng(3).random(4 * n))]
    psis = [rand_state(n, s) for s in range(3)]
    want = [run_k2(eng, psi, ops, n) for psi, ops in zip(psis, progs)]
    wantp = [eng.probabilities(eng.upload(w), [0, 3, 5], n).numpy() for w in want]
    errors = []

    def work(i, reps):
        try:
            for _ in range(reps):
                st = eng.upload(psis[i])
                eng.apply_program(st, n, progs[i])
                if not np.array_equal(st.numpy(), want[i]):
                    errors.append(("state", i))
                if not np.array_equal(eng.probabilities(st, [0, 3, 5], n).numpy(), wantp[i]):
                    errors.append(("probs", i))
                u = np.random.default_rng(i).random(64)
                eng.sample(eng.probabilities(st, list(range(n)), n), u)
        except Exception as exc:  # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=work, args=(i % 3, 6)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(300)
    assert not errors, errors[:3]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_wide_unitaries(eng, dtype):
    """Unitary on more than 6 targets (the reference takes any width, gates/gates.py:2774; gates.I(*range(7)) too):
    K8 permutation + library GEMM + K8 back (Engine.apply_wide), alone, controlled, and inside a longer program."""
    from helpers import rand_unitary

    n = 13
    rng = np.random.default_rng(8)
    psi = rand_state(n, 4, dtype)
    wide7 = Op(rand_unitary(7, rng), (12, 0, 5, 3, 9, 1, 7))
    wide8c = Op(rand_unitary(8, rng), (2, 4, 6, 8, 10, 11, 0, 1), (5, 12))
    ident = Op(np.eye(128), tuple(range(7)))
    t = tol(dtype)
    for op in (wide7, wide8c, ident):
        assert np.abs(run_k1(eng, psi, [op], n) - oracle_run(psi, [op], n)).max() < t
    ops = [Op(orc.gate_matrix("H"), (3,)), wide7, Op(orc.gate_matrix("CNOT"), (0, 12)), ident, wide8c, Op(orc.gate_matrix("RY", 0.3), (6,))]
    ref = oracle_run(psi, ops, n)
    assert np.abs(run_k2(eng, psi, ops, n) - ref).max() < t
    prog = eng.compile(n, dtype, ops)
    st = eng.upload(psi)
    eng.run_program(prog, st)
    assert np.abs(st.numpy() - ref).max() < t


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_compiled_program_parameter_slots(eng, dtype):
    """qb_program_set_params through the C ABI: angles for the gate families, host matrices for everything else, on a
    program that also holds a SWAP run (K8 segment) -- against the oracle on the updated ops."""
    from qibo_b200 import _lib

    n = 15
    rng = np.random.default_rng(2)

    def build(th):
        named = []
        for q in range(n):
            named.append(("RY", (q,), (th[q],)))
        for q in range(0, n - 1, 2):
            named.append(("CZ", (q, q + 1), ()))
        for q in range(n):
            named.append(("RX", (q,), (th[n + q],)))
        for q in range(1, n - 1, 2):
            named.append(("CU1", (q, q + 1), (th[2 * n + q],)))
        named.append(("CRZ", (0, n - 1), (th[3 * n],)))
        ops = ops_from_named(named)
        ops += ops_from_named([("SWAP", (q, n - 1 - q), ()) for q in range(3)])
        ops += ops_from_named([("RZ", (q,), (th[3 * n + 1 + q],)) for q in range(n)])
        return named, ops

    th0 = rng.uniform(0.1, 6, 4 * n + 2)
    named, ops = build(th0)
    prog = eng.compile(n, dtype, ops)
    psi = rand_state(n, 1, dtype)
    fam = {"RY": _lib.QB_GATE_RY, "RX": _lib.QB_GATE_RX, "CU1": _lib.QB_GATE_CU1, "CRZ": _lib.QB_GATE_CRZ, "RZ": _lib.QB_GATE_RZ}
    for step in range(3):
        th = rng.uniform(0.1, 6, 4 * n + 2)
        _, new_ops = build(th)
        updates = []
        for i, (old, new) in enumerate(zip(ops, new_ops)):
            if old.name in fam:
                if step == 1 and i % 2:  # every other one as a host matrix
                    updates.append((i, _lib.QB_GATE_MATRIX, [], new.data, False))
                else:
                    theta = {"RY": None}.get("x")
                    updates.append((i, fam[old.name], [_theta_of(new)], None, False))
        prog.set_params(updates)
        st = eng.upload(psi)
        eng.run_program(prog, st)
        assert np.abs(st.numpy() - oracle_run(psi, new_ops, n)).max() < tol(dtype), step
    prog.close()


def _theta_of(op):
    """Angle of an RX / RY / RZ / CU1 / CRZ op from its matrix (test helper)."""
    m = op.data
    if op.name == "RY":
        return 2 * np.arctan2(m[1, 0].real, m[0, 0].real)
    if op.name == "RX":
        return 2 * np.arctan2(-m[0, 1].imag, m[0, 0].real)
    if op.name == "RZ":
        return 2 * np.angle(m[1, 1])
    if op.name == "CU1":
        return np.angle(m[3, 3])
    if op.name == "CRZ":
        return 2 * np.angle(m[3, 3])
    raise KeyError(op.name)


# ------------------------------------------------------------------------------------------ permuting sweeps
@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("seed", range(4))
def test_trailing_permutation_rides_on_last_sweep(eng, dtype, seed):
    """qb_apply_program_permuted / qb_program_create_permuted + run: gates followed by a SWAP run, the permutation written
    by the last sweep (out of place, destination tensor map) -- against the oracle; the two-launch form (K8) must agree
    bit for bit; with and without a caller-provided second buffer."""
    rng = np.random.default_rng(50 + seed)
    n = int(rng.integers(14, 21))
    psi = rand_state(n, seed, dtype)
    keep = int(rng.integers(0, 4))
    dests = [list(range(n - 1, -1, -1)), list(range(keep)) + list(range(n - 1, keep - 1, -1)), rng.permutation(n).tolist()]
    from qibo_b200.engine import swaps_for_permutation

    for dest in dests:
        for ngates in (0, 30):
            gops = random_zoo(n, ngates, seed) if ngates else []
            ops = gops + swaps_for_permutation(dest)
            if len(ops) - len(gops) < 3:
                continue
            ref = oracle_run(psi, ops, n)
            st = eng.upload(psi)
            stats = eng.apply_program(st, n, ops)
            assert np.abs(st.numpy() - ref).max() < tol(dtype), (n, dest, ngates)
            fused = stats.nperm_fused
            # the two-launch form
            eng.fuse_permutations = False
            try:
                st2 = eng.upload(psi)
                stats2 = eng.apply_program(st2, n, ops)
            finally:
                eng.fuse_permutations = True
            assert stats2.nperm_fused == 0
            if fused:
                assert stats.nperm == 0 and stats2.nperm == 1 and stats.nsweeps <= stats2.nsweeps
            assert np.abs(st2.numpy() - ref).max() < tol(dtype)
            # compiled, twice, ping-ponging between two caller-owned buffers
            prog = eng.compile(n, dtype, ops)
            a, b = eng.upload(psi), eng.empty((1 << n,), dtype)
            eng.run_program(prog, a, alt=b)
            assert np.array_equal(a.numpy(), st.numpy())
            b.tensor.copy_(eng.upload(psi).tensor)
            eng.run_program(prog, b, alt=a)
            assert np.array_equal(b.numpy(), st.numpy())
            prog.close()


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_qft_with_fused_reversal_vs_oracle(eng, dtype):
    from qibo_b200 import circuits

    for n in (13, 14, 15, 18, 21):
        psi = rand_state(n, n, dtype)
        ref = orc.run_ops(psi, orc.qft_ops(n), n, dtype=dtype)
        st = eng.upload(psi)
        stats = eng.apply_program(st, n, circuits.qft(n))
        assert stats.nperm_fused == (1 if n > (12 if dtype == "complex128" else 13) else 0)
        assert np.abs(st.numpy() - ref).max() < tol(dtype)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_first_sweep_makes_the_zero_state(eng, dtype):
    """QB_PROGRAM_INPUT_ZERO: the buffer handed in is uninitialised (filled with NaNs here) and stands for |0...0>; the
    first sweep's loader makes the tiles.  Bit-identical to running on a written zero state -- compiled and uncompiled,
    with a fused permutation, for programs that start with a permutation / have no gate at all, and below 4 qubits."""
    from qibo_b200 import circuits
    from qibo_b200.engine import swaps_for_permutation

    rng = np.random.default_rng(9)
    cases = []
    for n in (2, 3, 9, 14, 17, 21):
        cases.append((n, circuits.qft(n)))
        if n >= 9:
            cases.append((n, random_zoo(n, 25, n)))
    cases.append((16, swaps_for_permutation(list(range(15, -1, -1))) + circuits.qft(16, with_swaps=False)))  # permutation first
    cases.append((15, []))
    cases.append((15, [Op(np.eye(2), (3,))]))  # canonicalises to nothing: no sweep runs
    for n, ops in cases:
        ref_state = eng.basis_state(n, dtype)
        eng.apply_program(ref_state, n, ops)
        ref = ref_state.numpy()
        for compiled in (False, True):
            st = eng.uninitialised_state(n, dtype)
            st.tensor.fill_(float("nan"))
            if compiled:
                prog = eng.compile(n, dtype, ops)
                eng.run_program(prog, st, input_zero=True)
                prog.close()
            else:
                eng.apply_program(st, n, ops, input_zero=True)
            out = st.numpy()
            assert not np.isnan(out).any(), (n, len(ops), compiled)
            assert np.array_equal(out, ref), (n, len(ops), compiled)
