"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/qibo_b200.h
declares, refuses to run without a GPU (no CPU fallback), and its host-only planner works."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from qibo_b200 import _lib, circuits
from qibo_b200.engine import plan_program
from qibo_b200.ops import Op

HAS_GPU = torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "qibo_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(qb_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 25
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.qb_version() == 100


def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 (no C++ in the signatures) and a C program that calls an entry
    point links against the shared library."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include "qibo_b200.h"\n'
                   "int main(void) { qb_program p = 0; qb_handle h = 0; (void)p; (void)h; return qb_version() == 100 ? 0 : 1; }\n")
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", str(src), "-I", os.path.join(root, "include"),
                    "-L", libdir, "-lqibo_b200", "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_struct_layout_matches_header():
    # qb_op: 2 + 6 + 32 + 2 int32 then a pointer; qb_program_stats: 4 int32, double, float, 3 int32
    assert ctypes.sizeof(_lib.QbOp) == 4 * (2 + 6 + 32 + 2) + 8
    assert ctypes.sizeof(_lib.QbProgramStats) == 40


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.qb_create(0, None, ctypes.byref(h))
    assert rc == _lib.QB_ERR_CUDA
    assert "no CPU fallback" in _lib.last_error()
    from qibo_b200.engine import Engine

    with pytest.raises(_lib.QiboB200Error):
        Engine(0)


def _assert_dependencies_kept(ops, sweep_of_op):
    """The scheduler may move a gate to an earlier sweep only past gates it commutes with: two gates that share a
    qubit and are not both diagonal must keep their order."""
    def diag_qubits(op):
        """qubit -> True when the gate never mixes its 0 and 1 subspaces (controls, diagonal targets)"""
        out = {q: True for q in op.controls}
        k = len(op.targets)
        m = np.diag(op.data) if op.is_diagonal else np.asarray(op.data)
        idx = np.arange(2**k)
        for pos, q in enumerate(op.targets):
            bit = (idx >> (k - 1 - pos)) & 1
            out[q] = not np.any(m[bit[:, None] != bit[None, :]])
        return out

    last = {}  # qubit -> list of (sweep, acts diagonally on it) of the gates seen so far
    for op, s in zip(ops, sweep_of_op):
        for q, d in diag_qubits(op).items():
            for s0, d0 in last.get(q, []):
                if not (d and d0):
                    assert s0 <= s
            last.setdefault(q, []).append((s, d))


def test_planner_packs_qft_into_few_sweeps():
    for n, dtype in ((30, "complex128"), (32, "complex128"), (31, "complex64")):
        ops = circuits.qft(n)
        stats, sweep_of_op = plan_program(n, dtype, ops)
        assert stats.nops == len(ops) == n * (n + 1) // 2 + n // 2
        assert stats.nsweeps <= 20, stats.nsweeps  # gate-by-gate would be ~500 sweeps
        _assert_dependencies_kept(ops, sweep_of_op)
        itemsize = 16 if dtype == "complex128" else 8
        assert stats.bytes_moved == stats.nsweeps * 2 * itemsize * 2.0**n
        stats1, _ = plan_program(n, dtype, ops, fuse=False)
        assert stats1.nsweeps == len(ops)
    # the QFT stages without the closing SWAPs (the engine turns those into one K8 permutation): 4 sweeps at 32 qubits,
    # every one a straight-line stage sweep -- what the lean two-team kernel instantiation runs
    stats, _ = plan_program(32, "complex128", circuits.qft(32, with_swaps=False))
    assert stats.nsweeps == 4 and stats.nstage_sweeps == 4
    stats, _ = plan_program(30, "complex128", circuits.qft(30, with_swaps=False))
    assert stats.nsweeps == 4 and stats.nstage_sweeps == 4


def test_planner_follows_light_cones():
    """Layered circuits: a sweep keeps taking gates of later layers as long as they commute with what stays behind."""
    n = 32
    ops = circuits.variational(n, 20, np.random.default_rng(7).random(2 * 20 * n))
    stats, sweep_of_op = plan_program(n, "complex64", ops)
    assert all(s >= 0 for s in sweep_of_op)
    _assert_dependencies_kept(ops, sweep_of_op)
    assert stats.nsweeps <= 60, stats.nsweeps  # prefix-greedy packing needs 149
    ops = circuits.random_circuit(30, 300, 11)
    stats, sweep_of_op = plan_program(30, "complex128", ops)
    _assert_dependencies_kept(ops, sweep_of_op)
    assert stats.nsweeps <= 40, stats.nsweeps  # prefix-greedy packing needs 53


def test_planner_argument_errors():
    h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    with pytest.raises(ValueError):
        plan_program(4, "complex128", [Op(h, (4,))])
    with pytest.raises(ValueError):
        plan_program(4, "complex128", [Op(np.eye(4), (1, 1))])
    with pytest.raises(ValueError):
        Op(np.eye(4), (1,))
    # blocks on more than 6 targets exist as Ops (Engine.apply_wide: permute + GEMM) but never enter a sweep program
    wide = Op(np.eye(128), tuple(range(7)))
    with pytest.raises(NotImplementedError):
        plan_program(8, "complex128", [wide])
    from qibo_b200.engine import split_segments

    assert [k for k, _ in split_segments([Op(h, (0,)), wide, Op(h, (1,))], 8)] == ["ops", "wide", "ops"]


def test_product_circuits_match_oracle_generators():
    """qibo_b200.circuits (product, used by bench/smoke) describes the same circuits as the pinned oracle."""
    from oracle import numpy_oracle as orc

    def same(ops, named):
        assert len(ops) == len(named)
        for op, (name, qubits, params) in zip(ops, named):
            assert op.targets == tuple(qubits)
            np.testing.assert_array_equal(op.data, orc.gate_matrix(name, *params))

    same(circuits.qft(9), orc.qft_ops(9))
    th = np.random.default_rng(3).random(2 * 2 * 6)
    same(circuits.variational(6, 2, th), orc.variational_ops(6, 2, th))
    same(circuits.random_circuit(7, 25, 11), orc.random_ops(7, 25, 11))


def test_product_does_not_import_oracle():
    import subprocess
    import sys

    code = "import sys; import qibo_b200, qibo_b200.engine, qibo_b200.circuits, qibo_b200.ops; assert not any(m.startswith('oracle') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for root, _, files in os.walk(os.path.join(ROOT, "qibo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp")):
                assert "oracle" not in open(os.path.join(root, f)).read().replace("no oracle", ""), f
