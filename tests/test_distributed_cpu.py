"""World-size 2 and 4 gloo tests (CPU) of the distributed HOST logic: planning (which qubits must be local,
exchanges, swap relabelling, canonicalisation), per-rank specialisation of gates with global controls /
diagonal qubits, and the pairwise half-shard exchange.  The local gate runs are executed by the oracle on
NumPy shards (a test hook); on the GPU the same plan drives the sweep kernels (tests/test_gpu_distributed.py)."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_apply(tensor, nlocal, ops):
    from helpers import oracle_run

    arr = tensor.numpy()
    arr[:] = oracle_run(arr.copy(), ops, nlocal)


def _worker(rank, world, port, case, n, seed, layout, out):
    sys.path[:0] = [ROOT, HERE]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import ops_from_named, oracle_run, rand_state, random_zoo
        from oracle import numpy_oracle as orc
        from qibo_b200.distributed import ShardedProgram

        if case == "qft":
            ops = ops_from_named(orc.qft_ops(n))
        elif case == "random":
            ops = ops_from_named(orc.random_ops(n, 30, seed=seed))
        elif case == "variational":
            ops = ops_from_named(orc.variational_ops(n, 2, np.random.default_rng(seed).random(4 * n) * 6))
        else:
            ops = random_zoo(n, 40, seed, max_dense=3)
        psi = rand_state(n, seed)
        kw = layout if isinstance(layout, dict) else dict(global_qubits=layout)
        prog = ShardedProgram(None, n, "complex128", ops, apply=_oracle_apply, staging_elems=8, **kw)
        nl = prog.nlocal
        if layout is None:
            assert np.array_equal(prog.shard_of(psi), psi[rank << nl : (rank + 1) << nl])
        shard = torch.from_numpy(prog.shard_of(psi).copy())
        r0, loc0 = prog.locate(5)
        basis = np.zeros(2**n)
        basis[5] = 1
        assert (prog.shard_of(basis)[loc0] == 1) == (r0 == rank) and prog.shard_of(basis).sum() == (1 if r0 == rank else 0)
        stats = prog.run(shard, timed=False)
        full = prog.gather(shard)
        ref = oracle_run(psi, ops, n)
        err = float(np.abs(full - ref).max())
        if rank == 0:
            out.put((err, stats.nexchanges, prog.plan.nexchanges))
    finally:
        dist.destroy_process_group()


def _run(world, case, n, seed=0, layout=None):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, n, seed, layout, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return out.get()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case", ["qft", "random", "variational", "zoo"])
def test_sharded_program_matches_oracle(world, case):
    err, nex, planned = _run(world, case, 7 if case != "qft" else 8, seed=3)
    assert err < 1e-12
    assert nex == planned


@pytest.mark.parametrize("world,case,layout", [(2, "qft", "auto"), (4, "qft", "auto"), (4, "zoo", "auto"), (4, "qft", (5, 2)),
                                               (2, "variational", (3,)), (4, "random", (7, 0)),
                                               (4, "qft", dict(global_qubits="auto", final_global_qubits=(0, 1))),
                                               (2, "zoo", dict(global_qubits=(4,), final_global_qubits=(1,)))])
def test_sharded_program_other_layouts(world, case, layout):
    """Global qubits other than the leading ones (cyclic layout for the QFT, arbitrary sets): shard_of / gather map the
    canonical state onto the ranks, and the plan returns to the same layout."""
    err, nex, planned = _run(world, case, 8, seed=4, layout=layout)
    assert err < 1e-12
    assert nex == planned
    if case == "qft" and layout == "auto":
        assert planned == world.bit_length() - 1
    if case == "qft" and isinstance(layout, dict):  # in through the cyclic layout, out in the block layout: two all-to-alls
        assert planned == 2 * (world.bit_length() - 1)  # one exchange per global qubit, as the reference's _DistributedQFT


def test_cyclic_layout_qft_plan():
    """QFT with the trailing qubits global (models/qft.py:66): one exchange per global qubit, back to back, so that the
    closing stages share one local segment; every rank-specialised segment reproduces the oracle through the product's
    planner + passes (tests/emul)."""
    sys.path[:0] = [ROOT, HERE]
    import emul
    from helpers import oracle_run, rand_state
    from qibo_b200 import circuits
    from qibo_b200.distributed import Plan, choose_layout, cyclic_layout, specialise

    for n, g in ((33, 1), (34, 2), (35, 3)):
        plan = choose_layout(n, g, circuits.qft(n))
        assert plan.nexchanges == g and plan.global_qubits in (cyclic_layout(n, g), cyclic_layout(n, g)[::-1])
        kinds = [s.kind for s in plan.segments]
        assert kinds == ["local"] + ["exchange"] * g + ["local"]
    # at the bench size every rank runs stages 0..31 as the same four stage-only sweeps (the phases of global control
    # qubits that are 1 on a rank fold into the fans instead of staying separate ops)
    from qibo_b200.engine import plan_program, split_swap_runs

    n, g = 35, 3
    plan = choose_layout(n, g, circuits.qft(n))
    for r in (0, 3, 5, 7):
        local = [o for o in (specialise(p, n - g, r) for p in plan.segments[0].ops) if o is not None]
        stats, _ = plan_program(n - g, "complex128", local)
        assert stats.nsweeps == 4 and stats.nstage_sweeps == 4, (r, stats.nsweeps, stats.nstage_sweeps)
        tail = [o for o in (specialise(p, n - g, r) for p in plan.segments[-1].ops) if o is not None]
        kinds = [k for k, _ in split_swap_runs(tail, n - g)]
        assert kinds == ["ops", "perm"]
        stats, _ = plan_program(n - g, "complex128", split_swap_runs(tail, n - g)[0][1])
        assert stats.nsweeps == 1
    n, g = 17, 3
    nl = n - g
    plan = Plan(n, g, circuits.qft(n), global_qubits=cyclic_layout(n, g))
    assert plan.nexchanges == g
    for r in range(1 << g):
        for si, seg in enumerate(plan.segments):
            if seg.kind != "local":
                continue
            local = [o for o in (specialise(p, nl, r) for p in seg.ops) if o is not None]
            psi = rand_state(nl, si + r)
            out, stats = emul.apply_program(psi, nl, local)
            assert np.abs(out - oracle_run(psi, local, nl)).max() < 1e-12
            if si == 0:
                assert stats.nsweeps <= 3  # the global controls fold into the fans: no extra ops on the other ranks


def test_alltoall_entries_equal_sequential_exchanges():
    """A run of exchanges as one all-to-all of chunks (K7b, alltoall_entries) moves every amplitude where the pairwise
    half-shard exchanges would, one after the other -- all ranks simulated in one process."""
    sys.path[:0] = [ROOT, HERE]
    from qibo_b200.distributed import Segment, alltoall_entries, alltoall_push_entries, exchange_runs

    rng = np.random.default_rng(0)
    for g, nlocal, pairs in ((3, 6, [(8, 3), (7, 4), (6, 5)]), (3, 5, [(5, 4), (7, 2), (6, 3)]), (2, 4, [(5, 3), (4, 2)]),
                             (1, 3, [(3, 2)]), (3, 5, [(6, 4), (5, 3)])):
        W = 1 << g
        shards = [rng.normal(size=1 << nlocal) for _ in range(W)]
        # reference: the pairwise exchanges in sequence (exchange_half semantics: rank r, bit value b of global bit j,
        # trades its half with local bit == 1 - b against the partner's half with local bit == b)
        ref = [s.copy() for s in shards]
        for gbit, lbit in pairs:
            j = gbit - nlocal
            new = [s.copy() for s in ref]
            for r in range(W):
                b = (r >> j) & 1
                idx = np.arange(1 << nlocal)
                mine = idx[((idx >> lbit) & 1) == 1 - b]
                new[r][mine] = ref[r ^ (1 << j)][mine ^ (1 << lbit)]
            ref = new
        got = [s.copy() for s in shards]
        for r in range(W):
            ent = alltoall_entries(r, nlocal, pairs)
            assert ent is not None and len(ent) == (1 << len(pairs)) - 1
            assert len({e[0] for e in ent}) == len(ent)  # every peer once
            for r2, a, b_, lo, hi in ent:
                x = got[r][a + lo : a + hi].copy()
                got[r][a + lo : a + hi] = got[r2][b_ + lo : b_ + hi]
                got[r2][b_ + lo : b_ + hi] = x
        for r in range(W):
            np.testing.assert_array_equal(got[r], ref[r])
        # out-of-place form: every chunk is copied into the destination rank's second buffer, each slot written once
        second = [np.full(1 << nlocal, np.nan) for _ in range(W)]
        for r in range(W):
            ent = alltoall_push_entries(r, nlocal, pairs)
            assert len(ent) == 1 << len(pairs)
            for r2, a, b_, lo, hi in ent:
                assert np.isnan(second[r2][b_ + lo : b_ + hi]).all()
                second[r2][b_ + lo : b_ + hi] = shards[r][a + lo : a + hi]
        for r in range(W):
            np.testing.assert_array_equal(second[r], ref[r])
    assert alltoall_entries(0, 6, [(8, 3), (7, 5)]) is None and alltoall_push_entries(0, 6, [(8, 3), (7, 5)]) is None  # local bits are not the leading ones: pairwise exchanges
    runs = exchange_runs([Segment("exchange", gbit=8, lbit=5), Segment("exchange", gbit=7, lbit=4), Segment("exchange", gbit=8, lbit=3),
                          Segment("local", ops=[]), Segment("exchange", gbit=6, lbit=5)])
    assert [(k, p) for k, p in runs if k == "exchange"] == [("exchange", [(8, 5), (7, 4)]), ("exchange", [(8, 3)]), ("exchange", [(6, 5)])]


def test_fused_exchange_permutation_equivalence(monkeypatch):
    """EXPERIMENTAL path (QB_A2A_FUSE_PERM=1): the closing permutation of the segment after a run of exchanges, written
    by the all-to-all itself (one K8 per chunk into the destination rank's second buffer) -- all ranks simulated in one
    process with the product's K8 index math (tests/emul), against exchanges -> gates -> permutation in sequence."""
    sys.path[:0] = [ROOT, HERE]
    import emul
    from helpers import oracle_run, rand_state
    from qibo_b200 import circuits
    from qibo_b200.distributed import (alltoall_push_entries, choose_layout, exchange_runs, specialise,
                                       split_trailing_permutation)
    from qibo_b200.engine import split_swap_runs

    for n, g in ((13, 2), (14, 3), (12, 1)):
        W, nl = 1 << g, n - g
        plan = choose_layout(n, g, circuits.qft(n))
        runs = exchange_runs(plan.segments)
        assert [k for k, _ in runs] == ["local", "exchange", "local"]
        pairs, tail = runs[1][1], runs[2][1]
        split = split_trailing_permutation(tail, nl, len(pairs))
        assert split is not None
        gates, sub_dest = split
        k, lo = len(pairs), nl - len(pairs)
        shards = [rand_state(nl, 40 + r) for r in range(W)]
        # reference: pairwise exchanges in sequence, then the whole tail segment (gates + SWAPs) on every rank
        ref = [x.copy() for x in shards]
        for gbit, lbit in pairs:
            j = gbit - nl
            new = [x.copy() for x in ref]
            idx = np.arange(1 << nl)
            for r in range(W):
                b = (r >> j) & 1
                mine = idx[((idx >> lbit) & 1) == 1 - b]
                new[r][mine] = ref[r ^ (1 << j)][mine ^ (1 << lbit)]
            ref = new
        for r in range(W):
            local = [o for o in (specialise(p, nl, r) for p in tail) if o is not None]
            ref[r] = oracle_run(ref[r], local, nl)
        # fused: every chunk lands permuted in the destination's second buffer, then only the gates run
        second = [np.zeros(1 << nl, dtype=complex) for _ in range(W)]
        csz = 1 << lo
        for r in range(W):
            for r2, a, b_, _, _ in alltoall_push_entries(r, nl, pairs):
                second[r2][b_ : b_ + csz] = emul.permute_qubits(shards[r][a : a + csz], lo, sub_dest)
        for r in range(W):
            local = [o for o in (specialise(p, nl, r) for p in gates) if o is not None]
            got = oracle_run(second[r], local, nl) if local else second[r]
            assert np.abs(got - ref[r]).max() < 1e-12, (n, g, r)
    # the same plan through ShardedProgram on gloo: no peer memory there, so the permutation runs right after the exchange
    monkeypatch.setenv("QB_A2A_FUSE_PERM", "1")
    for world in (2, 4):
        err, nex, planned = _run(world, "qft", 9, seed=6, layout="auto")
        assert err < 1e-12 and nex == planned == world.bit_length() - 1


@pytest.mark.parametrize("block", range(4))
def test_plans_on_random_circuits_all_ranks_simulated(block):
    """Plan + specialise + exchanges on random circuits, layouts (auto, random global sets, with and without batched
    exchanges) and world sizes 2/4/8, every rank simulated in this process: local segments by the oracle, exchanges as
    NumPy half-shard swaps or as the out-of-place all-to-all -- the gathered state equals the oracle's."""
    sys.path[:0] = [ROOT, HERE]
    from helpers import oracle_run, rand_state, random_zoo
    from qibo_b200 import circuits
    from qibo_b200 import distributed as D

    for seed in range(12 * block, 12 * block + 12):
        rng = np.random.default_rng(seed)
        g = int(rng.integers(1, 4))
        W = 1 << g
        n = int(rng.integers(g + 3, 10))
        kind = seed % 4
        if kind == 0:
            ops = random_zoo(n, int(rng.integers(5, 40)), seed, max_dense=min(3, n - g))
        elif kind == 1:
            ops = circuits.qft(n)
        elif kind == 2:
            ops = circuits.random_circuit(n, int(rng.integers(5, 30)), seed)
        else:
            ops = circuits.variational(n, 2, rng.random(8 * n) * 6)
        lay = seed % 3
        if lay == 0:
            plan = D.choose_layout(n, g, ops)
        elif lay == 1:  # any set of global qubits, and another one to leave the state in
            plan = D.Plan(n, g, ops, global_qubits=tuple(rng.permutation(n)[:g].tolist()),
                          final_global_qubits=tuple(rng.permutation(n)[:g].tolist()) if seed % 2 else None)
        else:
            plan = D.Plan(n, g, ops, batch_exchanges=bool(seed % 2))
        if seed % 5 == 0:  # fewest exchanges on the way in, block layout on the way out (what the sharded measurement takes)
            plan = D.choose_layout(n, g, ops, final_global_qubits=D.block_layout(n, g))
        nl = n - g
        axes = list(plan.global_qubits) + list(plan.local_qubits)
        psi = rand_state(n, seed)
        t = psi.reshape((2,) * n).transpose(axes).reshape(W, 1 << nl)
        shards = [t[r].copy() for r in range(W)]
        nex = 0
        for kind_, payload in D.exchange_runs(plan.segments):
            if kind_ == "local":
                for r in range(W):
                    local = [o for o in (D.specialise(p, nl, r) for p in payload) if o is not None]
                    if local:
                        shards[r] = oracle_run(shards[r], local, nl)
                continue
            nex += len(payload)
            if D.alltoall_push_entries(0, nl, payload) is not None and seed % 2:
                second = [np.zeros(1 << nl, complex) for _ in range(W)]
                for r in range(W):
                    for r2, a, b_, lo_, hi_ in D.alltoall_push_entries(r, nl, payload):
                        second[r2][b_ + lo_ : b_ + hi_] = shards[r][a + lo_ : a + hi_]
                shards = second
                continue
            for gbit, lbit in payload:
                j = gbit - nl
                new = [x.copy() for x in shards]
                idx = np.arange(1 << nl)
                for r in range(W):
                    b = (r >> j) & 1
                    mine = idx[((idx >> lbit) & 1) == 1 - b]
                    new[r][mine] = shards[r ^ (1 << j)][mine ^ (1 << lbit)]
                shards = new
        final_axes = list(plan.final_global_qubits) + list(plan.final_local_qubits)
        full = np.stack(shards).reshape((2,) * n).transpose(np.argsort(final_axes)).reshape(-1)
        assert np.abs(full - oracle_run(psi, ops, n)).max() < 1e-12, seed
        assert nex == plan.nexchanges


def _verify_worker(rank, world, port, n, layout, break_it, out):
    sys.path[:0] = [ROOT, HERE]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace

        import bench
        from qibo_b200 import circuits
        from qibo_b200.distributed import ShardedProgram

        calls = {"n": 0}

        def apply(tensor, nlocal, ops):
            _oracle_apply(tensor, nlocal, ops)
            calls["n"] += 1
            if break_it and rank == 1:
                tensor[3] += 1e-3  # a wrong amplitude on one rank must fail the check on every rank, on every transport

        prog = ShardedProgram(None, n, "complex128", circuits.qft(n), apply=apply, staging_elems=8, global_qubits=layout)
        state = SimpleNamespace(tensor=torch.zeros(1 << prog.nlocal, dtype=torch.complex128))
        args = SimpleNamespace(dtype="complex128")
        try:
            res = bench.verify_sharded_qft(prog, state, args, None, dist)
        except AssertionError as e:
            res = {"failed": str(e)}
        r0, loc0 = prog.locate(0)
        ok_reset = bool(state.tensor.abs().sum().item() == (1.0 if r0 == rank else 0.0)) and bool(r0 != rank or state.tensor[loc0] == 1)
        if rank == 0:
            out.put((res, ok_reset))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,layout,break_it", [(2, "auto", False), (4, "auto", False), (4, None, False), (2, "auto", True)])
def test_bench_closed_form_check(world, layout, break_it):
    """bench.py's pre-timing check of a sharded QFT (closed form exp(2 pi i x k / 2^n) / sqrt(2^n) on every amplitude of
    every rank) on gloo with the oracle as the shard executor: passes on a correct run, fails on every rank when one rank
    holds a wrong amplitude, and leaves |0...0> behind."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_verify_worker, args=(r, world, port, 10, layout, break_it, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res, ok_reset = out.get()
    if break_it:
        assert "failed" in res
    else:
        assert res["max_rel_err"] < 1e-12 and res["amplitudes_checked_per_rank"] == 2 ** 10 // world and ok_reset
        assert res["exchange_path"] == "pipelined-dma"  # (no peer memory on gloo: the name of the configured path)


class _FakeEngine:
    """TEST INFRASTRUCTURE: stands in for qibo_b200.engine.Engine on CPU shards (oracle arithmetic) so that the plumbing of
    ShardedProgram.run -- compiled programs on first use, the uncompiled e2e path, statistics -- runs without a GPU."""

    def __init__(self):
        self.compiled, self.launched, self.applied = 0, 0, 0

    def compile(self, nqubits, dtype, ops):
        from types import SimpleNamespace

        self.compiled += 1
        return SimpleNamespace(nqubits=nqubits, dtype=np.dtype(dtype), ops=list(ops))

    def _stats(self, ops):
        from types import SimpleNamespace

        return SimpleNamespace(nsweeps=1, elapsed_ms=0.5, perm_ms=0.0, nperm=0, nops=len(ops))

    def run_program(self, prog, state, timed=False, alt=None, spans=None):
        assert alt is None  # only peer-memory shards carry a second buffer
        _oracle_apply(state.tensor, prog.nqubits, prog.ops)
        self.launched += 1
        return self._stats(prog.ops)

    def apply_program(self, state, nqubits, ops, timed=False, alt=None, spans=None):
        _oracle_apply(state.tensor, nqubits, ops)
        self.applied += 1
        return self._stats(ops)


def _engine_worker(rank, world, port, n, out):
    sys.path[:0] = [ROOT, HERE]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace

        from helpers import oracle_run, rand_state
        from qibo_b200 import circuits
        from qibo_b200.distributed import ShardedProgram

        ops = circuits.qft(n)
        eng = _FakeEngine()
        prog = ShardedProgram(eng, n, "complex128", ops, staging_elems=8, global_qubits="auto")
        psi = rand_state(n, 9)
        state = SimpleNamespace(tensor=torch.from_numpy(prog.shard_of(psi).copy()))
        s1 = prog.run(state)  # compiles every local segment once ...
        s2 = prog.run(state, timed=False)  # ... and only launches afterwards
        # (the tail of a local segment that would ride on a chunk-pipelined exchange runs as a program of its own on
        # transports without peer memory)
        nlocal_segments = sum(1 for seg in prog.segments if seg[0] == "local" and seg[1]) + sum(
            1 for seg in prog.segments if seg[0] == "exchange" and seg[4] is not None and seg[4].nlocal_ops)
        assert eng.compiled == nlocal_segments and eng.launched == 2 * nlocal_segments and eng.applied == 0
        s3 = prog.run(state, timed=False, compiled=False)  # the e2e leg: host gate matrices on every call
        assert eng.applied == nlocal_segments and eng.compiled == nlocal_segments
        ref = oracle_run(oracle_run(oracle_run(psi, ops, n), ops, n), ops, n)
        err = float(np.abs(prog.gather(state) - ref).max())
        assert s1.nsweeps == s2.nsweeps == s3.nsweeps == nlocal_segments and s1.nexchanges == prog.plan.nexchanges
        assert s1.nexchange_launches == s1.nexchanges and s1.exchange_bytes > 0  # pairwise exchanges on gloo
        if rank == 0:
            out.put(err)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_program_engine_plumbing(world):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_engine_worker, args=(r, world, port, 9, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get() < 1e-12


def test_api_layout_choice():
    sys.path[:0] = [ROOT, HERE]
    from qibo_b200 import circuits
    from qibo_b200.distributed import Plan, api_layout, block_layout, choose_layout

    assert api_layout(12, 30, False, 4) == dict(global_qubits="auto")
    assert api_layout(12, 30, True, 4) == dict(global_qubits="auto")
    assert api_layout(35, 30, False, 8) == dict(global_qubits="auto", final_global_qubits=(0, 1, 2))
    assert api_layout(35, 30, True, 8) == {}
    # what the large-register path buys for a QFT: two runs of exchanges (all-to-alls) instead of seven scattered ones
    ops = circuits.qft(35)
    through = choose_layout(35, 3, ops, final_global_qubits=block_layout(35, 3))
    assert through.final_global_qubits == (0, 1, 2) and through.nexchanges == 6
    assert Plan(35, 3, ops).nexchanges == 7


def test_locate_and_canonical_index_round_trip():
    """ShardedProgram.locate (initial layout) and canonical_index (final layout) are inverse bit shuffles -- bench.py's
    closed-form check of the sharded QFT relies on them."""
    sys.path[:0] = [ROOT, HERE]
    from types import SimpleNamespace

    from qibo_b200.distributed import ShardedProgram

    rng = np.random.default_rng(2)
    for n, g in ((7, 2), (9, 3), (6, 1), (35, 3)):
        for _ in range(4):
            gq = tuple(rng.permutation(n)[:g].tolist())
            lq = tuple(q for q in range(n) if q not in gq)
            ns = SimpleNamespace(n=n, g=g, global_qubits=gq, local_qubits=lq, final_global_qubits=gq, final_local_qubits=lq)
            for index in [0, (1 << n) - 1] + [int(x) for x in rng.integers(0, 1 << n, size=20)]:
                r, loc = ShardedProgram.locate(ns, index)
                assert 0 <= r < (1 << g) and 0 <= loc < (1 << (n - g))
                assert ShardedProgram.canonical_index(ns, r, loc) == index
    # block layout: rank = leading bits, shard index = the rest; cyclic layout: rank = trailing bits
    ns = SimpleNamespace(n=5, g=2, global_qubits=(0, 1), local_qubits=(2, 3, 4), final_global_qubits=(0, 1), final_local_qubits=(2, 3, 4))
    assert ShardedProgram.locate(ns, 0b10110) == (0b10, 0b110)
    ns = SimpleNamespace(n=5, g=2, global_qubits=(3, 4), local_qubits=(0, 1, 2), final_global_qubits=(3, 4), final_local_qubits=(0, 1, 2))
    assert ShardedProgram.locate(ns, 0b10110) == (0b10, 0b101)


def test_plan_properties():
    """Device-free planner checks in the spirit of tests/test_models_distcircuit.py:95-103: no mixing target is
    ever on a global bit inside a local segment, and the layout is canonical at the end."""
    sys.path[:0] = [ROOT, HERE]
    from helpers import ops_from_named
    from oracle import numpy_oracle as orc
    from qibo_b200.distributed import Plan

    for n in (28, 31, 35):
        ops = ops_from_named(orc.qft_ops(n))
        for g in (1, 2, 3):
            plan = Plan(n, g, ops)
            nl = n - g
            for seg in plan.segments:
                if seg.kind == "local":
                    for p in seg.ops:
                        if not p.is_diagonal:
                            rows, cols = np.nonzero(p.data)
                            mixed = 0
                            for d in np.unique(rows ^ cols):
                                mixed |= int(d)
                            k = len(p.tbits)
                            for i, b in enumerate(p.tbits):
                                if (mixed >> (k - 1 - i)) & 1:
                                    assert b < nl
                else:
                    assert seg.gbit >= nl > seg.lbit >= nl - 8  # contiguous chunks of >= 2^(nl-8) amplitudes
            assert plan.nexchanges <= 4 * g + 2, (n, g, plan.nexchanges)


def test_specialise_global_control_and_diagonal():
    sys.path[:0] = [ROOT, HERE]
    from oracle import numpy_oracle as orc
    from qibo_b200.distributed import PhysOp, specialise

    cnot = PhysOp(orc.gate_matrix("CNOT"), (5, 1), (), False)  # control on global bit 5 (nlocal = 4), target local bit 1
    assert specialise(cnot, 4, 0) is None  # rank bit 1 (bit 5 - 4) is 0 on rank 0
    op = specialise(cnot, 4, 2)
    np.testing.assert_array_equal(op.data, orc.gate_matrix("X"))
    assert op.targets == (2,) and op.controls == ()
    cu1 = PhysOp(orc.gate_matrix("CU1", 0.3), (4, 5), (), False)  # both qubits global
    assert specialise(cu1, 4, 1) is None and specialise(cu1, 4, 2) is None
    op = specialise(cu1, 4, 3)
    assert op.targets == () and abs(op.data[0, 0] - np.exp(0.3j)) < 1e-15
    rz = PhysOp(np.array([np.exp(-0.1j), np.exp(0.1j)]), (4,), (2,), True)
    op = specialise(rz, 4, 1)
    assert op.is_diagonal and op.targets == () and op.controls == (1,) and abs(op.data[0] - np.exp(0.1j)) < 1e-15


# ---- sharded measurement (qibo_b200/dist_measure.py): collectives over gloo, shard-local arithmetic by the oracle ----
class _OracleLocal:
    """TEST INFRASTRUCTURE: NumPy stand-in for dist_measure.EngineLocal (CPU shards are plain torch tensors)."""

    def probabilities(self, shard, qubits, nlocal):
        from oracle import numpy_oracle as orc

        return torch.from_numpy(np.ascontiguousarray(orc.calculate_probabilities(shard.numpy(), list(qubits), nlocal)))

    def cdf(self, probs):
        c = np.cumsum(probs.numpy().astype(np.float64))
        return torch.from_numpy(c / c[-1])

    def search(self, cdf, uniforms):
        return torch.from_numpy(np.searchsorted(cdf.numpy(), uniforms.numpy(), side="right").astype(np.int64))

    def collapse(self, shard, nlocal, qubits, outcome):
        arr = shard.numpy()
        keep = np.ones(arr.shape[0], dtype=bool)
        idx = np.arange(arr.shape[0])
        for pos, q in enumerate(qubits):
            bit = (outcome >> (len(qubits) - 1 - pos)) & 1
            keep &= ((idx >> (nlocal - 1 - q)) & 1) == bit
        arr[~keep] = 0

    def zero(self, shard):
        shard.zero_()

    def norm2(self, shard):
        return float((shard.abs() ** 2).sum())

    def scale(self, shard, nlocal, factor):
        shard.mul_(factor)

    def device(self, shard):
        return shard.device


def _measure_worker(rank, world, port, n, out):
    sys.path[:0] = [ROOT, HERE]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import rand_state
        from oracle import numpy_oracle as orc
        from qibo_b200.dist_measure import ShardMeasure

        psi = rand_state(n, 21)
        sm = ShardMeasure(n, _OracleLocal())
        nl = sm.nlocal
        shard = torch.from_numpy(psi[rank << nl : (rank + 1) << nl].copy())
        errs = []
        for qubits in ([0], [n - 1], [1, 3], [3, 0, 2], [4, 1], list(range(n))[::-1], [2, 5, 0, 1]):
            p, sharded = sm.probabilities(shard, qubits)
            assert not sharded
            errs.append(float(np.abs(p.numpy() - orc.calculate_probabilities(psi, qubits, n)).max()))
        p, sharded = sm.probabilities(shard, list(range(n)))
        assert sharded and p.numel() == 1 << nl
        full = np.abs(psi) ** 2
        errs.append(float(np.abs(p.numpy() - full[rank << nl : (rank + 1) << nl]).max()))
        # sampling from the sharded distribution == inverse CDF on the full one (same uniforms)
        u = np.random.default_rng(5).random(2000)
        s = sm.sample(p, u, sharded=True).numpy()
        c = np.cumsum(full)
        ref = np.searchsorted(c / c[-1], u, side="right")
        mismatch = int((s != ref).sum())
        # collapse over a mix of global and local qubits
        for qubits, outcome in (([0, 2], 1), ([1], 1), ([0, 1, 4], 5), ([3, 5], 2)):
            sh = shard.clone()
            sm.collapse(sh, qubits, outcome)
            parts = [torch.empty_like(sh) for _ in range(world)]
            dist.all_gather(parts, sh)
            got = torch.cat(parts).numpy()
            want = orc.collapse_statevector(psi, qubits, [outcome], n)
            errs.append(float(np.abs(got - want).max()))
        if rank == 0:
            out.put((max(errs), mismatch))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_measurement_matches_oracle(world):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_measure_worker, args=(r, world, port, 7, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    err, mismatch = out.get()
    assert err < 1e-12
    assert mismatch <= 2  # a uniform within rounding of a rank boundary of the CDF may land on the neighbouring bin


def test_pipelined_exchange_splits_the_qft(monkeypatch):
    """What the chunk pipeline does to the sharded QFT: the sweeps after the first one touch none of the leading local
    qubits, so they ride on the exchange chunk by chunk (host logic; the numbers are checked above); with the opt-in
    QB_SINK_LOW_QUBITS=5 the stages on the five lowest bits move behind the exchange."""
    from qibo_b200 import circuits
    from qibo_b200.distributed import ShardedProgram

    monkeypatch.setenv("QB_SINK_LOW_QUBITS", "5")

    prog = ShardedProgram(None, 24, "complex128", circuits.qft(24), global_qubits="auto", world_=8, rank_=5)
    kinds = [seg[0] for seg in prog.segments]
    assert kinds.count("exchange") == 1
    x = prog.segments[kinds.index("exchange")]
    assert len(x[1]) == 3 and x[4] is not None and len(x[4].ops) > 80
    assert all(min(op.targets + op.controls) >= 0 for op in x[4].ops)
    head = prog.segments[kinds.index("exchange") - 1][1]
    assert len(head) > 0
    # the stages on the five lowest state bits sink behind the exchange, in front of the stages on the arrived qubits: the
    # last segment is then ONE full sweep (with the closing permutation), the chunk sweeps lose five stages
    after = prog.segments[kinds.index("exchange") + 1][1]
    from qibo_b200.distributed import _is_diagonal_op

    mixing = [op.targets[0] for op in after if len(op.targets) == 1 and not _is_diagonal_op(op)]
    assert mixing[:5] == [16, 17, 18, 19, 20] and sorted(mixing[5:8]) == [0, 1, 2]
    assert not any(q >= 16 for op in x[4].nlocal_ops if len(op.targets) == 1 and not _is_diagonal_op(op) for q in op.targets)


@pytest.mark.parametrize("sink", ["0", "5", "3"])
def test_sharded_qft_with_and_without_sunk_stages(sink, monkeypatch):
    """The gates that sink behind an exchange (QB_SINK_LOW_QUBITS low state bits; 0 = none): the gathered state of a
    sharded QFT is the oracle's either way (gloo, world size 4, real exchanges, NumPy shard executor)."""
    monkeypatch.setenv("QB_SINK_LOW_QUBITS", sink)
    err, nex, planned = _run(4, "qft", 10, seed=6, layout="auto")
    assert err < 1e-12 and nex == planned == 2


@pytest.mark.parametrize("case,n,world", [("qft", 17, 8), ("qft", 15, 2), ("variational", 16, 4), ("zoo", 16, 4)])
def test_pipelined_tail_is_chunk_separable(case, n, world):
    """The ops that ride on a chunk-pipelined exchange, applied chunk by chunk in the chunk's own qubit numbering, equal
    the same ops applied to the whole shard (oracle arithmetic, every rank)."""
    sys.path[:0] = [ROOT, HERE]
    from helpers import ops_from_named, oracle_run, rand_state, random_zoo
    from oracle import numpy_oracle as orc
    from qibo_b200.distributed import ShardedProgram

    if case == "qft":
        ops = ops_from_named(orc.qft_ops(n))
    elif case == "variational":
        ops = ops_from_named(orc.variational_ops(n, 3, np.random.default_rng(1).random(6 * n) * 6))
    else:
        ops = random_zoo(n, 60, 5, max_dense=3)
    seen = 0
    for r in range(world):
        prog = ShardedProgram(None, n, "complex128", ops, global_qubits="auto", world_=world, rank_=r)
        for seg in prog.segments:
            if seg[0] != "exchange" or seg[4] is None or not seg[4].ops:
                continue
            k, nl = len(seg[1]), prog.nlocal
            psi = rand_state(nl, 3 + r)
            whole = oracle_run(psi, seg[4].nlocal_ops, nl)
            csz = 1 << (nl - k)
            parts = [oracle_run(psi[c * csz : (c + 1) * csz].copy(), seg[4].ops, nl - k) for c in range(1 << k)]
            assert np.abs(np.concatenate(parts) - whole).max() < 1e-13
            seen += 1
    assert seen > 0 or case != "qft"  # (layered circuits touch every qubit in every sweep: nothing to pipeline)
