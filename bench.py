#!/usr/bin/env python
"""bench.py -- QFT gates/s and HBM GB/s per sweep on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--nqubits n] [--dtype complex128] [--impl reference]

A "step" is one execution of the whole QFT(n) gate queue (n(n+1)/2 + n/2 gates) on a state resident in HBM.
N = 1: QFT(32) complex128 (the north star's single-GPU target; 64 GiB state).  N > 1 (torchrun, one rank per
GPU): QFT(32 + log2 N) sharded over log2 N global qubits -- 64 GiB per GPU, i.e. weak scaling, QFT(35) at N = 8.
One JSON line on stdout (rank 0).  `--impl reference` times the reference's CPU implementation instead: bounded samples of
the same workload per step plus ONE whole QFT(26) through `Circuit()` on the NumpyBackend -- the same-config anchor that
the GPU arm also runs (`anchor` in both lines).
"""

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "qft_gates_per_second"
ANCHOR_N = 26
CPU_THREADS_CAP = 16


def n_gates(n):
    return n * (n + 1) // 2 + n // 2


def bench_config(n, dtype, world):
    """Identical in both arms (the driver compares the reference line's config with ours)."""
    g = int(round(math.log2(world)))
    itemsize = 16 if dtype == "complex128" else 8
    return {
        "workload": f"QFT({n}) {dtype}, {n_gates(n)} gates, {2 ** (n - g) * itemsize / 2 ** 30:.0f} GiB of state per GPU on {world} GPU(s)",
        "nqubits": n, "gates": n_gates(n), "gpus": world, "global_qubits": g,
        "l2": "state (>= 16 GiB) is far larger than the 126 MB L2",
    }


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(n, dtype):
    """dram bytes per sweep launch from the committed ncu --set full capture, if one exists for this size."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return d.get(f"{dtype}:{n}")
    except Exception:
        return None


def import_qibo():
    """The unmodified reference package from baseline/_ref (travels to the GPU box); None when it is absent."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "qibo")):
        return None
    if ref_dir not in sys.path:
        sys.path.append(ref_dir)
    os.environ.setdefault("QIBO_LOG_LEVEL", "3")
    try:
        import qibo

        return qibo
    except Exception:
        return None


def pin_cpu_threads():
    """The same BLAS thread count under plain python and under torchrun (which exports OMP_NUM_THREADS=1)."""
    cores = max(1, min(os.cpu_count() or 1, CPU_THREADS_CAP))
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=cores)
    except Exception:
        pass
    return cores


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arms
class CpuQft:
    """The reference's CPU path for QFT(n_cpu): the real NumpyBackend (kind "reference") when baseline/_ref is importable,
    else the oracle's restatement of it (kind "port")."""

    def __init__(self, n_cpu, dtype):
        self.n, self.dtype = n_cpu, dtype
        self.cores = pin_cpu_threads()
        self.kind = "port"
        self.qibo = import_qibo()
        if self.qibo is not None:
            from qibo import gates
            from qibo.backends import NumpyBackend

            self.be = NumpyBackend()
            self.be.set_dtype(dtype)
            self.glist = []
            for i1 in range(n_cpu):
                self.glist.append(gates.H(i1))
                for i2 in range(i1 + 1, n_cpu):
                    self.glist.append(gates.CU1(i2, i1, math.pi / 2 ** (i2 - i1)))
            self.total = len(self.glist)
            self.kind = "reference"
        else:
            from oracle import numpy_oracle as orc

            self.orc = orc
            self.named = orc.qft_ops(n_cpu)
            self.mats = [orc.gate_matrix(nm, *pr, dtype=dtype) for nm, _, pr in self.named]
            self.total = len(self.named)

    def apply(self, state, k):
        if self.kind == "reference":
            return self.be.apply_gate(self.glist[k], state, self.n)
        return self.orc.apply_gate(state, self.mats[k], self.named[k][1], self.n)

    def sample(self, budget_s, state=None):
        """As many leading gates as fit the time budget -> (gates done, seconds)."""
        if state is None:
            state = np.zeros(2**self.n, dtype=self.dtype)
            state[0] = 1
        t0 = time.perf_counter()
        done = 0
        while done < self.total and (done == 0 or time.perf_counter() - t0 < budget_s):
            state = self.apply(state, done)
            done += 1
        return done, time.perf_counter() - t0

    def whole_circuit(self):
        """ONE complete QFT(n) through the reference's own public API -> seconds (None without the reference package)."""
        if self.kind != "reference":
            return None
        from qibo.models import QFT

        c = QFT(self.n)
        t0 = time.perf_counter()
        res = self.be.execute_circuit(c)
        dt = time.perf_counter() - t0
        amp = complex(res.state()[0])
        assert abs(amp - 2.0 ** (-self.n / 2)) < 1e-9
        return dt


def cpu_reference_sample(n_target, dtype, budget_s=15.0):
    """`cpu_baseline` of the GPU arm: a bounded sample on the host cores, scaled to n_target by the 2^n cost of a gate."""
    cpu = CpuQft(ANCHOR_N, dtype)
    cpu.sample(0.0)  # warm-up (page faults, BLAS init): one gate
    done, dt = cpu.sample(budget_s)
    rate = done / dt
    scale = 2.0 ** (n_target - ANCHOR_N)
    return {
        "value": rate / scale, "unit": "gates/s", "cores": cpu.cores, "kind": cpu.kind,
        "sample": f"first {done} gates of QFT({ANCHOR_N}) {dtype} in {dt:.1f} s on the host ({rate:.2f} gates/s at n={ANCHOR_N}), "
        f"scaled by 2^({n_target}-{ANCHOR_N}) to n={n_target}; NumPy transposes are single-threaded, BLAS pinned to {cpu.cores} threads",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    g = int(round(math.log2(args.gpus)))
    n = args.nqubits + g
    cpu = CpuQft(ANCHOR_N, args.dtype)
    anchor = None
    if not args.no_anchor:
        cpu_s = cpu.whole_circuit()
        if cpu_s is not None:
            anchor = {"n": ANCHOR_N, "dtype": args.dtype, "gates": n_gates(ANCHOR_N), "cpu_s": cpu_s, "cpu_gates_per_s": n_gates(ANCHOR_N) / cpu_s,
                      "api": "qibo.models.QFT(26) -> NumpyBackend.execute_circuit, whole circuit, measured (no extrapolation)", "cores": cpu.cores}
    # K timed steps (+ W warm-up steps), each a bounded sample: the leading gates of QFT(26) for step_budget seconds
    total_steps = max(1, args.warmup + args.steps)
    budget = min(args.cpu_budget, max(1.0, 100.0 / total_steps))
    rates, walls = [], []
    for i in range(total_steps):
        done, dt = cpu.sample(budget)
        if i >= args.warmup or total_steps == 1:
            rates.append(done / dt)
            walls.append(dt)
    scale = 2.0 ** (n - ANCHOR_N)
    # same unit as the GPU arm: gates x shards of 2^(n - log2 N) amplitudes (one host applies each gate to all N shards)
    v = args.gpus * float(np.mean(rates)) / scale
    base = {"value": v, "unit": "gates/s", "cores": cpu.cores, "kind": cpu.kind,
            "sample": f"{len(rates)} steps, each the leading gates of QFT({ANCHOR_N}) {args.dtype} for {budget:.1f} s on the host "
                      f"({float(np.mean(rates)):.2f} gates/s at n={ANCHOR_N}), scaled by 2^({n}-{ANCHOR_N}) to n={n}; BLAS pinned to {cpu.cores} threads"}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "gates/s", "n_gpus": args.gpus, "steps": len(rates),
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128" if args.dtype == "complex128" else "c64", "data": "synthetic",
        "config": bench_config(n, args.dtype, args.gpus),
        "cpu_baseline": base, "anchor": anchor,
        "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "step_definition": "one bounded sample of the workload on the host (ms_per_step is its wall time); value extrapolates the "
                           "sample's gate rate to the configured size; `anchor` is a complete, un-extrapolated run both arms share",
    }
    print(json.dumps(line), flush=True)


def verify_sharded_qft(runner, state, args, eng, dist):
    """Outside the timed region: QFT|x> on the shards against the closed form exp(2 pi i x k / 2^n) / sqrt(2^n) on EVERY
    amplitude of every rank (qibo_b200/checks.py: evaluated on the device, int64 phase arithmetic).  A misplaced chunk or a
    wrong rank-specialised phase shows; the norm alone would not.  If a transport fails the check, fall back to the next
    one (chunk-pipelined DMA -> all-to-all kernel out of place -> in place -> pairwise) and check again."""
    import torch

    from qibo_b200.checks import generic_basis_state, qft_basis_state_error

    n = runner.n
    x = generic_basis_state(n)
    rk, loc = runner.locate(x)
    tol = 1e-9 if args.dtype == "complex128" else 2e-3
    out = {}
    # canonical index of shard-local index `loc` in the final layout, vectorised
    fg, fl, g = runner.final_global_qubits, runner.final_local_qubits, runner.g

    def canonical(loc_t):
        idx = torch.zeros_like(loc_t)
        for j, q in enumerate(fg):
            idx |= ((runner.rank >> (g - 1 - j)) & 1) << (n - 1 - q)
        nl = len(fl)
        for k, q in enumerate(fl):
            idx |= ((loc_t >> (nl - 1 - k)) & 1) << (n - 1 - q)
        return idx

    def path():
        if not runner.alltoall:
            return "pairwise"
        if runner.alltoall_push and runner.pipeline:
            return "pipelined-dma"
        fused = "+perm" if (runner.fuse_perm and runner.alltoall_push) else ""  # experimental QB_A2A_FUSE_PERM=1
        return ("alltoall-push" if runner.alltoall_push else "alltoall-swap") + fused

    for _ in range(4):
        attempt = path()
        st = state
        st.tensor.zero_()
        if rk == runner.rank:
            st.tensor[loc] = 1
        runner.run(st, timed=False)
        err = qft_basis_state_error(st.tensor, n, x, canonical_index=canonical)
        t = torch.tensor([err], device=st.tensor.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = {"basis_state": x, "amplitudes_checked_per_rank": int(st.tensor.numel()), "max_rel_err": float(t.item()), "exchange_path": attempt}
        if out["max_rel_err"] < tol or attempt == "pairwise":
            break
        # every rank sees the same all-reduced error, so they switch together
        if runner.fuse_perm:
            raise AssertionError(f"QB_A2A_FUSE_PERM=1 (experimental) fails the closed-form check: {out}")
        if runner.pipeline and runner.alltoall_push:
            runner.pipeline = False
        elif runner.alltoall_push:
            runner.alltoall_push = False
        else:
            runner.alltoall = False
    if not out["max_rel_err"] < tol:
        raise AssertionError(f"sharded QFT does not match the closed form: {out}")
    state.tensor.zero_()
    if runner.rank == runner.locate(0)[0]:
        state.tensor[0] = 1
    return out


def verify_single_qft(eng, compiled, n, dtype):
    """N = 1: the same closed-form check on every amplitude, for the compiled program that is about to be timed."""
    import torch

    from qibo_b200.checks import generic_basis_state, qft_basis_state_error

    x = generic_basis_state(n)
    st = eng.basis_state(n, dtype, x)
    eng.run_program(compiled, st)
    err = qft_basis_state_error(st.tensor, n, x)
    tol = 1e-9 if dtype == "complex128" else 2e-3
    out = {"basis_state": x, "amplitudes_checked": int(st.tensor.numel()), "max_rel_err": err, "path": "qb_program_run (compiled program)"}
    if not err < tol:
        raise AssertionError(f"QFT({n}) does not match the closed form: {out}")
    del st
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------- plugin-level legs (N = 1)
def plugin_backend(dtype):
    qibo = import_qibo()
    if qibo is None:
        return None, None
    qibo.set_backend("qibo_b200")
    be = qibo.get_backend()
    be.set_dtype(dtype)
    return qibo, be


def plugin_qft_step(n, nmeasured, nshots):
    """What a Qibo user writes: host Circuit objects in, host Counter out (models/circuit.py:1071-1110)."""
    from qibo import gates
    from qibo.models import QFT

    c = QFT(n)
    c.add(gates.M(*range(nmeasured)))
    res = c(nshots=nshots)
    return res.frequencies(binary=False)


def config3_leg(qibo, be, torch, small=False):
    """BASELINE config 3: 32-qubit RY/CZ ansatz, depth 20, complex64, unfused and through circuit.fuse(max_qubits=2..5)
    (examples/benchmarks/circuits.py:7-22; tests/test_models_circuit_fuse.py:109-140), through the plugin."""
    from qibo import Circuit, gates

    n, layers = (24, 4) if small else (32, 20)
    be.set_dtype("complex64")
    theta = iter(2 * np.pi * np.random.default_rng(7).random(2 * layers * n))
    c = Circuit(n)
    for _ in range(layers):
        c.add(gates.RY(i, theta=float(next(theta))) for i in range(n))
        c.add(gates.CZ(i, i + 1) for i in range(0, n - 1, 2))
        c.add(gates.RY(i, theta=float(next(theta))) for i in range(n))
        c.add(gates.CZ(i, i + 1) for i in range(1, n - 2, 2))
        c.add(gates.CZ(0, n - 1))
    out = {"nqubits": n, "layers": layers, "dtype": "complex64", "gates": len(c.queue), "api": "qibo Circuit()() through B200Backend.execute_circuit"}

    def run(circ):
        circ()  # warm-up: planner caches, allocator
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = circ()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = be.engine_gpu.last_stats
        nrm = be.engine_gpu.norm2(res.state())
        del res
        return dt, st.nsweeps, nrm

    dt, nsweeps, nrm = run(c)
    out["unfused"] = {"seconds": dt, "gates_per_s": len(c.queue) / dt, "sweeps": nsweeps, "GBps_per_sweep": nsweeps * 2 * 8 * 2.0**n / dt / 1e9, "norm2": nrm}
    # VQE-style loop on the same Circuit object: new angles every step (Circuit.set_parameters, models/circuit.py:788-857),
    # the backend patches its compiled program in place (qb_program_set_params) instead of planning again
    rng = np.random.default_rng(3)
    nparams = len(c.get_parameters("flatlist"))
    set_ms, patch_ms, exec_s = [], [], []
    for _ in range(4):
        theta = 2 * np.pi * rng.random(nparams)
        t0 = time.perf_counter()
        c.set_parameters(theta)
        set_ms.append(1e3 * (time.perf_counter() - t0))
        t0 = time.perf_counter()
        be._compiled_circuit(c, n, False, np.dtype("complex64"))  # the parameter patch alone (execute_circuit would do it)
        patch_ms.append(1e3 * (time.perf_counter() - t0))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = c()
        torch.cuda.synchronize()
        exec_s.append(time.perf_counter() - t0)
        del res
    out["vqe_loop"] = {"parameters": nparams, "steps": len(exec_s), "qibo_set_parameters_ms": float(np.median(set_ms)),
                       "backend_parameter_patch_ms": float(np.median(patch_ms)), "execute_seconds": float(np.median(exec_s)),
                       "note": "qibo_set_parameters_ms is the reference's own Python loop over the gates; the patch is qb_program_set_params "
                               "(angles only, schedule kept) plus the re-upload of the sweep programs"}
    for k in (2, 3, 4, 5):
        t0 = time.perf_counter()
        fc = c.fuse(max_qubits=k)
        t_fuse = time.perf_counter() - t0
        dt, nsweeps, nrm = run(fc)
        out[f"fuse{k}"] = {"seconds": dt, "gates_per_s": len(c.queue) / dt, "fused_gates": len(fc.queue), "host_fuse_seconds": t_fuse,
                           "sweeps": nsweeps, "GBps_per_sweep": nsweeps * 2 * 8 * 2.0**n / dt / 1e9, "norm2": nrm}
    be.set_dtype("complex128")
    return out


def config4_leg(qibo, be, torch, small=False):
    """BASELINE config 4: 30-qubit random circuit (tests/test_models_circuit_fuse.py:124-140, seed 11), 10^6 shots and the
    marginals of SURVEY 8d, through the plugin; plus the sampling contract at scale: how many of the 10^6 samples differ
    between the parallel-scan CDF and the numpy-exact sequential one on the same probabilities and uniforms."""
    from qibo import Circuit, gates

    from qibo_b200 import _lib

    n, ngates, nshots = (24, 60, 10**5) if small else (30, 300, 10**6)
    be.set_dtype("complex128")
    np.random.seed(11)
    one, two = [gates.RX, gates.RY, gates.RZ], [gates.CNOT, gates.CZ, gates.SWAP]
    thetas = np.pi * np.random.random((ngates,))
    c = Circuit(n)
    for i in range(ngates):
        g1 = one[int(np.random.randint(0, 3))]
        c.add(g1(int(np.random.randint(0, n)), thetas[i]))
        g2 = two[int(np.random.randint(0, 3))]
        q0, q1 = np.random.randint(0, n, (2,))
        while q0 == q1:
            q0, q1 = np.random.randint(0, n, (2,))
        c.add(g2(int(q0), int(q1)))
    c.add(gates.M(*range(n)))
    c(nshots=10)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = c(nshots=nshots)
    torch.cuda.synchronize()
    t_circ = time.perf_counter() - t0
    eng = be.engine_gpu
    out = {"nqubits": n, "gates": len(c.queue) - 1, "dtype": "complex128", "circuit_seconds": t_circ, "gates_per_s": (len(c.queue) - 1) / t_circ,
           "sweeps": eng.last_stats.nsweeps, "api": "qibo Circuit()(nshots) through B200Backend.execute_circuit"}
    be.set_seed(1234)
    t0 = time.perf_counter()
    samples = res.samples(binary=False)
    t_s = time.perf_counter() - t0
    out["nshots"] = nshots
    out["sampling_seconds"] = t_s  # all-qubit probabilities (K3) + CDF (K4 scan) + search + D2H of the shots
    out["shots_per_s"] = nshots / t_s
    state = res.state()
    for q in ([0], [0, 5, 7], list(range(10)), [1, 5, 2, 0]):
        be.calculate_probabilities(state, q, n)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m = be.calculate_probabilities(state, q, n)
        torch.cuda.synchronize()
        out[f"marginal_{len(q)}q_seconds"] = time.perf_counter() - t0
        out[f"marginal_{len(q)}q_sum"] = float(m.numpy().sum())
    # the sampling contract at scale (SURVEY 8a hazard 1.iii)
    probs = eng.probabilities(state, list(range(n)), n)
    np.random.seed(1234)
    u = np.random.random_sample(nshots)
    par = eng.sample(probs, u, mode=_lib.QB_SCAN_PARALLEL)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    exact = eng.sample(probs, u, mode=_lib.QB_SCAN_EXACT)
    t_exact = time.perf_counter() - t0
    out["sampling_contract"] = {
        "bins": int(probs.size), "shots": nshots, "samples_differing_parallel_vs_sequential_cdf": int((par != exact).sum()),
        "exact_scan_seconds": t_exact, "default_above_2^22_bins": "parallel scan (QB_EXACT_SCAN_MAX_BINS / B200Backend.exact_sampling opt in to the sequential one)",
        "samples_equal_default_path": bool(np.array_equal(par, samples)),
    }
    return out


# ----------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch

    from qibo_b200 import circuits
    from qibo_b200.engine import Engine, resolve_spans

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = int(round(math.log2(world)))
    n = args.nqubits + g
    nlocal = n - g
    eng = Engine(local_rank)
    ops = circuits.qft(n)
    itemsize = 16 if args.dtype == "complex128" else 8

    if world > 1:
        from qibo_b200 import distributed

        runner = distributed.ShardedProgram(eng, n, args.dtype, ops, global_qubits=args.layout if args.layout == "auto" else None)
        state = runner.basis_state() if args.exchange == "nccl" else runner.peer_shard(0)
        step = lambda: runner.run(state)  # noqa: E731
        barrier = dist.barrier
        verify = verify_sharded_qft(runner, state, args, eng, dist)
    else:
        # the gate program is compiled once and resident on the device, like the state ("inputs already resident in HBM
        # when the timed region starts"); the e2e leg below sends host Circuit objects through the plugin on every step
        compiled = eng.compile(n, args.dtype, ops)
        verify = verify_single_qft(eng, compiled, n, args.dtype)
        state = eng.basis_state(n, args.dtype)
        alt = eng.empty((1 << n,), args.dtype)  # the second buffer the permuting last sweep writes to: resident like the state

        class _S:  # one step's counters + event spans (nothing waits for the GPU inside the step loop)
            def __init__(self):
                self.spans = []

        def step():
            s = _S()
            st = eng.run_program(compiled, state, alt=alt, spans=s.spans)
            s.nsweeps, s.nstage_sweeps, s.nperm = st.nsweeps, st.nstage_sweeps, st.nperm
            return s

        barrier = lambda: None  # noqa: E731

    for _ in range(args.warmup):
        step()
    barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    all_stats = [step() for _ in range(args.steps)]
    ev1.record()
    barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    spans = resolve_spans([sp for s in all_stats for sp in s.spans])
    sweep_ms, sweep_launches = spans.get("sweep", (0.0, 0))
    perm_ms, nperm = spans.get("perm", (0.0, 0))
    xch_ms, _ = spans.get("exchange", (0.0, 0))
    nsweeps = sum(s.nsweeps for s in all_stats)
    nstage = sum(getattr(s, "nstage_sweeps", 0) for s in all_stats)
    nchunk = sum(getattr(s, "nchunk_sweeps", 0) for s in all_stats)
    xch_bytes = sum(getattr(s, "exchange_bytes", 0) for s in all_stats)
    nxch = sum(getattr(s, "nexchanges", 0) for s in all_stats)
    nxl = sum(getattr(s, "nexchange_launches", 0) for s in all_stats)
    # units: one gate applied to one 2^nlocal-amplitude shard; every rank applies every gate of the circuit to its
    # shard, so N ranks process N * gates units per step (tier rule: "the units all ranks processed / that time")
    circuit_gates_per_s = n_gates(n) * args.steps / (ms / 1e3)
    value = world * circuit_gates_per_s

    # roofline of the dominant kernel (sweep_kernel): algorithmic bytes per launch = 2 * B * 2^nlocal, launch time from the
    # CUDA events recorded around the full-shard sweep launches inside the timed region (the sweeps on single chunks of a
    # pipelined exchange overlap with the DMA copies and are reported with the exchange instead)
    bytes_per_sweep = 2.0 * itemsize * 2.0**nlocal
    full_launches = max(sweep_launches, 1)
    avg_sweep_ms = sweep_ms / full_launches
    achieved = bytes_per_sweep / (avg_sweep_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": profiled_traffic(nlocal, args.dtype), "kernel": "sweep_kernel", "peak_source": peak_src,
        "bytes_per_launch": bytes_per_sweep, "avg_launch_ms": avg_sweep_ms, "launches": sweep_launches,
        "frac_of_nominal_8TBs": achieved / 8000.0,
    }
    if nperm:
        roofline["k8_permute"] = {"launches": nperm, "avg_launch_ms": perm_ms / nperm, "GBps": bytes_per_sweep / (perm_ms / nperm * 1e-3) / 1e9,
                                  "frac": bytes_per_sweep / (perm_ms / nperm * 1e-3) / 1e9 / peak}
        both = (sweep_ms + perm_ms) / (sweep_launches + nperm)
        roofline["all_launches"] = {"count": sweep_launches + nperm, "avg_ms": both, "GBps": bytes_per_sweep / (both * 1e-3) / 1e9,
                                    "frac": bytes_per_sweep / (both * 1e-3) / 1e9 / peak}

    e2e, e2e_engine, anchor, configs = None, None, None, None
    if world == 1:
        del state, alt
        torch.cuda.empty_cache()
        mq = list(range(min(10, n)))
        h2d_prog = sum(op.data.nbytes for op in ops) + len(ops) * 176

        # ---- e2e_engine: the C-ABI-level call with host inputs (kept from round 1 for continuity)
        def engine_step():
            st = eng.basis_state(n, args.dtype)
            eng.apply_program(st, n, ops)
            return eng.probabilities(st, mq, n).numpy()

        engine_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            probs = engine_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert abs(float(probs.sum()) - 1.0) < 1e-6
        e2e_engine = {"value": n_gates(n) * args.steps / dt, "unit": "gates/s", "h2d_bytes_per_step": int(h2d_prog),
                      "d2h_bytes_per_step": int(8 * 2 ** len(mq) * (1 if args.dtype == "complex128" else 0.5)),
                      "api": "Engine.basis_state + apply_program(host gate matrices) + probabilities(10 qubits) -> host"}
        torch.cuda.empty_cache()

        # ---- e2e: through the reference-facing plugin, as a Qibo user calls it
        qibo, be = plugin_backend(args.dtype)
        if be is not None:
            nshots = 1000
            plugin_qft_step(n, len(mq), nshots)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                freq = plugin_qft_step(n, len(mq), nshots)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert sum(freq.values()) == nshots
            e2e = {"value": n_gates(n) * args.steps / dt, "unit": "gates/s", "ms_per_step": 1e3 * dt / args.steps,
                   "h2d_bytes_per_step": int(h2d_prog + 8 * nshots), "d2h_bytes_per_step": int(8 * nshots + 8),
                   "api": f"qibo.set_backend('qibo_b200'); c = QFT({n}); c.add(M(*range({len(mq)}))); c(nshots={nshots}).frequencies() "
                          "-- host Circuit objects in (gate matrices built per step), host Counter out"}
            torch.cuda.empty_cache()
            # ---- same-config anchor: the whole QFT(26) through the same call; the reference arm runs the same circuit whole
            na = ANCHOR_N
            plugin_qft_step(na, 10, nshots)
            torch.cuda.synchronize()
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                plugin_qft_step(na, 10, nshots)
            torch.cuda.synchronize()
            gpu_ms = 1e3 * (time.perf_counter() - t0) / reps
            anchor = {"n": na, "dtype": args.dtype, "gates": n_gates(na), "gpu_ms": gpu_ms, "gpu_gates_per_s": n_gates(na) / (gpu_ms / 1e3),
                      "api": f"QFT({na}) + M(10 qubits), nshots={nshots}, through the plugin (wall clock, host objects in and out)",
                      "cpu_s": None, "ratio": None}
            if args.anchor_cpu:
                cpu_s = CpuQft(na, args.dtype).whole_circuit()
                if cpu_s is not None:
                    anchor["cpu_s"], anchor["ratio"] = cpu_s, cpu_s / (gpu_ms / 1e3)
            else:
                anchor["note"] = "cpu_s: the reference arm's `anchor.cpu_s` (bench.py --impl reference) on the same box; --anchor-cpu measures it here"
            if not args.no_configs:
                configs = {}
                try:
                    configs["c3"] = config3_leg(qibo, be, torch, small=args.small_configs)
                    torch.cuda.empty_cache()
                    configs["c4"] = config4_leg(qibo, be, torch, small=args.small_configs)
                except Exception as exc:  # the headline stands; say what went wrong
                    configs["error"] = repr(exc)
                torch.cuda.empty_cache()
        else:
            e2e = dict(e2e_engine)
            e2e["note"] = "reference package not importable (baseline/_ref missing): Engine-level call instead of the plugin"
    else:
        # N > 1: fresh zero shards, host gate program in (specialised per rank), global norm out (all-reduce + D2H)
        h2d = sum(op.data.nbytes for seg in runner.segments if seg[0] == "local" for op in seg[1]) + 176 * sum(
            len(seg[1]) for seg in runner.segments if seg[0] == "local")

        def e2e_step():
            if args.exchange == "nccl":
                st = runner.basis_state()
            else:
                st = state
                st.tensor.zero_()
                if rank == 0:
                    st.tensor[0] = 1
            runner.run(st, timed=False, compiled=False)  # host gate matrices in (planning + H2D) on every step
            arr = st.array if hasattr(st, "array") else st
            nrm = torch.tensor([eng.norm2(arr)], device="cuda", dtype=torch.float64)
            dist.all_reduce(nrm)
            return float(nrm.item())

        e2e_step()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            nrm = e2e_step()
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        assert abs(nrm - 1.0) < 1e-9, nrm
        e2e = {"value": world * n_gates(n) * args.steps / dt, "unit": "gates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8,
               "api": "zero shards + ShardedProgram.run(host gate matrices, per-rank specialisation) + global norm (all-reduce) -> host"}

    cpu = cpu_reference_sample(n, args.dtype, budget_s=args.cpu_budget) if (rank == 0 and world == 1 and not args.no_cpu) else None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128" if args.dtype == "complex128" else "c64", "data": "synthetic",
            "config": bench_config(n, args.dtype, world),
            "details": {
                "resident": "initial state and compiled gate program resident in HBM when the timed region starts",
                "sweeps_per_step": nsweeps // args.steps,
                **({"stage_only_sweeps_per_step": nstage // args.steps} if world == 1 else {"chunk_sweeps_per_step": nchunk // args.steps}),
                "parallelism": f"global-qubit sharding x{world}" if world > 1 else "single GPU",
                **({"layout": f"global qubits {list(runner.global_qubits)} (--layout {args.layout}); rank = their bits, shard index = "
                    "the other qubits in ascending order"} if world > 1 else {}),
                "host_sync_in_timed_loop": False,
            },
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(nsweeps + nxl), "whole_circuit_wall_s": ms_per_step / 1e3,
            "circuit_gates_per_s": circuit_gates_per_s, "verify": verify,
            "value_definition": "gates applied x shards (N ranks each apply every gate to their 2^nlocal-amplitude shard) per second; equals circuit gates/s at N=1",
        }
        if e2e_engine is not None:
            line["e2e_engine"] = e2e_engine
        if anchor is not None:
            line["anchor"] = anchor
        if configs is not None:
            line["configs"] = configs
        if world > 1:
            exposed = max(ms_per_step - (sweep_ms + perm_ms) / args.steps, 0.0)
            line["exchange"] = {
                "count_per_step": nxch // args.steps, "launches_per_step": nxl // args.steps,
                "span_ms_per_step_rank0": xch_ms / args.steps,
                "exposed_ms_per_step_rank0": exposed,
                "exposed_definition": "ms_per_step minus the CUDA-event time of the full-shard sweep and permutation launches: what the "
                                      "exchange (with the chunk sweeps it overlaps) adds to the step",
                "GBps_per_direction_rank0": (xch_bytes / 2) / max(xch_ms, 1e-9) / 1e6,
                "bytes_per_step_rank0": xch_bytes // args.steps,
                "transport": verify.get("exchange_path") if args.exchange == "p2p" else "NCCL send/recv over NVLink, half-shard pairwise",
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--nqubits", type=int, default=32, help="qubits per GPU shard + log2(gpus) global qubits")
    ap.add_argument("--dtype", default="complex128", choices=["complex128", "complex64"])
    ap.add_argument("--impl", default="qibo_b200", choices=["qibo_b200", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE config 3 / 4 legs (N = 1)")
    ap.add_argument("--small-configs", action="store_true", help="config 3 / 4 legs at reduced size (smoke runs)")
    ap.add_argument("--no-anchor", action="store_true", help="reference arm: skip the whole QFT(26) run (about 3 minutes)")
    ap.add_argument("--anchor-cpu", action="store_true", help="GPU arm: also run the whole QFT(26) on the NumpyBackend here (about 3 minutes)")
    ap.add_argument("--layout", default="auto", choices=["auto", "block"],
                    help="N > 1: which qubits are global -- 'auto' picks the layout with the fewest exchanges (trailing qubits for a QFT, "
                         "as the reference's _DistributedQFT), 'block' the leading ones")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="global<->local exchange: NVLink peer memory (DMA / kernels) or NCCL send/recv")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
