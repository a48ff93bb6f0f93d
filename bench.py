#!/usr/bin/env python
"""bench.py -- QFT gates/s and HBM GB/s per sweep on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--nqubits n] [--dtype complex128] [--impl reference]

A "step" is one execution of the whole QFT(n) gate queue (n(n+1)/2 + n/2 gates) on a state resident in HBM.
N = 1: QFT(32) complex128 (the north star's single-GPU target; 64 GiB state).  N > 1 (torchrun, one rank per
GPU): QFT(32 + log2 N) sharded over log2 N global qubits -- 64 GiB per GPU, i.e. weak scaling, QFT(35) at N = 8.
One JSON line on stdout (rank 0).  `--impl reference` times the reference's CPU implementation instead.
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


def n_gates(n):
    return n * (n + 1) // 2 + n // 2


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
def cpu_reference_sample(n_target, dtype, budget_s=15.0):
    """The reference's CPU path on a bounded sample: as many leading gates of QFT(n_cpu) as fit the time budget,
    scaled to n_target by the 2^n cost of a sweep.  Uses the real NumpyBackend when the reference package is
    importable (baseline/_ref), else the oracle's restatement of it (kind "port")."""
    n_cpu = 26
    threads = 1
    try:
        from threadpoolctl import threadpool_info

        threads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        pass
    kind = "port"
    apply = None
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_dir, "qibo")):
        try:
            sys.path.append(ref_dir)
            os.environ.setdefault("QIBO_LOG_LEVEL", "3")
            from qibo import gates
            from qibo.backends import NumpyBackend

            be = NumpyBackend()
            be.set_dtype(dtype)
            glist = []
            for i1 in range(n_cpu):
                glist.append(gates.H(i1))
                for i2 in range(i1 + 1, n_cpu):
                    glist.append(gates.CU1(i2, i1, math.pi / 2 ** (i2 - i1)))
            apply = lambda st, k: be.apply_gate(glist[k], st, n_cpu)  # noqa: E731
            total = len(glist)
            kind = "reference"
        except Exception:
            apply = None
    if apply is None:
        from oracle import numpy_oracle as orc

        named = orc.qft_ops(n_cpu)
        mats = [orc.gate_matrix(nm, *pr, dtype=dtype) for nm, _, pr in named]
        apply = lambda st, k: orc.apply_gate(st, mats[k], named[k][1], n_cpu)  # noqa: E731
        total = len(named)
    state = np.zeros(2**n_cpu, dtype=dtype)
    state[0] = 1
    state = apply(state, 0)  # warm-up (page faults, BLAS init)
    t0 = time.perf_counter()
    done = 0
    while done < total - 1 and time.perf_counter() - t0 < budget_s:
        state = apply(state, 1 + done)
        done += 1
    dt = time.perf_counter() - t0
    rate_cpu_n = done / dt
    scale = 2.0 ** (n_target - n_cpu)
    return {
        "value": rate_cpu_n / scale,
        "unit": "gates/s",
        "cores": threads,
        "kind": kind,
        "sample": f"first {done} gates of QFT({n_cpu}) {dtype} in {dt:.1f} s on the host ({rate_cpu_n:.2f} gates/s at n={n_cpu}), "
        f"scaled by 2^({n_target}-{n_cpu}) to n={n_target}; NumPy transposes are single-threaded, BLAS uses {threads} threads",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    g = int(round(math.log2(args.gpus)))
    n = args.nqubits + g
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(1, args.warmup + args.steps)):
        vals.append(cpu_reference_sample(n, args.dtype, budget_s=args.cpu_budget))
        if time.perf_counter() - t0 > 150:
            break
    vals = vals[min(args.warmup, len(vals) - 1):]
    base = vals[-1]
    # same unit as the GPU arm: gates x shards of 2^(n - log2 N) amplitudes (one host applies each gate to all N shards)
    v = args.gpus * float(np.mean([x["value"] for x in vals]))
    base["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "gates/s", "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": 1e3 * n_gates(n) / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128" if args.dtype == "complex128" else "c64", "data": "synthetic",
        "config": {"workload": f"QFT({n}) {args.dtype}, gate-by-gate on the host CPU (bounded sample, scaled)", "nqubits": n},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def verify_sharded_qft(runner, state, args, eng, dist):
    """Outside the timed region: QFT|x> on the shards against the closed form exp(2 pi i x k / 2^n) / sqrt(2^n) at sample
    amplitudes of every rank (a misplaced chunk or a wrong rank-specialised phase shows; the norm alone would not).
    If the all-to-all transport fails the check, fall back to the pairwise exchange kernel and check again."""
    import cmath

    import torch

    n = runner.n
    x = int("1011001110001011010111001010011011"[: n - 1] + "1", 2)
    rk, loc = runner.locate(x)
    rng = np.random.default_rng(17 + runner.rank)
    sample = sorted({0, (1 << runner.nlocal) - 1, *rng.integers(0, 1 << runner.nlocal, size=256).tolist()})
    tol = 1e-9 if args.dtype == "complex128" else 1e-4
    out = {}
    def path():
        if not runner.alltoall:
            return "pairwise"
        fused = "+perm" if (runner.fuse_perm and runner.alltoall_push) else ""  # experimental QB_A2A_FUSE_PERM=1
        return ("alltoall-push" if runner.alltoall_push else "alltoall-swap") + fused

    for _ in range(3):  # all-to-all out of place -> all-to-all in place -> pairwise exchanges
        attempt = path()
        st = state
        st.tensor.zero_()
        if rk == runner.rank:
            st.tensor[loc] = 1
        runner.run(st, timed=False)
        got = st.tensor[torch.as_tensor(sample, device=st.tensor.device)].cpu().numpy()
        err = 0.0
        for v, l in zip(got, sample):
            k = runner.canonical_index(runner.rank, int(l))
            phase = (x * k) % (1 << n)
            want = cmath.exp(2j * cmath.pi * (phase / float(1 << n))) / (2.0 ** (n / 2))
            err = max(err, abs(complex(v) - want) * 2.0 ** (n / 2))
        t = torch.tensor([err], device=st.tensor.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = {"basis_state": x, "samples_per_rank": len(sample), "max_rel_err": float(t.item()), "exchange_path": attempt}
        if out["max_rel_err"] < tol or attempt == "pairwise":
            break
        # every rank sees the same all-reduced error, so they switch together
        if runner.fuse_perm:
            raise AssertionError(f"QB_A2A_FUSE_PERM=1 (experimental) fails the closed-form check: {out}")
        if runner.alltoall_push:
            runner.alltoall_push = False
        else:
            runner.alltoall = False
    if not out["max_rel_err"] < tol:
        raise AssertionError(f"sharded QFT does not match the closed form: {out}")
    state.tensor.zero_()
    if runner.rank == runner.locate(0)[0]:
        state.tensor[0] = 1
    return out


# ----------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch

    from qibo_b200 import circuits
    from qibo_b200.engine import Engine

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
    else:
        state = eng.basis_state(n, args.dtype)
        # the gate program is compiled once and resident on the device, like the state ("inputs already resident in HBM
        # when the timed region starts"); the e2e leg below sends the host gate matrices in on every step instead
        compiled = eng.compile(n, args.dtype, ops)
        step = lambda: eng.run_program(compiled, state, timed=True)  # noqa: E731
        barrier = lambda: None  # noqa: E731

    verify = None
    if world > 1:
        verify = verify_sharded_qft(runner, state, args, eng, dist)
    for _ in range(args.warmup):
        stats = step()
    barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sweep_ms, nsweeps, xch_ms, xch_bytes, nxch, perm_ms, nperm, nxl, nstage = 0.0, 0, 0.0, 0, 0, 0.0, 0, 0, 0
    for _ in range(args.steps):
        stats = step()
        sweep_ms += stats.elapsed_ms
        nsweeps += stats.nsweeps
        nstage += getattr(stats, "nstage_sweeps", 0)
        perm_ms += getattr(stats, "perm_ms", 0.0)
        nperm += getattr(stats, "nperm", 0)
        xch_ms += getattr(stats, "exchange_ms", 0.0)
        xch_bytes += getattr(stats, "exchange_bytes", 0)
        nxch += getattr(stats, "nexchanges", 0)
        nxl += getattr(stats, "nexchange_launches", 0)
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
    # units: one gate applied to one 2^nlocal-amplitude shard; every rank applies every gate of the circuit to its
    # shard, so N ranks process N * gates units per step (tier rule: "the units all ranks processed / that time")
    circuit_gates_per_s = n_gates(n) * args.steps / (ms / 1e3)
    value = world * circuit_gates_per_s

    # roofline of the dominant kernel (sweep_kernel): algorithmic bytes per launch = 2 * B * 2^nlocal
    # (the out-of-place permutation kernel k8_permute that applies the final SWAP run moves the same bytes; it is
    # reported beside it, not averaged into the sweep kernel's launch time)
    bytes_per_sweep = 2.0 * itemsize * 2.0**nlocal
    avg_sweep_ms = (sweep_ms - perm_ms) / max(nsweeps - nperm, 1)
    achieved = bytes_per_sweep / (avg_sweep_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": profiled_traffic(nlocal, args.dtype), "kernel": "sweep_kernel", "peak_source": peak_src,
        "bytes_per_launch": bytes_per_sweep, "avg_launch_ms": avg_sweep_ms, "launches": nsweeps - nperm,
        "frac_of_nominal_8TBs": achieved / 8000.0,
        "all_launches": {"count": nsweeps, "avg_ms": sweep_ms / max(nsweeps, 1), "GBps": bytes_per_sweep / (sweep_ms / max(nsweeps, 1) * 1e-3) / 1e9,
                         "frac": bytes_per_sweep / (sweep_ms / max(nsweeps, 1) * 1e-3) / 1e9 / peak},
    }
    if nperm:
        roofline["k8_permute"] = {"launches": nperm, "avg_launch_ms": perm_ms / nperm, "GBps": bytes_per_sweep / (perm_ms / nperm * 1e-3) / 1e9}

    # e2e: the plugin-level call with host inputs: zero state, host gate program in, marginal probabilities out
    e2e = None
    if world == 1:
        mq = list(range(min(10, n)))
        h2d = sum(op.data.nbytes for op in ops) + len(ops) * 176
        d2h = 8 * 2 ** len(mq) * (1 if args.dtype == "complex128" else 0.5)

        def e2e_step():
            st = eng.basis_state(n, args.dtype)
            eng.apply_program(st, n, ops)
            return eng.probabilities(st, mq, n).numpy()

        del state
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            probs = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert abs(float(probs.sum()) - 1.0) < 1e-6
        e2e = {"value": n_gates(n) * args.steps / dt, "unit": "gates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "Engine.basis_state + apply_program(host gate matrices) + probabilities(10 qubits) -> host"}

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
            "config": {
                "workload": f"QFT({n}) {args.dtype}, {n_gates(n)} gates, zero initial state and compiled gate program resident in HBM, "
                f"{2 ** nlocal * itemsize / 2 ** 30:.0f} GiB per GPU", "nqubits": n, "global_qubits": g,
                "sweeps_per_step": nsweeps // args.steps,
                **({"stage_only_sweeps_per_step": nstage // args.steps} if world == 1 else {}), "l2": "state (>= 16 GiB) is far larger than the 126 MB L2",
                "parallelism": f"global-qubit sharding x{world}" if world > 1 else "single GPU",
                **({"layout": f"global qubits {list(runner.global_qubits)} (--layout {args.layout}); rank = their bits, shard index = "
                    "the other qubits in ascending order"} if world > 1 else {}),
            },
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(nsweeps + nxl), "whole_circuit_wall_s": ms_per_step / 1e3,
            "circuit_gates_per_s": circuit_gates_per_s,
            "value_definition": "gates applied x shards (N ranks each apply every gate to their 2^nlocal-amplitude shard) per second; equals circuit gates/s at N=1",
        }
        if world > 1:
            line["exchange"] = {
                "count_per_step": nxch // args.steps, "launches_per_step": nxl // args.steps, "ms_per_step_rank0": xch_ms / args.steps,
                "GBps_per_direction_rank0": (xch_bytes / 2) / max(xch_ms, 1e-9) / 1e6,
                "bytes_per_step_rank0": xch_bytes // args.steps, "transport": (("one all-to-all kernel per run of exchanges over NVLink peer memory (CUDA IPC)" + (", out of place: remote stores only" if runner.alltoall_push else ", in-place chunk swaps")) if (nxl < nxch or runner.alltoall_min <= 1) and runner.alltoall else "one swap kernel over NVLink peer memory (CUDA IPC)") if args.exchange == "p2p" else "NCCL send/recv over NVLink, half-shard pairwise",
            }
            line["verify"] = verify
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
    ap.add_argument("--layout", default="auto", choices=["auto", "block"],
                    help="N > 1: which qubits are global -- 'auto' picks the layout with the fewest exchanges (trailing qubits for a QFT, "
                         "as the reference's _DistributedQFT), 'block' the leading ones")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="global<->local exchange: NVLink peer-memory kernel or NCCL send/recv")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
