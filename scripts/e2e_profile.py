"""Where the time of the plugin-level e2e step goes: phases separated by device synchronisation + allocator statistics."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
import torch  # noqa: E402

qibo, be = bench.plugin_backend("complex128")
from qibo import gates  # noqa: E402
from qibo.models import QFT  # noqa: E402


def stats():
    s = torch.cuda.memory_stats()
    return {k: s[k] for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "num_ooms")} | {
        "reserved_GiB": round(torch.cuda.memory_reserved() / 2**30, 1), "allocated_GiB": round(torch.cuda.memory_allocated() / 2**30, 1)}


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for step in range(4):
    t0 = sync()
    c = QFT(n)
    c.add(gates.M(*range(10)))
    t1 = sync()
    res = c(nshots=1000)
    t2 = sync()
    f = res.frequencies(binary=False)
    t3 = sync()
    print(f"step {step}: build {1e3 * (t1 - t0):.1f} ms, execute {1e3 * (t2 - t1):.1f} ms, frequencies {1e3 * (t3 - t2):.1f} ms", stats(), flush=True)
    del c, res, f
# the pieces of execute
eng = be.engine_gpu
from qibo_b200 import circuits  # noqa: E402

for step in range(3):
    t0 = sync()
    st = eng.basis_state(n, "complex128")
    t1 = sync()
    eng.apply_program(st, n, circuits.qft(n))
    t2 = sync()
    p = eng.probabilities(st, list(range(10)), n)
    t3 = sync()
    print(f"engine {step}: zero state {1e3 * (t1 - t0):.1f} ms, program {1e3 * (t2 - t1):.1f} ms, marginal {1e3 * (t3 - t2):.1f} ms", stats(), flush=True)
    del st, p

# the pieces of the plugin's execute_circuit, host time per call (no synchronisation added: what the host waits for)
import functools  # noqa: E402

acc = {}


def wrap(obj, name):
    fn = getattr(obj, name)

    @functools.wraps(fn)
    def inner(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[name] = acc.get(name, 0.0) + 1e3 * (time.perf_counter() - t)

    setattr(obj, name, inner)


for o, nm in ((be, "zero_state"), (be, "_compiled_circuit"), (eng, "run_program"), (eng, "_alloc"), (eng, "_reclaim"), (be, "calculate_probabilities"),
              (be, "sample_shots")):
    wrap(o, nm)
for step in range(int(os.environ.get("DETAIL_STEPS", 3))):
    acc.clear()
    t0 = sync()
    c = QFT(n)
    c.add(gates.M(*range(10)))
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    res = c(nshots=1000)
    e1.record()
    th = time.perf_counter()
    t1 = sync()
    acc["device_span_ms"] = e0.elapsed_time(e1)
    f = res.frequencies(binary=False)
    t2 = sync()
    print(f"detail {step}: host returns after {1e3 * (th - t0):.1f} ms, device done {1e3 * (t1 - t0):.1f} ms, frequencies {1e3 * (t2 - t1):.1f} ms; host ms per call:",
          {k: round(v, 1) for k, v in acc.items()}, flush=True)
    del c, res, f
