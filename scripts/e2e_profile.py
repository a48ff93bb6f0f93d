"""Where the host time of the plugin-level e2e step goes: cProfile of `QFT(n) + M -> c(nshots).frequencies()`."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
import torch  # noqa: E402

qibo, be = bench.plugin_backend("complex128")
for _ in range(2):
    bench.plugin_qft_step(n, 10, 1000)
torch.cuda.synchronize()
t0 = time.perf_counter()
bench.plugin_qft_step(n, 10, 1000)
torch.cuda.synchronize()
print("step wall ms", 1e3 * (time.perf_counter() - t0))
pr = cProfile.Profile()
pr.enable()
bench.plugin_qft_step(n, 10, 1000)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue())
