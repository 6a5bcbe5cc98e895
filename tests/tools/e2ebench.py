"""whole-set end-to-end calls (pageable input, glibc's untouched malloc) under the environment's FLBGPU_* knobs:
   python tests/tools/e2ebench.py [json|apache] [events] -> one line with M lines/s and the host phases"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench          # noqa: E402
import util           # noqa: E402


class A:
    lines = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    l2m_total = 100_000_000
    steps = 3
    warmup = 1


wl = sys.argv[1] if len(sys.argv) > 1 else "json"
L = util.pkg.load()
ctx = util.pkg.Context(0, lib=L)
w = bench.Workload(A, wl, L, ctx, 0, 1)
w.step_host()
t0 = time.perf_counter()
for _ in range(A.steps):
    out = w.step_host()
dt = (time.perf_counter() - t0) / A.steps
st = w.chain.stats()
print("%s %s: %.1f M lines/s, %.1f ms/step, last call total %.1f ms, out %d MB" % (
    wl, " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("FLBGPU_")), w.n_lines / dt / 1e6, 1e3 * dt, st.phase_ms[3], out >> 20))
