"""diagnostic: the reference arm's per-core rate under different task shapes (run on the GPU box)"""
import ctypes as C
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench          # noqa: E402
import util           # noqa: E402

WL = sys.argv[1] if len(sys.argv) > 1 else "json"


def fresh_worker(args):
    """round-1 shape: a fresh pipeline and buffer per task, whole 100k-event block"""
    block, reps = args
    ref = util.Ref()
    ref.parser(**bench.parser_kw(WL))
    for p, props in bench.WORKLOADS[WL]["filters"]:
        ref.filter(p, props)
    buf = C.create_string_buffer(block, len(block))
    nrec = ref.L.flbref_count_records(C.cast(buf, C.c_void_p), len(block))
    t0 = time.perf_counter()
    for _ in range(reps):
        out, n = C.c_void_p(), C.c_size_t()
        r = ref.L.flbref_filter_do(ref.cfg, C.cast(buf, C.c_void_p), len(block), nrec, b"bench", C.byref(out), C.byref(n))
        if r == 1 and out.value:
            ref.L.flbref_free(out)
    return time.perf_counter() - t0, nrec * reps


if __name__ == "__main__":
    cores = os.cpu_count()
    block = bench.make_block(WL)
    for n in (cores, cores // 2, cores // 4):
        with mp.get_context("fork").Pool(n) as pool:
            res = pool.map(fresh_worker, [(block, 1)] * n)
        print("fresh pipelines, %3d workers x 100k events: slowest %.3f s fastest %.3f s -> %.1f M lines/s" % (
            n, max(r[0] for r in res), min(r[0] for r in res), sum(r[1] for r in res) / max(r[0] for r in res) / 1e6), flush=True)
    pool = bench.RefPool(cores, [WL])
    for per in (100000, 78125, 31250):
        for rep in range(3):
            res = pool.pool.map(bench._ref_task, [(WL, pool.offs[WL][per], 1)] * cores, chunksize=1)
            print("persistent pipelines, %d workers x %d events (pass %d): slowest %.3f s fastest %.3f s -> %.1f M lines/s" % (
                cores, per, rep, max(r[0] for r in res), min(r[0] for r in res), sum(r[1] for r in res) / max(r[0] for r in res) / 1e6), flush=True)
    pool.close()
