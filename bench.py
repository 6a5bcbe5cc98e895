#!/usr/bin/env python
"""bench.py -- log lines/s through the parser+filter chain (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): the north-star chain of BASELINE.json -- filter_parser with the
apache-combined regex parser (conf/parsers.conf:1-6) + filter_grep + filter_modify -- over
synthetic apache access-log events {"log": line} (SURVEY.md section 8d, C1 shape), 10 M events per
GPU per step by default (weak scaling).  A step is one pass of the chain over that batch.

  value ....... events/s with the batch resident in HBM (flbgpu_chain_do_device), CUDA events
  e2e ......... the same batch through flbgpu_chain_do(): pinned host input, host<->device copies
                and the malloc()ed host result inside the timed region
  roofline .... evaluation kernel (k_chain<false>: record decode + regex + filters) timed by CUDA
                events inside the library; algorithmic bytes = chain input + chain output
  cpu_baseline  the UNMODIFIED reference (oracle/_ref) on this box's host cores, bounded sample
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BASE_LINES = 100_000          # distinct synthetic lines; the batch tiles this block

# BASELINE.json configs[1]: "in_dummy 10M JSON lines -> filter_parser(json) + filter_grep + filter_modify on 1xB200"
# (value shapes of SURVEY.md section 8d C2) -- the default workload.
# BASELINE.json north_star: "apache-combined parser + grep + modify chain" -- reported beside it (key "north_star").
WORKLOADS = {
    "json": {
        "name": "configs[1]: 10M JSON lines -> filter_parser(json)+filter_grep(level ^(warn|error)$)+filter_modify(Add,Rename,Remove)",
        "filters": [("parser", [("Key_Name", "log"), ("Parser", "json")]),
                    ("grep", [("Regex", "level ^(warn|error)$")]),
                    ("modify", [("Add", "env prod"), ("Rename", "msg message"), ("Remove", "debug")])],
    },
    "apache": {
        "name": "north-star chain: filter_parser(apache regex)+filter_grep(method ^(GET|POST)$)+filter_modify over apache-combined events",
        "filters": [("parser", [("Key_Name", "log"), ("Parser", "apache")]),
                    ("grep", [("Regex", "method ^(GET|POST)$")]),
                    ("modify", [("Add", "env prod"), ("Rename", "code status"), ("Remove", "agent")])],
    },
    # configs[3] in small: filter_log_to_metrics histogram, 32 label sets, integer-valued observations (exact
    # fp64 sums), logs discarded; with N>1 every step ends with the NCCL all-reduce of the metric tables.
    # Not a default bench line: `--workload l2m`.
    "l2m": {
        "name": "configs[3]: filter_log_to_metrics histogram(duration) by color,direction; discard_logs; table all-reduce per step",
        "filters": [("log_to_metrics", [("metric_mode", "histogram"), ("metric_name", "duration"), ("metric_description", "d"),
                                        ("tag", "m"), ("value_field", "duration"), ("label_field", "color"),
                                        ("label_field", "direction"), ("discard_logs", "on")])],
    },
}
WL = "json"


def make_block(rank=0, wl=None):
    import util
    wl = wl or WL
    if wl == "l2m":
        import random
        rng = random.Random(0xF1B1 + 4 + rank)
        colors = [b"red", b"green", b"blue", b"cyan", b"black", b"white", b"pink", b"grey"]
        dirs = [b"north", b"south", b"east", b"west"]
        return b"".join(util.event(1700000000 + i, 0, [(b"duration", util.mp_str(str(rng.randint(0, 9999)).encode())),
                                                       (b"color", util.mp_str(rng.choice(colors))),
                                                       (b"direction", util.mp_str(rng.choice(dirs)))]) for i in range(BASE_LINES))
    if wl == "json":
        lines = util.json_lines(BASE_LINES, seed=0xF1B1 + 2 + rank)
    else:
        lines = util.apache_lines(BASE_LINES, seed=0xF1B1 + 1 + rank)
    return util.chunk_from_lines(lines)


def parser_kw(wl=None):
    import util
    if (wl or WL) == "json":
        return dict(name="json", format="json", time_fmt="%d/%b/%Y:%H:%M:%S %z", time_key="time")
    return dict(name="apache", format="regex", regex=util.APACHE_RX, time_fmt=util.APACHE_TIME_FMT, time_key="time")


# ------------------------------------------------------------------ reference arm
def _ref_worker(args):
    block, reps = args
    try:                     # the GPU arm binds its process near its GPU; the reference gets every core
        os.sched_setaffinity(0, range(os.cpu_count()))
    except Exception:
        pass
    import util
    ref = util.Ref()
    ref.parser(**parser_kw())
    for p, props in WORKLOADS[WL]["filters"]:
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


def reference_throughput(block, cores, reps=1):
    """All host cores, each its own reference pipeline (filters are single-threaded per
    pipeline in the reference) over its own copy of the block.  Returns (lines/s, seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(block, reps)] * cores)
        wall = time.perf_counter() - t0
    lines = sum(r[1] for r in res)
    busy = max(r[0] for r in res)
    return lines / busy, wall, lines


def run_reference(args):
    import util
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not util.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libflbref.so missing"}))
        return
    cores = os.cpu_count() or 1
    block = make_block()
    for _ in range(args.warmup):
        reference_throughput(block, cores, 1)
    lines, dt = 0, 0.0
    for _ in range(args.steps):
        v, _, n = reference_throughput(block, cores, 1)
        lines += n
        dt += n / v                    # slowest worker's time inside the reference calls (pool start-up excluded)
    val = lines / dt
    sample = "%d cores x %d-event block per step, %d steps" % (cores, BASE_LINES, args.steps)
    print(json.dumps({
        "impl": "reference", "metric": "log lines/sec through parser+filter chain", "value": val, "unit": "lines/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOADS[WL]["name"], "events_per_step": cores * BASE_LINES, "host_cores": cores},
        "cpu_baseline": {"value": val, "unit": "lines/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------ our arm
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measure(args, wl, L, ctx, torch, dist, rank, world, local):
    """value / e2e / kernel times of one workload on this rank's GPU"""
    import util
    pkg = util.pkg
    if wl != "l2m":
        ctx.parser(**parser_kw(wl))
    filters = [ctx.filter(p, props) for p, props in WORKLOADS[wl]["filters"]]
    chain = ctx.chain(filters)
    block = make_block(rank, wl)
    reps = max(1, args.lines // BASE_LINES)
    n_lines = reps * BASE_LINES
    nbytes = len(block) * reps
    pageable = os.environ.get("FLBGPU_BENCH_PAGEABLE") == "1"   # experiment: ordinary malloc()ed input, staged by the library
    if pageable:
        libc0 = C.CDLL(None)
        libc0.malloc.restype = C.c_void_p; libc0.malloc.argtypes = [C.c_size_t]
        h_in = libc0.malloc(nbytes)
    else:
        h_in = L.flbgpu_host_alloc(ctx.h, nbytes)       # pinned
    for i in range(reps):
        C.memmove(h_in + i * len(block), block, len(block))
    d_in = L.flbgpu_dev_alloc(ctx.h, nbytes + 64)
    out_cap = nbytes + nbytes // 2 + 64
    d_out = L.flbgpu_dev_alloc(ctx.h, out_cap)
    assert h_in and d_in and d_out, "allocation failed"
    L.flbgpu_dev_upload(ctx.h, d_in, h_in, nbytes)
    if os.environ.get("FLBGPU_BENCH_DEBUG"):
        torch.cuda.synchronize()
        t_up = time.perf_counter()
        L.flbgpu_dev_upload(ctx.h, d_in, h_in, nbytes)
        torch.cuda.synchronize()
        t_up = time.perf_counter() - t_up
        sys.stderr.write("plain pinned H2D of the input: %.1f ms = %.1f GB/s\n" % (1e3 * t_up, nbytes / t_up / 1e9))
    stream = torch.cuda.ExternalStream(L.flbgpu_stream(ctx.h), device=torch.device("cuda", local))
    osz = C.c_size_t()

    def step_device():
        r = L.flbgpu_chain_do_device(chain.h, d_in, nbytes, d_out, out_cap, C.byref(osz))
        if r != pkg.FILTER_MODIFIED:
            raise RuntimeError("chain_do_device -> %d: %s" % (r, ctx.err()))
        if wl == "l2m" and world > 1:
            filters[0].l2m_allreduce()          # the one exchange of the path: metric tables over NCCL
            L.flbgpu_l2m_reset(filters[0].h)    # "flushed": the next interval starts from zero

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    st0 = chain.stats()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    kms = [0.0, 0.0, 0.0]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    ms3 = (C.c_float * 3)()
    for _ in range(args.steps):
        step_device()
        L.flbgpu_kernel_ms(ctx.h, ms3)
        for k in range(3):
            kms[k] += ms3[k]
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    st1 = chain.stats()
    launches = int(st1.kernel_launches - st0.kernel_launches)
    out_bytes = osz.value
    t = torch.tensor([dev_ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * n_lines * args.steps / (dev_ms / 1000.0)

    # ---- end to end through the host-buffer C ABI
    out_p = C.c_void_p()
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]

    def step_host():
        r = L.flbgpu_chain_do(chain.h, h_in, nbytes, b"bench", 5, C.byref(out_p), C.byref(osz))
        if r != pkg.FILTER_MODIFIED:
            raise RuntimeError("chain_do -> %d: %s" % (r, ctx.err()))
        if wl == "l2m" and world > 1:
            filters[0].l2m_allreduce()
            L.flbgpu_l2m_reset(filters[0].h)
        t_free = time.perf_counter()
        libc.free(out_p)
        free_s[0] += time.perf_counter() - t_free

    free_s = [0.0]
    e2e_steps = max(1, args.steps)
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    if os.environ.get("FLBGPU_BENCH_DEBUG"):
        sys.stderr.write("e2e %s: %.1f ms/step, of which free() of the result %.1f ms/step\n" % (wl, 1e3 * e2e_s / e2e_steps, 1e3 * free_s[0] / (e2e_steps + 1)))
    t = torch.tensor([e2e_s], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e = world * n_lines * e2e_steps / e2e_s
    phases = [round(float(x), 2) for x in chain.stats().phase_ms]
    L.flbgpu_dev_free(ctx.h, d_in)
    L.flbgpu_dev_free(ctx.h, d_out)
    if pageable:
        libc.free(h_in)
    else:
        L.flbgpu_host_free(ctx.h, h_in)
    return {"value": value, "dev_ms": dev_ms, "e2e": e2e, "e2e_steps": e2e_steps, "kms": [k / args.steps for k in kms],
            "launches": launches, "n_lines": n_lines, "nbytes": nbytes, "out_bytes": out_bytes, "clocks": clocks,
            "phases": phases}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import util
    pkg = util.pkg
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    L = pkg.load()                                  # raises if the CUDA library is missing
    ctx = pkg.Context(local, lib=L)
    # Host allocator policy of the embedding process: keep freed result buffers in the heap instead of
    # returning them to the kernel (glibc: no mmap for big blocks, no trimming) -- what Fluent Bit's
    # default jemalloc build does with its retained extents.  Without it every step pays ~300k page
    # faults for the fresh result buffer.  FLBGPU_BENCH_DEFAULT_MALLOC=1 turns the tuning off.
    if os.environ.get("FLBGPU_BENCH_DEFAULT_MALLOC") != "1":
        libc = C.CDLL(None)
        libc.mallopt(-4, 0)              # M_MMAP_MAX = 0
        libc.mallopt(-1, (1 << 31) - 1)  # M_TRIM_THRESHOLD
        libc.mallopt(-2, 64 << 20)       # M_TOP_PAD (small, so that a freed GB-size result does not push the top over the trim threshold)

    m = measure(args, WL, L, ctx, torch, dist, rank, world, local)
    other = "apache" if WL == "json" else "json"
    m2 = None
    if not args.primary_only:
        m2 = measure(args, other, L, ctx, torch, dist, rank, world, local)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:
        per_event = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_chain_eval_dram_bytes_per_event_" + WL)
        traffic = per_event * m["n_lines"] if per_event else None      # ncu dram read+write of the evaluation launches of one step
    except Exception:
        pass

    def roof(mm):
        eval_ms = mm["kms"][1]
        alg = mm["nbytes"] + mm["out_bytes"]
        ach = alg / (eval_ms / 1000.0) / 1e9 if eval_ms > 0 else None
        return eval_ms, alg, ach

    eval_ms, alg_bytes, achieved = roof(m)

    cpu = None
    if util.have_ref():
        cores = os.cpu_count() or 1
        v, wall, n = reference_throughput(make_block(), cores, 1)
        cpu = {"value": v, "unit": "lines/s", "cores": cores, "kind": "reference",
               "sample": "%d cores x one %d-event block each (%.1f s wall)" % (cores, BASE_LINES, wall)}

    line = {
        "metric": "log lines/sec through parser+filter chain", "value": m["value"], "unit": "lines/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["dev_ms"] / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOADS[WL]["name"], "events_per_gpu_per_step": m["n_lines"], "input_bytes_per_gpu": m["nbytes"],
                   "output_bytes_per_gpu": m["out_bytes"], "distinct_lines": BASE_LINES,
                   "l2": "input (%.0f MB) and output larger than the 126 MB L2" % (m["nbytes"] / 1e6),
                   "parallelism": ("record shards; one NCCL all-reduce of the metric table per step" if WL == "l2m" else "record shards, no data-path collective"),
                   "host_malloc": "default" if os.environ.get("FLBGPU_BENCH_DEFAULT_MALLOC") == "1" else "glibc tuned to retain freed result buffers (M_MMAP_MAX=0, M_TRIM_THRESHOLD=2GiB)"},
        "e2e": {"value": m["e2e"], "unit": "lines/s", "h2d_bytes_per_step": m["nbytes"], "d2h_bytes_per_step": m["out_bytes"],
                "steps": m["e2e_steps"], "timing": "wall clock between device-synchronising barriers",
                "host_phase_ms_last_call": dict(zip(["upload+index+evaluate", "size_scan", "emit+download", "total"], m["phases"]))},
        "gpu_launches": m["launches"],
        "kernel_ms_per_step": {"index": m["kms"][0], "evaluate": eval_ms, "emit": m["kms"][2]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "kernel": "k_chain_eval (evaluation pass)", "algorithmic_bytes_per_launch": alg_bytes,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
        "cpu_baseline": cpu,
        "clocks": m["clocks"],
    }
    if m2:
        e2, a2, ach2 = roof(m2)
        line["north_star" if other == "apache" else "configs1_json"] = {
            "workload": WORKLOADS[other]["name"], "value": m2["value"], "e2e": m2["e2e"], "unit": "lines/s",
            "events_per_gpu_per_step": m2["n_lines"], "input_bytes_per_gpu": m2["nbytes"], "output_bytes_per_gpu": m2["out_bytes"],
            "kernel_ms_per_step": {"index": m2["kms"][0], "evaluate": e2, "emit": m2["kms"][2]},
            "roofline_frac": (ach2 / peak) if ach2 else None, "gpu_launches": m2["launches"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lines", type=int, default=10_000_000, help="events per GPU per step")
    ap.add_argument("--workload", default="json", choices=["json", "apache", "l2m"], help="primary workload (the other one is reported beside it)")
    ap.add_argument("--primary-only", action="store_true")
    args = ap.parse_args()
    global WL
    WL = args.workload
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
