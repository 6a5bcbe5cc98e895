"""filter_log_to_metrics: metric tables and filter results against the UNMODIFIED reference
(oracle/_ref), on the CPU emulation and on the GPU; plus the cross-rank merge (gloo, world 2)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import pytest

import l2m_cases
import util

pkg = util.pkg
TS = re.compile(r"^\S+Z ", re.M)


def ref_text(ref, f):
    ref.L.flbref_l2m_cmt_text.restype = C.c_void_p
    ref.L.flbref_l2m_cmt_text.argtypes = [C.c_void_p]
    p = ref.L.flbref_l2m_cmt_text(f)
    t = C.string_at(p).decode(errors="replace")
    ref.L.flbref_cfree(C.c_void_p(p))
    return TS.sub("", t)


def run_case(lib, parsers, filters, chunk, k, calls=2):
    ctx = pkg.Context(0, lib=lib)
    ref = util.Ref()
    for kw in parsers:
        ctx.parser(**kw)
        ref.parser(**kw)
    fs = [ctx.filter(p, props) for p, props in filters]
    rfs = [ref.filter(p, props) for p, props in filters]
    chain = ctx.chain(fs)
    for _ in range(calls):                       # the table accumulates over calls, like ctx->cmt
        want = ref.chain_do(chunk)
        got = chain.do(chunk)
        assert got[0] == want[0]
        assert got[1] == want[1]
        assert fs[k].l2m_text() == ref_text(ref, rfs[k])


@pytest.mark.parametrize("case", l2m_cases.L2M_CASES, ids=[c[0] for c in l2m_cases.L2M_CASES])
def test_l2m_hostsim(case, sim_lib, ref_available):
    _, parsers, filters, mk, k = case
    run_case(sim_lib, parsers, filters, mk(), k)


@pytest.mark.gpu
@pytest.mark.parametrize("case", l2m_cases.L2M_CASES, ids=[c[0] for c in l2m_cases.L2M_CASES])
def test_l2m_gpu(case, gpu_lib, ref_available):
    _, parsers, filters, mk, k = case
    run_case(gpu_lib, parsers, filters, mk(), k)


def check_golden(lib):
    vec = json.load(open(os.path.join(util.ROOT, "tests", "golden", "l2m_vectors.json")))
    assert len(vec) >= 8
    for v in vec:
        ctx = pkg.Context(0, lib=lib)
        for kw in v["parsers"]:
            ctx.parser(**kw)
        fs = [ctx.filter(p, [tuple(x) for x in props]) for p, props in v["filters"]]
        r, out = ctx.chain(fs).do(bytes.fromhex(v["in_hex"]))
        assert r == v["ret"], v["name"]
        assert (out.hex() if out is not None else None) == v["out_hex"], v["name"]
        assert fs[v["k"]].l2m_text() == v["text"], v["name"]


def test_l2m_golden_hostsim(sim_lib):
    check_golden(sim_lib)


@pytest.mark.gpu
def test_l2m_golden_gpu(gpu_lib):
    check_golden(gpu_lib)


@pytest.mark.parametrize("bad", [
    [("metric_mode", "counter"), ("metric_description", "d")],                              # no tag
    [("metric_mode", "counter"), ("tag", "t")],                                             # no description
    [("metric_mode", "histogram"), ("metric_description", "d"), ("tag", "t")],              # no value_field
    [("metric_mode", "nope"), ("metric_description", "d"), ("tag", "t")],
    [("metric_description", "d"), ("tag", "t"), ("regex", "onlyfield")],
    [("metric_description", "d"), ("tag", "t"), ("add_label", "just_one")],
])
def test_l2m_config_errors(bad, sim_lib, ref_available):
    ctx = pkg.Context(0, lib=sim_lib)
    with pytest.raises(pkg.FlbGpuError):
        ctx.filter("log_to_metrics", bad)
    with pytest.raises(RuntimeError):
        util.Ref().filter("log_to_metrics", bad)


def test_l2m_not_supported_is_loud(sim_lib):
    """a value text sscanf("%lf") cannot convert leaves the previous record's value in place in the
    reference (order dependent): refused, not guessed"""
    ctx = pkg.Context(0, lib=sim_lib)
    for mode in ("gauge", "histogram"):
        f = ctx.filter("log_to_metrics", [("metric_mode", mode), ("value_field", "x"), ("metric_description", "d"), ("tag", "t")])
        chunk = util.event(1700000000, 0, [(b"x", util.mp_str(b"1.5"))]) + util.event(1700000001, 0, [(b"x", util.mp_str(b"fast"))])
        with pytest.raises(pkg.FlbGpuError):
            f.cb(chunk)


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch.distributed as dist
import util, l2m_cases
pkg = util.pkg
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
name, parsers, filters, mk, k = [c for c in l2m_cases.L2M_CASES if c[0] == {case!r}][0]
chunk = mk()
recs = util.split_records(chunk)
per = (len(recs) + world - 1) // world
mine = recs[rank * per:(rank + 1) * per]
shard = chunk[mine[0][0]: mine[-1][0] + mine[-1][1]]
ctx = pkg.Context(0, lib=pkg.load(util.HOSTSIM_SO))
for kw in parsers:
    ctx.parser(**kw)
fs = [ctx.filter(p, props) for p, props in filters]
ctx.chain(fs).do(shard)
fs[k].l2m_allreduce()
open({out!r} + str(rank), "w").write(fs[k].l2m_text())
dist.destroy_process_group()
"""


@pytest.mark.parametrize("case", ["counter_labels", "histogram_default_buckets", "after_parser_and_grep", "gauge_labels"])
def test_l2m_allreduce_world2(case, sim_lib, ref_available, tmp_path):
    """Two shards of one chunk, one table per rank, merged with the all-reduce: every rank ends
    with exactly the table the reference builds from the whole chunk (same order, same values)."""
    name, parsers, filters, mk, k = [c for c in l2m_cases.L2M_CASES if c[0] == case][0]
    ref = util.Ref()
    for kw in parsers:
        ref.parser(**kw)
    rfs = [ref.filter(p, props) for p, props in filters]
    ref.chain_do(mk())
    want = ref_text(ref, rfs[k])
    out = str(tmp_path / "t")
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=util.ROOT, case=case, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29600 + os.getpid() % 300), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    for r in range(2):
        assert open(out + str(r)).read() == want


LIB_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import util, l2m_cases
pkg = util.pkg
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
gpu = {gpu!r}
name, parsers, filters, mk, k = [c for c in l2m_cases.L2M_CASES if c[0] == {case!r}][0]
chunk = mk()
recs = util.split_records(chunk)
per = (len(recs) + world - 1) // world
mine = recs[rank * per:(rank + 1) * per]
shard = chunk[mine[0][0]: mine[-1][0] + mine[-1][1]]
lib = pkg.load() if gpu else pkg.load(util.HOSTSIM_SO)
ctx = pkg.Context(rank if gpu else 0, lib=lib)
# the 128-byte communicator id travels over the embedding process's own channel: here a file
idf = {out!r} + ".id"
if rank == 0:
    uid = ctx.comm_unique_id()
    open(idf + ".tmp", "wb").write(uid); os.rename(idf + ".tmp", idf)
else:
    import time
    while not os.path.exists(idf):
        time.sleep(0.01)
    uid = open(idf, "rb").read()
ctx.comm_init(world, rank, uid)
for kw in parsers:
    ctx.parser(**kw)
fs = [ctx.filter(p, props) for p, props in filters]
ctx.chain(fs).do(shard)
fs[k].l2m_allreduce_lib()              # flbgpu_l2m_allreduce(): the library's own exchange
open({out!r} + str(rank), "w").write(fs[k].l2m_text())
"""


def _lib_allreduce(case, tmp_path, gpu):
    name, parsers, filters, mk, k = [c for c in l2m_cases.L2M_CASES if c[0] == case][0]
    ref = util.Ref()
    for kw in parsers:
        ref.parser(**kw)
    rfs = [ref.filter(p, props) for p, props in filters]
    ref.chain_do(mk())
    want = ref_text(ref, rfs[k])
    out = str(tmp_path / "t")
    script = tmp_path / "w.py"
    script.write_text(LIB_WORKER.format(root=util.ROOT, case=case, out=out, gpu=gpu))
    env = dict(os.environ, WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    for r in range(2):
        assert open(out + str(r)).read() == want


@pytest.mark.parametrize("case", ["counter_labels", "histogram_default_buckets", "after_parser_and_grep", "gauge_labels", "counter_no_labels"])
def test_l2m_library_allreduce_world2(case, sim_lib, ref_available, tmp_path):
    """flbgpu_l2m_allreduce() -- the merge logic of the PRODUCT library (runtime.c), two processes; the collectives
    underneath are the CPU emulation's (shared memory) here and NCCL on the GPU box"""
    _lib_allreduce(case, tmp_path, gpu=False)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["counter_labels", "histogram_default_buckets", "gauge_labels"])
def test_l2m_library_allreduce_nccl(case, gpu_lib, ref_available, tmp_path):
    """the same over NCCL: needs two GPUs"""
    if gpu_lib.flbgpu_device_count() < 2:
        pytest.skip("one GPU: the NCCL exchange needs two")
    _lib_allreduce(case, tmp_path, gpu=True)
