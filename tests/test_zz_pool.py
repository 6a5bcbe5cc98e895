"""flbgpu_pool_*: a batch of chunks over several chains (one per device on a GPU box; several contexts of the CPU emulation
here), every result in its chunk's slot and equal to what one chain makes of that chunk."""
import ctypes as C
import random

import pytest

import util

pkg = util.pkg


def _pool(lib, contexts):
    ap = dict(name="apache", format="regex", regex=util.APACHE_RX, time_fmt=util.APACHE_TIME_FMT, time_key="time")
    flt = [("parser", [("Key_Name", "log"), ("Parser", "apache")]), ("grep", [("Regex", "method ^(GET|POST)$")]), ("modify", [("Add", "env prod")])]
    ctxs, chains = [], []
    for d in contexts:
        ctx = pkg.Context(d, lib=lib)
        ctx.parser(**ap)
        ctxs.append(ctx)
        chains.append(ctx.chain([ctx.filter(p, props) for p, props in flt]))
    one = chains[0]
    rng = random.Random(51)
    chunks = [util.chunk_from_lines(util.apache_lines(rng.choice([1, 50, 400, 3000]), seed=s)) for s in range(23)]
    chunks += [util.chunk_from_lines([b"not an access line"] * 5), b"\xc1\xc1"]             # nothing parsed (still re-encoded), nothing decodable
    want = [one.do(c, tag="bench") for c in chunks]
    L = ctxs[0].L
    arr = (C.c_void_p * len(chains))(*[ch.h for ch in chains])
    pool = L.flbgpu_pool_new(arr, len(chains))
    assert pool
    n = len(chunks)
    bufs = [C.create_string_buffer(c, len(c)) for c in chunks]
    data = (C.c_void_p * n)(*[C.cast(b, C.c_void_p) for b in bufs])
    sizes = (C.c_size_t * n)(*[len(c) for c in chunks])
    outs, osz, rets = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
    for _ in range(3):
        assert L.flbgpu_pool_do(pool, n, data, sizes, b"bench", 5, outs, osz, rets) == 0
        for i in range(n):
            got = (rets[i], C.string_at(outs[i], osz[i]) if rets[i] == pkg.FILTER_MODIFIED and osz[i] else (b"" if rets[i] == pkg.FILTER_MODIFIED else None))
            assert got == want[i], i
            if outs[i]:
                pkg._libc.free(C.c_void_p(outs[i]))
                outs[i] = None
    assert L.flbgpu_pool_do(pool, 0, None, None, b"", 0, None, None, None) == 0
    L.flbgpu_pool_destroy(pool)


def test_pool_hostsim(sim_lib):
    _pool(sim_lib, [0, 0, 0])


@pytest.mark.gpu
def test_pool_gpu(gpu_lib):
    n = gpu_lib.flbgpu_device_count()
    _pool(gpu_lib, list(range(min(n, 4))) if n > 1 else [0, 0])
