"""The C-ABI boundary: libflbgpu.so loads on a box without a GPU, exports every function that
include/flbgpu.h declares (no compute is called here), refuses to initialise without a device, and
never links or loads anything under oracle/."""
import ctypes as C
import os
import re
import subprocess

import pytest

import util

HDR = os.path.join(util.ROOT, "include", "flbgpu.h")
LIB = os.path.join(util.ROOT, "fluent-bit_b200", "libflbgpu.so")


def declared():
    text = open(HDR).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flbgpu_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    names = declared()
    assert len(names) >= 30
    lib = C.CDLL(LIB)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_exports_are_plain_c():
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(not s.startswith("_Z") for s in syms if s.startswith("flbgpu") or "flbgpu" in s)
    assert set(declared()) <= set(syms)


def test_product_does_not_reach_the_oracle():
    out = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    assert "oracle" not in out and "flbref" not in out and "hostsim" not in out
    blob = open(LIB, "rb").read()
    assert b"liboracle" not in blob and b"libflbref" not in blob and b"libhostsim" not in blob


def test_init_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = util.pkg.load()
    with pytest.raises(util.pkg.FlbGpuError) as e:
        util.pkg.Context(0, lib=lib)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_stats_after_a_call(sim_lib, ref_available):
    """flbgpu_chain_stats: records and bytes in/out of the last call"""
    import cases
    ctx = util.pkg.Context(0, lib=sim_lib)
    ctx.parser(**cases.AP)
    chain = ctx.chain([ctx.filter(p, props) for p, props in [cases.P, ("grep", [("Regex", "method ^(GET|POST)$")])]])
    lines = util.apache_lines(500, seed=2)
    chunk = util.chunk_from_lines(lines)
    r, out = chain.do(chunk)
    st = chain.stats()
    assert st.records_in == len(lines) and st.bytes_in == len(chunk) and st.bytes_out == len(out)
    assert st.records_out == len(util.split_records(out)) and 0 < st.records_out < st.records_in
