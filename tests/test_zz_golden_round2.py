"""Committed golden vectors of filter_multiline and of the chunk -> JSON conversion (tests/golden/multiline_vectors.json,
tojson_vectors.json: outputs of the UNMODIFIED reference, made by tests/golden/make_round2_rows.py).  They need neither
/root/reference nor oracle/_ref: the CPU emulation of the device code here, the GPU on the box."""
import base64
import json
import os

import pytest

import util

pkg = util.pkg
G = os.path.join(util.ROOT, "tests", "golden")
D = lambda s: None if s is None else base64.b64decode(s)


def _multiline(lib):
    vec = json.load(open(os.path.join(G, "multiline_vectors.json")))
    assert len(vec) >= 9
    for v in vec:
        ctx = pkg.Context(0, lib=lib)
        if not v.get("builtin"):
            ctx.ml_parser(v["name"], type=v["type"], rules=[tuple(r) for r in v["rules"]], match_string=v["match_string"], negate=v["negate"])
        ch = ctx.chain([ctx.filter("multiline", [tuple(p) for p in v["props"]])])
        for k, call in enumerate(v["calls"]):
            assert ch.do(D(call["in"])) == (call["ret"], D(call["out"])), (v["name"], k)


def _tojson(lib):
    vec = json.load(open(os.path.join(G, "tojson_vectors.json")))
    ctx = pkg.Context(0, lib=lib)
    n = 0
    for v in vec:
        got, und = ctx.to_json(D(v["in"]), v["json_format"], v["date_format"], v["date_key"], v["escape_unicode"])
        if und:
            continue
        assert got == D(v["out"]), v
        n += 1
    assert n >= 30


def test_multiline_golden_hostsim(sim_lib):
    _multiline(sim_lib)


def test_tojson_golden_hostsim(sim_lib):
    _tojson(sim_lib)


@pytest.mark.gpu
def test_multiline_golden_gpu(gpu_lib):
    _multiline(gpu_lib)


@pytest.mark.gpu
def test_tojson_golden_gpu(gpu_lib):
    _tojson(gpu_lib)
