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


def _ml_reference_tables(lib):
    """tests/internal/multiline.c: the input / expected-output tables of the java, ruby, python, go, elastic and endswith tests
    (tests/golden/ml_scenarios.json, read from the C file by make_ml_scenarios.py).  The inputs as one chunk of {"log": line}
    events: byte for byte what the reference's filter_multiline makes of them, and -- where that path and the test's own
    (flb_ml_append_text, which puts a line feed behind a message when it flushes) agree -- the messages of the test's table."""
    vec = json.load(open(os.path.join(G, "ml_scenarios.json")))
    assert [v["name"] for v in vec] == ["java", "ruby", "python", "go", "elastic", "endswith"]
    for v in vec:
        ctx = pkg.Context(0, lib=lib)
        if "rules" in v:
            ctx.ml_parser(v["parser"], rules=[tuple(r) for r in v["rules"]])
        if v["name"] == "endswith":
            ctx.ml_parser(v["parser"], type=v["type"], match_string=v["match_string"], negate=v["negate"])
        ch = ctx.chain([ctx.filter("multiline", [("multiline.parser", v["parser"]), ("multiline.key_content", "log"), ("buffer", "off")])])
        chunk = util.chunk_from_lines([D(x) for x in v["input"]])
        ret, out = ch.do(chunk)
        assert (ret, out) == (v["ret"], D(v["out"])), v["name"]
        msgs = []
        for o, l in util.split_records(out):
            rec = out[o:o + l]
            k = rec.index(b"\xa3log") + 4
            h = rec[k]
            n, hl = (h & 31, 1) if h < 0xc0 else (rec[k + 1], 2) if h == 0xd9 else (int.from_bytes(rec[k + 1:k + 3], "big"), 3) if h == 0xda else (int.from_bytes(rec[k + 1:k + 5], "big"), 5)
            msgs.append(rec[k + hl:k + hl + n])
        exp = [D(x) for x in v["expected"]]
        assert len(msgs) == len(exp), v["name"]
        if v["name"] in ("java", "python", "elastic"):
            assert msgs == exp, v["name"]


def test_multiline_reference_tables_hostsim(sim_lib):
    _ml_reference_tables(sim_lib)


@pytest.mark.gpu
def test_multiline_reference_tables_gpu(gpu_lib):
    _ml_reference_tables(gpu_lib)


def _write_str_cases(lib):
    """tests/internal/utils.c: the reference's own cases for flb_utils_write_str() (escaped and raw; invalid leading / trailing
    bytes, special bytes, edge cases), each as the value of a string member: the text between its quotes is the test's
    expected output"""
    vec = json.load(open(os.path.join(G, "write_str_cases.json")))
    assert len(vec) >= 20
    ctx = pkg.Context(0, lib=lib)
    for c in vec:
        ev = util.event(1700000000, 0, [(b"s", util.mp_str(D(c["input"]))), (b"tail", util.mp_str(b"a plain key behind it"))])
        got, und = ctx.to_json(ev, 3, 2, "d", c["escape_unicode"])
        assert und == 0 and got == b'{"d":1700000000,"s":"' + D(c["expected"]) + b'","tail":"a plain key behind it"}\n', c["test"]


def _pack_fixtures(lib):
    """tests/internal/pack.c: test_utf8_to_json (data/pack/*.mp, each a msgpack string, against the .json beside it) and
    test_json_date_* (one legacy event, the date key in every format)"""
    ctx = pkg.Context(0, lib=lib)
    vec = json.load(open(os.path.join(G, "pack_to_json_cases.json")))
    assert len(vec) >= 7
    for v in vec:
        ev = util.event(1700000000, 0, [(b"s", D(v["msgpack"])), (b"tail", util.mp_str(b"a plain key behind it"))])
        got, und = ctx.to_json(ev, 3, 2, "d", True)
        assert und == 0 and got == b'{"d":1700000000,"s":' + D(v["json"]) + b',"tail":"a plain key behind it"}\n', v["name"]
    legacy = bytes([0x92, 0xd7, 0x00, 0x07, 0x5b, 0xcd, 0x15, 0x07, 0x5b, 0xcd, 0x15, 0x81, 0xa2, 0x61, 0x61, 0xa2, 0x62, 0x62])
    for fmt, want in ((1, b"1973-11-29T21:33:09.123456Z"), (0, b"123456789.123456"), (3, b"1973-11-29 21:33:09.123456"), (2, b'"date":123456789,'),
                      (4, b'"date":123456789123,')):
        got, und = ctx.to_json(legacy, 1, fmt, "date", True)
        assert want in got and got.startswith(b"[{") and got.endswith(b'"aa":"bb"}]'), (fmt, got)


def test_pack_fixtures_hostsim(sim_lib):
    _pack_fixtures(sim_lib)


@pytest.mark.gpu
def test_pack_fixtures_gpu(gpu_lib):
    _pack_fixtures(gpu_lib)


def test_write_str_cases_hostsim(sim_lib):
    _write_str_cases(sim_lib)


@pytest.mark.gpu
def test_write_str_cases_gpu(gpu_lib):
    _write_str_cases(gpu_lib)


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
