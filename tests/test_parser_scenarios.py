"""The reference's internal parser tests (tests/internal/parser_{regex,json,ltsv,logfmt}.c) replayed
through flbgpu_parser_do(): the same parser definition and input line, compare_msgpack()'s check
(every expected key/value-text pair is in the map, values compared the way that helper does), the
timestamp the test expects, and byte equality with the unmodified reference's map and time
(tests/golden/parser_scenarios.json, made by tests/golden/make_parser_scenarios.py)."""
import json
import os

import msgpack
import pytest

import util

pkg = util.pkg
SCENARIOS = json.load(open(os.path.join(util.ROOT, "tests", "golden", "parser_scenarios.json")))
IDS = ["%s:%s" % (s["parser"]["format"], s["test"]) for s in SCENARIOS]


def value_matches(text, v):
    """msgpack_strncmp() of the reference's test helper"""
    if isinstance(v, bytes):
        return v == text.encode()
    if isinstance(v, bool):
        return text.lower() == ("true" if v else "false")
    if isinstance(v, int):
        return int(text) == v
    if isinstance(v, float):
        return abs(float(text) - v) < 2.220446049250313e-16
    return False


def check(parse, sc):
    r, data, t = parse(sc["input"].encode())
    assert r == sc["ret"] and r != -1
    assert data.hex() == sc["out_hex"]                        # byte equality with the reference
    assert list(t) == sc["out_time"]
    m = msgpack.unpackb(data, raw=True, strict_map_key=False, object_pairs_hook=list)
    found = sum(1 for k, text in sc["pairs"] if any(mk == k.encode() and value_matches(text, mv) for mk, mv in m))
    assert found == len(sc["pairs"])                          # compare_msgpack(): num == expected pairs
    if sc["time"]:
        assert list(t) == sc["time"]


def test_every_format_is_covered():
    assert len(SCENARIOS) == 17
    assert {s["parser"]["format"] for s in SCENARIOS} == {"regex", "json", "ltsv", "logfmt"}


@pytest.mark.parametrize("sc", SCENARIOS, ids=IDS)
def test_parser_scenario_hostsim(sc, sim_lib):
    ctx = pkg.Context(0, lib=sim_lib)
    check(ctx.parser(**sc["parser"]).do, sc)


@pytest.mark.gpu
@pytest.mark.parametrize("sc", SCENARIOS, ids=IDS)
def test_parser_scenario_gpu(sc, gpu_lib):
    ctx = pkg.Context(0, lib=gpu_lib)
    check(ctx.parser(**sc["parser"]).do, sc)


@pytest.mark.parametrize("sc", SCENARIOS, ids=IDS)
def test_parser_scenario_pins_the_oracle(sc):
    o = util.Oracle()
    p = o.parser(**sc["parser"])
    check(lambda line: o.parser_do(p, line), sc)
