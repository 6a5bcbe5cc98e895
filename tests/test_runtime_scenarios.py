"""The reference's own runtime tests (tests/runtime/filter_modify.c, filter_record_modifier.c,
filter_grep.c, filter_parser.c, filter_log_to_metrics.c), scenario by scenario: same filter properties, same pushed events, and the test's
own assertion -- the substring it looks for in the JSON output, the number of records that come
out, or that the configuration is refused -- plus byte equality with what the unmodified reference
produced for the same chunk (tests/golden/runtime_scenarios.json, made by
tests/golden/make_runtime_scenarios.py).  On the CPU emulation of the device code and on the GPU."""
import hashlib
import json
import os

import pytest

import scenario_util
import util

pkg = util.pkg
SCENARIOS = json.load(open(os.path.join(util.ROOT, "tests", "golden", "runtime_scenarios.json")))
IDS = ["%s:%s" % (s["filter"], s["test"].replace("flb_test_", "").replace("flb_", "")) for s in SCENARIOS]


def run_scenario(lib, sc):
    ctx = pkg.Context(0, lib=lib)
    props = [tuple(p) for p in sc["props"]]
    if sc["init_error"]:                                     # flb_start() fails in the reference test
        with pytest.raises(pkg.FlbGpuError):
            ctx.filter(sc["filter"], props)
        return
    for kw in sc.get("parsers") or []:
        ctx.parser(**kw)
    chunk = scenario_util.chunk_from_json_events(scenario_util.scenario_events(sc))
    f = ctx.filter(sc["filter"], props)
    ret, out = f.cb(chunk)
    # byte equality with the reference
    assert ret == sc["ret"]
    assert (None if out is None else len(out)) == sc["out_len"]
    if sc["out_sha256"]:
        assert hashlib.sha256(out).hexdigest() == sc["out_sha256"]
    else:
        assert (None if out is None else out.hex()) == sc["out_hex"]
    if sc["filter"] == "log_to_metrics":                     # the metric table: the reference's, and the test's assertion on it
        text = f.l2m_text()
        assert text == sc["text"]
        for want in sc["present"]:
            assert want in text, (want, text)
        return
    # what the reference test itself asserts
    result = chunk if ret == 2 else (out or b"")             # FLB_FILTER_NOTOUCH: the engine keeps the input
    texts = scenario_util.records_as_lib_lines(result)
    joined = "\n".join(texts)
    for want in sc["present"]:
        assert want in joined, (want, joined[:300])
    for unwanted in sc["absent"]:
        assert unwanted not in joined, (unwanted, joined[:300])
    if sc["count"] is not None:
        assert len(texts) == sc["count"]


def test_scenarios_cover_the_reference_files():
    by = {}
    for s in SCENARIOS:
        by[s["filter"]] = by.get(s["filter"], 0) + 1
    assert by == {"modify": 36, "record_modifier": 6, "grep": 13, "parser": 14, "log_to_metrics": 8}
    assert sum(bool(s["present"] or s["absent"]) for s in SCENARIOS) >= 60
    assert sum(s["count"] is not None for s in SCENARIOS) >= 8
    assert sum(s["init_error"] for s in SCENARIOS) >= 6


@pytest.mark.parametrize("sc", SCENARIOS, ids=IDS)
def test_runtime_scenario_hostsim(sc, sim_lib):
    run_scenario(sim_lib, sc)


@pytest.mark.gpu
@pytest.mark.parametrize("sc", SCENARIOS, ids=IDS)
def test_runtime_scenario_gpu(sc, gpu_lib):
    run_scenario(gpu_lib, sc)


@pytest.mark.parametrize("sc", SCENARIOS, ids=IDS)
def test_runtime_scenario_pins_the_oracle(sc):
    """the plain-C restatement (oracle/liboracle.so) against the same scenarios: it is what smoke() and the
    GPU-box parity tests lean on, so it has to agree with the reference's own tests too"""
    o = util.Oracle()
    props = [tuple(p) for p in sc["props"]]
    if sc["init_error"]:
        with pytest.raises(RuntimeError):
            o.filter(sc["filter"], props)
        return
    for kw in sc.get("parsers") or []:
        o.parser(**kw)
    f = o.filter(sc["filter"], props)
    ret, out = o.chain_do(scenario_util.chunk_from_json_events(scenario_util.scenario_events(sc)))
    assert ret == sc["ret"]
    assert (None if out is None else len(out)) == sc["out_len"]
    if sc["out_sha256"]:
        assert hashlib.sha256(out).hexdigest() == sc["out_sha256"]
    else:
        assert (None if out is None else out.hex()) == sc["out_hex"]
    if sc["filter"] == "log_to_metrics":
        assert o.l2m_text(f) == sc["text"]
