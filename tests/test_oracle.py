"""Pins the plain-C restatement under oracle/ (liboracle.so): it must reproduce the committed golden
vectors (made by the unmodified reference build, tests/golden/make_golden.py) and, when oracle/_ref
is present, the live reference on the seeded parity cases."""
import json
import os

import pytest

import cases
import l2m_cases
import util

G = os.path.join(util.ROOT, "tests", "golden")
TIME = json.load(open(os.path.join(G, "time_vectors.json")))
REGEX = json.load(open(os.path.join(G, "regex_vectors.json")))
CHAIN = json.load(open(os.path.join(G, "chain_vectors.json")))
L2M = json.load(open(os.path.join(G, "l2m_vectors.json")))

# What the restatement does not cover (documented in oracle/orc_parsers.c): yyjson's handling of invalid
# \u escapes / lone surrogates, which only the real reference build defines.
NOT_RESTATED = {"json_parser_edge", "json_parser_edge_plain_reserve"}


def test_time_vectors():
    for i, v in enumerate(TIME):
        o = util.Oracle(now=v.get("now"))
        p = o.parser("t%d" % i, "regex", r"^(?<time>.+)$", time_fmt=v["fmt"], time_key="time", time_offset=v["offset"], time_keep=True)
        r, data, (sec, nsec) = o.parser_do(p, v["str"].encode())
        assert (r >= 0) == (v["ref_ret"] >= 0), v
        if r < 0:
            continue
        assert nsec == v["ref_nsec"], v
        if not v["no_year"]:
            assert sec == v["ref_sec"], v
        assert data.hex() == v["ref_map_hex"], v


def test_regex_vectors():
    o = util.Oracle()
    n = 0
    for pat in REGEX:
        for c in pat["cases"]:
            got = o.regex_search(pat["pattern"], bytes.fromhex(c["s"]))
            want = None if c["m"] is None else [tuple(x) for x in c["m"]]
            assert got == want, (pat["pattern"], c)
            n += 1
    assert n >= 80


@pytest.mark.parametrize("v", CHAIN, ids=[v["name"] for v in CHAIN])
def test_chain_vectors(v):
    if v["name"] in NOT_RESTATED:
        pytest.skip("yyjson quirk cases are covered by oracle/_ref only")
    o = util.Oracle()
    for kw in v["parsers"]:
        o.parser(**kw)
    for p, props in v["filters"]:
        o.filter(p, [tuple(x) for x in props])
    r, out = o.chain_do(bytes.fromhex(v["in_hex"]))
    assert r == v["ret"]
    assert (out.hex() if out is not None else None) == v["out_hex"]


@pytest.mark.parametrize("v", L2M, ids=[v["name"] for v in L2M])
def test_l2m_vectors(v):
    o = util.Oracle()
    for kw in v["parsers"]:
        o.parser(**kw)
    fs = [o.filter(p, [tuple(x) for x in props]) for p, props in v["filters"]]
    r, out = o.chain_do(bytes.fromhex(v["in_hex"]))
    assert r == v["ret"]
    assert (out.hex() if out is not None else None) == v["out_hex"]
    assert o.l2m_text(fs[v["k"]]) == v["text"]


@pytest.mark.parametrize("case", cases.CASES, ids=[c[0] for c in cases.CASES])
def test_live_against_reference(case, ref_available):
    name, parsers, filters, mk = case
    if name in NOT_RESTATED:
        pytest.skip("yyjson quirk cases are covered by oracle/_ref only")
    chunk = mk()
    o, ref = util.Oracle(), util.Ref()
    for kw in parsers:
        o.parser(**kw); ref.parser(**kw)
    for p, props in filters:
        o.filter(p, props); ref.filter(p, props)
    assert o.chain_do(chunk) == ref.chain_do(chunk)
