"""Per-instance back-end state: several filter instances called from different threads at the same time (what
flb_processor_run() does, /root/reference/src/flb_processor.c:1352-1378), the small-chunk form against the sliced
forms, and a slice with more broken candidate links than the one-CTA chain walk holds."""
import os
import threading

import pytest

import cases
import util

pkg = util.pkg


def _ref_result(parsers, filters, chunk):
    ref = util.Ref()
    for kw in parsers:
        ref.parser(**kw)
    for p, props in filters:
        ref.filter(p, props)
    return ref.chain_do(chunk)


def _concurrent(lib, n_threads=6, rounds=8):
    north = [c for c in cases.CASES if c[0] == "north_star_chain"][0]
    jsonc = [c for c in cases.CASES if c[0] == "json_chain_config1"][0]
    work = []
    for t in range(n_threads):
        case = north if t % 2 == 0 else jsonc
        lines = util.apache_lines(1500 + 100 * t, seed=100 + t) if t % 2 == 0 else util.json_lines(1500 + 100 * t, seed=100 + t)
        chunk = util.chunk_from_lines(lines)
        work.append((case, chunk, _ref_result(case[1], case[2], chunk)))
    ctx = pkg.Context(0, lib=lib)
    for kw in north[1] + jsonc[1]:
        ctx.parser(**kw)
    chains = [ctx.chain([ctx.filter(p, props) for p, props in w[0][2]]) for w in work]
    errors = []

    def run(t):
        try:
            for _ in range(rounds):
                got = chains[t].do(work[t][1])
                if got != work[t][2]:
                    errors.append("thread %d: result differs from the reference" % t)
                    return
        except Exception as e:                      # noqa: BLE001
            errors.append("thread %d: %r" % (t, e))

    th = [threading.Thread(target=run, args=(t,)) for t in range(n_threads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors


def test_concurrent_instances_hostsim(sim_lib, ref_available):
    _concurrent(sim_lib, n_threads=3, rounds=2)


@pytest.mark.gpu
def test_concurrent_instances_gpu(gpu_lib, ref_available):
    _concurrent(gpu_lib)


def _forms(lib, monkeypatch):
    """the same chunk through the small-chunk form, the streaming slices and the classic two-pass form"""
    for name, lines in (("north_star_chain", util.apache_lines(9000, seed=51)), ("json_chain_config1", util.json_lines(9000, seed=52)),
                        ("parser_modify_recmod", util.apache_lines(5000, seed=53))):
        case = [c for c in cases.CASES if c[0] == name][0]
        chunk = util.chunk_from_lines(lines)
        want = _ref_result(case[1], case[2], chunk)
        for env in ({"FLBGPU_SMALL_MB": "8"}, {"FLBGPU_SMALL_MB": "0", "FLBGPU_SLICE_MB": "1"}, {"FLBGPU_STREAM": "0"}):
            for k in ("FLBGPU_SMALL_MB", "FLBGPU_SLICE_MB", "FLBGPU_STREAM"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            ctx = pkg.Context(0, lib=lib)
            for kw in case[1]:
                ctx.parser(**kw)
            chain = ctx.chain([ctx.filter(p, props) for p, props in case[2]])
            for cut in (len(chunk), len(chunk) - 29):
                data = chunk[:cut]
                w = want if cut == len(chunk) else _ref_result(case[1], case[2], data)
                assert chain.do(data) == w, (name, env, cut)
                assert chain.do(data) == w, (name, env, cut, "second call")


def test_call_forms_hostsim(sim_lib, ref_available, monkeypatch):
    _forms(sim_lib, monkeypatch)


@pytest.mark.gpu
def test_call_forms_gpu(gpu_lib, ref_available, monkeypatch):
    _forms(gpu_lib, monkeypatch)


def _nested_pairs_chunk(n):
    """every record carries "pair": [1, {"a": 1}] -- bytes that frame as a legacy [ts, body] event, so each record
    costs the record index one false candidate and one broken link"""
    S = util.mp_str
    pair = b"\x92\x01\x81" + S(b"a") + b"\x01"
    return b"".join(util.event(1700000000 + i, i, [(b"log", S(b"GET /x%d" % i)), (b"pair", pair), (b"n", bytes([i % 100]))]) for i in range(n))


def _many_breaks(lib, n):
    chunk = _nested_pairs_chunk(n)
    for filters in ([("grep", [("Regex", "log 7")])], [("modify", [("Add", "env prod")])], [("record_modifier", [("Remove_key", "pair")])]):
        want = _ref_result([], filters, chunk)
        ctx = pkg.Context(0, lib=lib)
        got = ctx.chain([ctx.filter(p, props) for p, props in filters]).do(chunk)
        assert got == want


def test_many_broken_links_hostsim(sim_lib, ref_available):
    _many_breaks(sim_lib, 3000)


@pytest.mark.gpu
def test_many_broken_links_gpu(gpu_lib, ref_available, monkeypatch):
    """more than 8192 broken links in one slice: the chain is decided by pointer doubling (k_link_*)"""
    _many_breaks(gpu_lib, 30000)                          # small-chunk form first, then the classic one
    monkeypatch.setenv("FLBGPU_SMALL_MB", "0")
    _many_breaks(gpu_lib, 30000)                          # streaming slices


@pytest.mark.gpu
def test_two_contexts_one_process_gpu(gpu_lib, ref_available):
    """a second flbgpu_init() (same or another device) has its own queues"""
    n_dev = gpu_lib.flbgpu_device_count()
    case = [c for c in cases.CASES if c[0] == "north_star_chain"][0]
    chunk = util.chunk_from_lines(util.apache_lines(2000, seed=77))
    want = _ref_result(case[1], case[2], chunk)
    ctxs = [pkg.Context(d % n_dev, lib=gpu_lib) for d in range(max(2, min(n_dev, 4)))]
    chains = []
    for ctx in ctxs:
        for kw in case[1]:
            ctx.parser(**kw)
        chains.append(ctx.chain([ctx.filter(p, props) for p, props in case[2]]))
    for _ in range(3):
        for ch in chains:
            assert ch.do(chunk) == want


def _fused_match(lib):
    """Match / Match_Regex / active decide per filter of the FUSED chain whether it sees the chunk, like flb_filter_do()
    (src/flb_filter.c:180-190, flb_router_match)"""
    chunk = cases.apache_chunk(500)
    variants = [
        [("parser", cases.P[1] + [("match", "app.*")]), ("grep", [("Regex", "method ^(GET|POST)$"), ("match", "*.web")]), ("modify", [("Add", "env prod"), ("match", "*")])],
        [("parser", cases.P[1] + [("match", "never"), ("match_regex", "^app\\.(web|db)$")]), ("grep", [("Regex", "log GET"), ("match", "a*b*")]), ("modify", [("Add", "k v"), ("active", "false")])],
        [("grep", [("Exclude", "log POST"), ("match", "sys"), ("match_regex", "^x")]), ("record_modifier", [("Record", "h n1"), ("match", "*s")])],
    ]
    for filters in variants:
        for tag in ("app.web", "app.db", "sys", "axxbyy", "xs", ""):
            ctx = pkg.Context(0, lib=lib)
            ctx.parser(**cases.AP)
            ref = util.Ref()
            ref.parser(**cases.AP)
            for p, props in filters:
                ref.filter(p, props)
            chain = ctx.chain([ctx.filter(p, props) for p, props in filters])
            for _ in range(2):
                assert chain.do(chunk, tag) == ref.chain_do(chunk, tag), (filters, tag)


def test_fused_match_routing_hostsim(sim_lib, ref_available):
    _fused_match(sim_lib)


@pytest.mark.gpu
def test_fused_match_routing_gpu(gpu_lib, ref_available):
    _fused_match(gpu_lib)


def _ascii_only_patterns(lib):
    """POSIX brackets, \\b / \\B and case-insensitive matching consult Onigmo's Unicode tables for non-ASCII subjects; those are
    not restated, so such a pattern meeting a non-ASCII value fails the call loudly instead of matching approximately"""
    ascii_chunk = util.chunk_from_lines([b"GET /a cafe", b"post /b CAFE", b"x"])
    uni_chunk = util.chunk_from_lines([b"GET /a caf\xc3\xa9", b"post /b CAF\xc3\x89"])
    for rule in ("log /caf/i", "log \\bGET\\b", "log [[:alpha:]]+ /", "log (?i)post"):
        filters = [("grep", [("Regex", rule)])]
        ctx = pkg.Context(0, lib=lib)
        chain = ctx.chain([ctx.filter(p, props) for p, props in filters])
        assert chain.do(ascii_chunk) == _ref_result([], filters, ascii_chunk), rule
        with pytest.raises(pkg.FlbGpuError, match="non-ASCII"):
            chain.do(uni_chunk)
    # a pattern without such constructs takes non-ASCII subjects as before
    filters = [("grep", [("Regex", "log caf.$")])]
    ctx = pkg.Context(0, lib=lib)
    assert ctx.chain([ctx.filter(p, props) for p, props in filters]).do(uni_chunk) == _ref_result([], filters, uni_chunk)


def test_ascii_only_patterns_hostsim(sim_lib, ref_available):
    _ascii_only_patterns(sim_lib)


@pytest.mark.gpu
def test_ascii_only_patterns_gpu(gpu_lib, ref_available):
    _ascii_only_patterns(gpu_lib)


def _two_stage(lib, monkeypatch):
    lines = util.json_lines(20000, seed=61) + [c for c in cases.JSON_EDGE] * 20
    chunk = util.chunk_from_lines(lines)
    for name in ("json_chain_config1",):
        case = [c for c in cases.CASES if c[0] == name][0]
        want = _ref_result(case[1], case[2], chunk)
        for env in ({}, {"FLBGPU_SMALL_MB": "0", "FLBGPU_SLICE_MB": "1"}, {"FLBGPU_EVAL_SPLIT": "0"}):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            ctx = pkg.Context(0, lib=lib)
            for kw in case[1]:
                ctx.parser(**kw)
            got = ctx.chain([ctx.filter(p, props) for p, props in case[2]]).do(chunk)
            assert got[0] == want[0] and len(got[1]) == len(want[1]), (env, got[0], want[0], len(got[1]), len(want[1]))
            if got[1] != want[1]:
                at = next(i for i in range(len(want[1])) if got[1][i] != want[1][i])
                raise AssertionError((env, "first difference at byte", at, got[1][max(0, at - 60):at + 40], want[1][max(0, at - 60):at + 40]))
    # parser only, with Reserve_Data and a record accessor key: every shape the walker meets
    filters = [("parser", [("Key_Name", "log"), ("Parser", "json"), ("Reserve_Data", "On")])]
    ctx = pkg.Context(0, lib=lib)
    ctx.parser(**cases.JS)
    assert ctx.chain([ctx.filter(p, props) for p, props in filters]).do(chunk) == _ref_result([cases.JS], filters, chunk)


def test_two_stage_json_tokenizer_sim(sim_lib, ref_available, monkeypatch):
    """FLBGPU_JSON_BM=1 as the emulation runs it: stage-1 bitmap, bit-scan walker, follow-up pass over the records put off"""
    monkeypatch.setenv("FLBGPU_JSON_BM", "1")
    _two_stage(sim_lib, monkeypatch)


@pytest.mark.gpu
def test_two_stage_json_tokenizer_gpu(gpu_lib, ref_available, monkeypatch):
    """FLBGPU_JSON_BM=1: the warp-cooperative stage-1 bitmap, the bit-scan walker and the follow-up launch over the records it
    puts off give the bytes of the default byte scanner (= the reference's)"""
    monkeypatch.setenv("FLBGPU_JSON_BM", "1")
    _two_stage(gpu_lib, monkeypatch)

