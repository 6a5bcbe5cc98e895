"""flb_pack_json_state() (src/flb_pack.c:758-829: jsmn tokeniser + tokens_to_msgpack) -- SURVEY 8 row a8 -- against the
UNMODIFIED reference: return value, msgpack bytes, state->last_byte and state->tokens_count, for whole documents, several
documents per buffer, buffers cut at every byte (the streaming contract: what in_tcp sees while a message arrives), the
number rules of pack_numeric_token() and the oddities jsmn's strict mode lets through."""
import json
import os
import random

import pytest

import cases
import util

pkg = util.pkg


def corpus():
    docs = [l for l in util.json_lines(60, seed=91) if l.startswith(b"{")]
    out = list(docs[:20])
    out += cases.JSON_EDGE
    out += [b'{"a":1}{"b":2}', b'{"a":1}\n{"b":[1,2,{"c":null}]}\n', b'{"a":1} {"b":', b'{"a":1}{"b":2}{"c"', b'[1,2,3]', b'[1,2,3] [4', b'"top"', b'"a" "b" 5 ',
            b'{"a" "b"}', b'{"a":"b" "c":"d"}', b'{"a":tru}', b'{"a":nope}', b'{"a":-}', b'{"a":-x1}', b'{"a":1x}', b'{"a":0x10}', b'{"a":1e5x}', b'{"a":1.2.3}',
            b'{"a":.5}', b'{"a":-.5e1}', b'{"a":1E+2}', b'{"a":1e}', b'{"a":12345678901234567890}', b'{"a":18446744073709551615}', b'{"a":18446744073709551616}',
            b'{"a":9223372036854775807}', b'{"a":9223372036854775808}', b'{"a":-9223372036854775808}', b'{"a":-9223372036854775809}', b'{"a":-0}', b'{"a":007}',
            b'{"a":1e400}', b'{"a":-1e400}', b'{"a":4.9e-324}', b'{"a":0.1e-400}', b'{"a":123456789012345678901234567890.5}',
            b'{1:2}', b'{"a":{"b":1}:2}', b'{{"a":1}:2}', b'{[1]:2}', b'{"a":[}', b'{"a":1]', b'}', b']', b'{"a":1}}', b'[[[[[[[[[[1]]]]]]]]]]', b'[' * 200 + b']' * 200,
            b'{"s":"\\u00e9\\u20ac\\ud83d\\ude00 \\n\\t\\"\\\\\\/\\b\\f\\r"}', b'{"s":"\\u12"}', b'{"s":"\\uZZZZ"}', b'{"s":"\\x41"}', b'{"s":"bad \\ud83d x"}', b'{"s":"\\u0000tail"}',
            b'{"s":"ctl\x01\x1f raw\ttab"}', b'{"s":"\xff\xfe \xc3\xa9"}', b'{"s":"it\'s"}', b'{"k\\n":1}', b'{"a":1}\x00{"b":2}', b'{"a":"x\x00y"}', b'\x00', b'',
            b' ', b'\n\n', b'{"a":1,}', b'{,}', b'{"a":1,,"b":2}', b'[1,,2]', b'[,]', b'{"a"::1}', b'{"a":1 "b":2}', b'true', b'true ', b'1', b'1 ', b'null,', b'nul ', b'{"a":true false}',
            b'{"a":[1,2,3],"b":{"c":[{"d":1},{"e":[]}]},"f":""}', b'{"key":"' + b"x" * 300 + b'"}', b'{"k":"' + b"y" * 70000 + b'"}', b'[' + b",".join(b"%d" % i for i in range(20)) + b"]",
            b'[' + b",".join(b'"%d"' % i for i in range(70000)) + b']']
    return out


def check(lib, bufs):
    ctx = pkg.Context(0, lib=lib)
    ref = util.Ref()
    got = ctx.pack_json_state(bufs)
    for b, g in zip(bufs, got):
        w = ref.pack_json_state(b)
        assert g[0] == w[0], (b[:80], g[0], w[0])
        if w[0] == 0:
            assert g[1] == w[1], b[:80]
            assert g[2] == w[2] and g[3] == w[3], (b[:80], g, w[2:])


def _all(lib):
    docs = corpus()
    check(lib, docs)
    # the streaming contract: a message cut at every byte
    stream = b'{"a":1,"s":"x\\ny","n":[1,2.5,-3e2,true,null]} {"b":{"c":"d"}}\n[1,"two",{"3":4}] 17 "s" '
    check(lib, [stream[:k] for k in range(len(stream) + 1)])
    # random concatenations and truncations
    rng = random.Random(5)
    bufs = []
    for _ in range(300):
        parts = [rng.choice(docs[:60]) for _ in range(rng.randint(1, 4))]
        b = rng.choice([b"", b" ", b"\n"]).join(parts)
        if rng.random() < 0.5 and b:
            b = b[:rng.randint(0, len(b))]
        bufs.append(b)
    check(lib, bufs)
    # mutations
    bufs = []
    for _ in range(400):
        b = bytearray(rng.choice(docs[:40]))
        for _ in range(rng.randint(1, 3)):
            if b:
                b[rng.randrange(len(b))] = rng.choice(b'{}[]":,\\ 0-e.tfnu\x00\x7f')
        bufs.append(bytes(b))
    check(lib, bufs)


def test_pack_json_state_hostsim(sim_lib, ref_available):
    _all(sim_lib)


@pytest.mark.gpu
def test_pack_json_state_gpu(gpu_lib, ref_available):
    _all(gpu_lib)


def _equals_yyjson(lib):
    """benchmarks/pack_json.c:17-59 of the reference: on well-formed documents the jsmn path and the yyjson path
    (flb_pack_json, used by the JSON parser) give the same msgpack -- replayed on the reference's own fixtures"""
    vec = json.load(open(os.path.join(util.ROOT, "tests", "golden", "json_scenarios.json")))
    ctx = pkg.Context(0, lib=lib)
    p = ctx.parser(name="plain", format="json")
    lines = [bytes.fromhex(v["line_hex"]) for v in vec if v.get("line_hex")][:200] if isinstance(vec, list) else []
    lines = [l for l in lines if l.strip().startswith(b"{")]
    got = ctx.pack_json_state(lines) if lines else []
    for l, g in zip(lines, got):
        r, data, _ = p.do(l)
        if g[0] == 0 and r >= 0 and b"\\u" not in l and b"." not in l and b"e" not in l.lower():
            assert g[1] == data, l[:80]


def test_jsmn_equals_yyjson_on_fixtures_hostsim(sim_lib):
    _equals_yyjson(sim_lib)
