"""Field decoders of a parser definition -- Decode_Field / Decode_Field_As with the json, escaped, escaped_utf8 and
mysql_quoted backends and the try_next / do_next actions (src/flb_parser_decoder.c:215-535) -- through
flbgpu_parser_do() and through filter_parser chains, byte for byte against the UNMODIFIED reference.  Includes the
reference's own decode_field tests (tests/internal/parser_{json,regex,ltsv}.c: test_decode_field_json)."""
import random

import pytest

import util

pkg = util.pkg

DOCKER = dict(name="docker", format="json", time_key="time", time_fmt="%Y-%m-%dT%H:%M:%S.%L", time_keep=True,
              decoders=[("Decode_Field_As", "escaped_utf8 log do_next"), ("Decode_Field_As", "json log")])
PARSERS = [
    DOCKER,
    dict(name="j_field", format="json", decoders=[("decode_field", "json json_str")]),                        # tests/internal/parser_json.c:396
    dict(name="j_both", format="json", decoders=[("Decode_Field", "json a try_next"), ("Decode_Field_As", "escaped a"), ("Decode_Field", "json b")]),
    dict(name="j_chain", format="json", decoders=[("Decode_Field_As", "escaped log try_next"), ("Decode_Field_As", "mysql_quoted log do_next"),
                                                  ("Decode_Field_As", "json log try_next"), ("Decode_Field", "json log")]),
    dict(name="j_time", format="json", time_key="t", time_fmt="%s", decoders=[("Decode_Field", "json inner")]),     # the time key may come out of a decoded field
    dict(name="j_mysql", format="json", decoders=[("Decode_Field_As", "mysql_quoted q")]),
    dict(name="j_esc", format="json", decoders=[("Decode_Field_As", "escaped e"), ("Decode_Field_As", "escaped_utf8 u")]),
    dict(name="r_field", format="regex", regex=r"^(?<key>[^ ]+) (?<json_str>.*)$", decoders=[("decode_field", "json json_str")]),   # parser_regex.c:396
    dict(name="r_as", format="regex", regex=r"^(?<time>[^ ]+) (?<log>.*)$", time_key="time", time_fmt="%Y-%m-%dT%H:%M:%S",
         decoders=[("Decode_Field_As", "json log try_next"), ("Decode_Field_As", "escaped log")], types="n:integer"),
    dict(name="l_field", format="ltsv", decoders=[("decode_field", "json json_str")]),                         # parser_ltsv.c:392
    dict(name="lf_as", format="logfmt", decoders=[("Decode_Field_As", "escaped_utf8 msg do_next"), ("Decode_Field_As", "json msg")]),
]
INNER = [r'{\"a\":1,\"b\":\"x\"}', r'{\"t\":\"1700000000\",\"deep\":{\"k\":[1,2.5,null,true]}}', r'[1,2]', r'{\"a\":1} trailing', r'{\"a\":1}{\"b\":2}', r'  {\"a\":1}', r'\t{\"a\":1}',
         r'not json', r'', r'{', r'{\"u\":\"caf\\u00e9 \\ud83d\\ude00\"}', r'{\"n\":12345678901234567890,\"f\":1e400}', r"'quoted\'s'", r'\"dq \\n \\0 \\Z \\x\"',
         r'tab\\tnl\\nbs\\\\ q\\\" a\\a v\\v x\\x', r'ends with backslash \\', r'\\', r'\\u00e9\\ud83d\\ude00\\u12\\uZZ', r'\x01ctl', 'h\u00e9llo'.encode().decode("latin1")]


def lines_for(kw, rng):
    out = []
    for inner in INNER:
        for key in ("log", "json_str", "a", "b", "inner", "q", "e", "u", "msg"):
            if rng.random() < 0.35:
                continue
            if kw["format"] == "json":
                out.append(('{"x":1,"%s":"%s","%s":"second %s","t":"1700000001","z":[1]}' % (key, inner, rng.choice(["b", "log", "y"]), inner)).encode("latin1"))
            elif kw["format"] == "regex":
                raw = inner.replace('\\"', '"').replace("\\\\", "\\")
                out.append(("2023-05-06T07:08:09 " + raw).encode("latin1"))
                out.append(("k1 " + raw).encode("latin1"))
            elif kw["format"] == "ltsv":
                raw = inner.replace('\\"', '"').replace("\\\\", "\\").replace("\t", " ")
                out.append(("key:v\t%s:%s\tother:1" % (key, raw)).encode("latin1"))
            else:
                out.append(('ts=1 %s="%s" other=2' % (key, inner)).encode("latin1"))
    out += [b'{"log":123,"json_str":{"a":1}}', b'{"json_str":"{\\"k\\":1}","json_str":"{\\"k2\\":2}"}', b'{"a":"{\\"p\\":1}","b":"{\\"q\\":2}"}', b"", b"{}"]
    return out


def _decoders(lib):
    rng = random.Random(3)
    for kw in PARSERS:
        ctx, ref = pkg.Context(0, lib=lib), util.Ref()
        p, rp = ctx.parser(**kw), ref.parser(**kw)
        lines = lines_for(kw, rng)
        got = p.do_batch(lines)
        for line, g in zip(lines, got):
            w = ref.parser_do(rp, line)
            assert g[0] == w[0], (kw["name"], line, g[0], w[0])
            if w[0] >= 0:
                assert g[1] == w[1], (kw["name"], line, g[1], w[1])
                assert g[2] == w[2], (kw["name"], line)


def test_decoders_hostsim(sim_lib, ref_available):
    _decoders(sim_lib)


@pytest.mark.gpu
def test_decoders_gpu(gpu_lib, ref_available):
    _decoders(gpu_lib)


def _in_chain(lib):
    """the docker parser with its decoders inside filter_parser + grep on a key of the decoded object + modify"""
    rng = random.Random(9)
    app = [l for l in util.json_lines(400, seed=77) if l.startswith(b"{")]
    lines = []
    for i, l in enumerate(app):
        inner = l.decode("latin1").replace("\\", "\\\\").replace('"', '\\"')
        if rng.random() < 0.2:
            inner = "plain text line %d" % i
        lines.append(('{"log":"%s\\n","stream":"%s","time":"2023-05-06T07:08:%02d.%dZ"}' % (inner, rng.choice(["stdout", "stderr"]), i % 60, i)).encode("latin1"))
    chunk = util.chunk_from_lines(lines)
    filters = [("parser", [("Key_Name", "log"), ("Parser", "docker"), ("Reserve_Data", "On")]),
               ("grep", [("Regex", "stream stdout")]),
               ("modify", [("Rename", "log payload"), ("Add", "node n1")])]
    ctx, ref = pkg.Context(0, lib=lib), util.Ref()
    ctx.parser(**DOCKER); ref.parser(**DOCKER)
    for p, props in filters:
        ref.filter(p, props)
    want = ref.chain_do(chunk)
    assert ctx.chain([ctx.filter(p, props) for p, props in filters]).do(chunk) == want
    # and one filter at a time, like flb_filter_do()
    cur = chunk
    for p, props in filters:
        r, out = ctx.filter(p, props).cb(cur)
        if r == pkg.FILTER_MODIFIED:
            cur = out
    assert cur == want[1]


def test_decoders_in_chain_hostsim(sim_lib, ref_available):
    _in_chain(sim_lib)


@pytest.mark.gpu
def test_decoders_in_chain_gpu(gpu_lib, ref_available):
    _in_chain(gpu_lib)


def test_decoder_definitions(sim_lib, ref_available):
    """the definitions the reference refuses are refused (src/flb_parser_decoder.c:638-700)"""
    for dec in ([("Decode_Field", "json")], [("Decode_Field_As", "nosuch log")], [("Decode_Field", "")]):
        ctx = pkg.Context(0, lib=sim_lib)
        with pytest.raises(pkg.FlbGpuError):
            ctx.parser(name="bad", format="json", decoders=dec)
        with pytest.raises(RuntimeError):
            util.Ref().parser(name="bad", format="json", decoders=dec)
