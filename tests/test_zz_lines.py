"""f4 (input side): raw text as log events -- in_tail's line loop (plugins/in_tail/tail_file.c:629-700, :338-391) against the
loop restated in the reference harness around the reference's own log event encoder (oracle/refshim/ref_harness.c:
flbref_lines_to_events), and against committed vectors made from it."""
import base64
import json
import os
import random

import pytest

import util

pkg = util.pkg
PIECES = [b"", b"\r", b"a", b"ab", b"line of text", b"x" * 31, b"y" * 255, b"z" * 256, b"w" * 257, b"crlf line\r", b"\r\r", b"tab\there",
          b"utf8 \xc3\xa9\xe4\xb8\x96", b"\x00\x01bin\xff", b"q" * 5000, b" ", b"ends with cr\r", b"{\"k\": 1}"]


def text(rng, n, tail=True):
    t = b"\n".join(rng.choice(PIECES) for _ in range(n)) + (b"\n" if n else b"")
    if tail and rng.random() < 0.6:
        t += rng.choice([b"unfinished", b"\r", b"x" * 300])
    return t


def _diff(lib, rounds):
    rng = random.Random(41)
    ref = util.Ref()
    ctx = pkg.Context(0, lib=lib)
    for r in range(rounds):
        t = text(rng, rng.choice([0, 1, 2, 5, 40, 700]))
        kw = dict(key=rng.choice(["log", "message", "m" * 40]), skip_empty_lines=rng.random() < 0.5, sec=1700000000 + r, nsec=rng.choice([0, 999999999]))
        if rng.random() < 0.4:
            kw.update(path_key="file", path="/var/log/containers/app-%d.log" % r)
        if rng.random() < 0.4:
            kw.update(offset_key="offset", stream_offset=rng.choice([0, 127, 2 ** 16, 2 ** 32 + 5]))
        assert ctx.lines_to_events(t, **kw) == ref.lines_to_events(t, **kw), (r, kw, t[:200])


def _large(lib):
    """several megabytes: every tile, block and scan boundary crossed; what comes back parses into one event per line"""
    rng = random.Random(42)
    ref = util.Ref()
    ctx = pkg.Context(0, lib=lib)
    t = text(rng, 60000)
    for skip in (True, False):
        got = ctx.lines_to_events(t, "log", skip, 1700000000, 1, offset_key="o", stream_offset=10)
        assert got == ref.lines_to_events(t, "log", skip, 1700000000, 1, offset_key="o", stream_offset=10)
        assert got[1] == t.rfind(b"\n") + 1 and len(util.split_records(got[0])) == got[2]


def _golden(lib):
    vec = json.load(open(os.path.join(util.ROOT, "tests", "golden", "lines_vectors.json")))
    ctx = pkg.Context(0, lib=lib)
    D = lambda s: None if s is None else base64.b64decode(s)
    assert len(vec) >= 20
    for v in vec:
        assert ctx.lines_to_events(D(v["text"]), **v["kw"]) == (D(v["out"]), v["consumed"], v["lines"]), v["kw"]


def test_lines_diff_hostsim(sim_lib, ref_available):
    _diff(sim_lib, 300)


def test_lines_large_hostsim(sim_lib, ref_available):
    _large(sim_lib)


def test_lines_golden_hostsim(sim_lib):
    _golden(sim_lib)


def test_lines_then_chain_hostsim(sim_lib, ref_available):
    """text -> events -> the apache chain -> JSON lines: the three conversions around the filter path, each against the reference"""
    import cases
    ref = util.Ref()
    ctx = pkg.Context(0, lib=sim_lib)
    t = b"\r\n".join(util.apache_lines(200, seed=7)) + b"\r\nrest"
    ev, used, n = ctx.lines_to_events(t, "log", True, 1700000000, 0)
    assert (ev, used, n) == ref.lines_to_events(t, "log", True, 1700000000, 0) and n == 200
    ap = dict(name="apache", format="regex", regex=util.APACHE_RX, time_fmt=util.APACHE_TIME_FMT, time_key="time")
    ctx.parser(**ap); ref.parser(**ap)
    flt = [("parser", [("Key_Name", "log"), ("Parser", "apache")]), ("grep", [("Regex", "method ^(GET|POST)$")])]
    fs = []
    for p, props in flt:
        ref.filter(p, props)
        fs.append(ctx.filter(p, props))
    got = ctx.chain(fs).do(ev)
    assert got == ref.chain_do(ev)
    text_out, und = ctx.to_json(got[1], 3, 1, "date", True)
    assert und == 0 and text_out == ref.to_json(got[1], 3, 1, "date", True)


@pytest.mark.gpu
def test_lines_diff_gpu(gpu_lib, ref_available):
    _diff(gpu_lib, 100)


@pytest.mark.gpu
def test_lines_large_gpu(gpu_lib, ref_available):
    _large(gpu_lib)


@pytest.mark.gpu
def test_lines_golden_gpu(gpu_lib):
    _golden(gpu_lib)
