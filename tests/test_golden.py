"""Committed golden vectors (tests/golden/*.json, produced from the UNMODIFIED reference by
tests/golden/make_golden.py).  These run without /root/reference and without oracle/_ref, on
the CPU emulation of the device code here and on the real GPU on the box."""
import ctypes as C
import json
import os

import pytest

import util

pkg = util.pkg
G = os.path.join(util.ROOT, "tests", "golden")
TIME = json.load(open(os.path.join(G, "time_vectors.json")))
REGEX = json.load(open(os.path.join(G, "regex_vectors.json")))
CHAIN = json.load(open(os.path.join(G, "chain_vectors.json")))


def test_reference_table_pins_the_oracle():
    """The reference's own expected epochs (tests/internal/parser.c:61-105) equal what the
    reference build under oracle/_ref produced when the fixtures were generated."""
    n = 0
    for v in TIME:
        if v["table_epoch"] is not None and not v["no_year"]:
            assert v["ref_ret"] >= 0
            assert v["ref_sec"] == v["table_epoch"], v
            assert v["ref_nsec"] == int(v["table_frac"] * 1000000000.0), v
            n += 1
    assert n >= 20


def check_time(lib):
    ctx = pkg.Context(0, lib=lib)
    for i, v in enumerate(TIME):
        p = ctx.parser("t%d" % i, "regex", r"^(?<time>.+)$", time_fmt=v["fmt"], time_key="time",
                       time_offset=v["offset"], time_keep=True)
        r, data, (sec, nsec) = p.do(v["str"].encode())
        assert (r >= 0) == (v["ref_ret"] >= 0), v
        if r < 0:
            continue
        assert nsec == v["ref_nsec"], v
        if not v["no_year"]:
            assert sec == v["ref_sec"] & 0xffffffff, v
        assert data.hex() == v["ref_map_hex"], v


def test_time_vectors_hostsim(sim_lib):
    check_time(sim_lib)


@pytest.mark.gpu
def test_time_vectors_gpu(gpu_lib):
    check_time(gpu_lib)


def test_regex_vectors_hostsim():
    import rxdiff_sim
    for entry in REGEX:
        h = rxdiff_sim.compile(entry["pattern"])
        assert h, entry["pattern"]
        for c in entry["cases"]:
            got = rxdiff_sim.search(h, bytes.fromhex(c["s"]))
            want = None if c["m"] is None else tuple(tuple(x) for x in c["m"])
            if want == ("m",) or (want is not None and len(want) == 1 and want[0] == ("m",)):
                assert got is not None
            else:
                assert got == want, (entry["pattern"], c)


def check_chain(lib):
    for v in CHAIN:
        ctx = pkg.Context(0, lib=lib)
        for kw in v["parsers"]:
            ctx.parser(**kw)
        fs = [ctx.filter(p, [tuple(x) for x in props]) for p, props in v["filters"]]
        ret, out = ctx.chain(fs).do(bytes.fromhex(v["in_hex"]))
        assert ret == v["ret"], v["name"]
        assert (out.hex() if out is not None else None) == v["out_hex"], v["name"]


def test_chain_vectors_hostsim(sim_lib):
    check_chain(sim_lib)


@pytest.mark.gpu
def test_chain_vectors_gpu(gpu_lib):
    check_chain(gpu_lib)


def check_parser_batch(lib):
    """flbgpu_parser_do_batch: every line's (ret, map, time) equals flb_parser_do of the reference, for all
    four formats, including lines that do not parse."""
    import cases
    ref = util.Ref()
    ctx = pkg.Context(0, lib=lib)
    sets = [(cases.AP, util.apache_lines(300, seed=5) + [b"", b"garbage line"]),
            (dict(name="js", format="json", time_key="t", time_fmt="%s"), util.json_lines(300, seed=6) + [b"[1]", b'{"t":"1700000000","a":1}']),
            (dict(name="lt", format="ltsv"), util.ltsv_lines(100) + [b"nolabel", b"a:1\tb:2"]),
            (dict(name="lf", format="logfmt"), util.logfmt_lines(100) + [b"  ", b"k=v bare"])]
    for kw, lines in sets:
        rp = ref.parser(**kw)
        got = ctx.parser(**kw).do_batch(lines)
        assert len(got) == len(lines)
        for line, (r, data, (sec, nsec)) in zip(lines, got):
            rr, rdata, (rsec, rnsec) = ref.parser_do(rp, line)
            assert r == rr, line                       # the position flb_parser_do() returns, or -1
            if r >= 0:
                assert data == rdata, line
                assert (sec, nsec) == (rsec, rnsec), line


def test_parser_batch_hostsim(sim_lib, ref_available):
    check_parser_batch(sim_lib)


@pytest.mark.gpu
def test_parser_batch_gpu(gpu_lib, ref_available):
    check_parser_batch(gpu_lib)


def check_apache_time_fast_path(lib):
    """the direct path for "%d/%b/%Y:%H:%M:%S %z" agrees with the reference on canonical and odd values"""
    ref = util.Ref()
    ctx = pkg.Context(0, lib=lib)
    kw = dict(name="t", format="regex", regex=r"^(?<time>.+)$", time_fmt="%d/%b/%Y:%H:%M:%S %z", time_key="time", time_keep=True)
    vals = ["10/Oct/2000:13:55:36 -0700", "01/jan/1970:00:00:00 +0000", "31/DEC/9999:23:59:60 +1400", "29/Feb/2023:12:00:00 +0530",
            "00/Jan/2023:00:00:00 +0000", "32/Jan/2023:00:00:00 +0000", "1/Jan/2023:00:00:00 +0000", "01/Jan/2023:24:00:00 +0000",
            "01/Jan/2023:00:60:00 +0000", "01/Jan/2023:00:00:61 +0000", "01/Foo/2023:00:00:00 +0000", "01/Jan/2023:00:00:00 +00:00",
            "01/Jan/2023:00:00:00 Z", "01/Jan/2023:00:00:00 +0", "01/Jan/2023:00:00:00 -0099", "01/Jan/0000:00:00:00 +0000",
            "01/March/2023:00:00:00 +0000", "01/Jan/2023:00:00:00  +0000", "01/Jan/2023:00:00:00 +000a", " 1/Jan/2023:00:00:00 +0000",
            "31/Apr/2024:07:08:09 -1234", "15/Sep/2038:03:14:08 +0000"]
    # out-of-range fields: strptime stops there and, when time_strict is off, the half-filled tm is what counts
    vals += ["28/Jul/2024:12:60:13 -0700", "01/May/1970:12:03:61 +0530", "01/Mar/2000:24:59:61 -1200", "31/Nov/0068:05:59:61 -0700", "09/May/2023:00:60:60 +9999"]
    rp = ref.parser(**kw)
    got = ctx.parser(**kw).do_batch([v.encode() for v in vals])
    kw2 = dict(kw, name="t2", time_strict=False)
    rp2 = ref.parser(**kw2)
    for v, (r, data, (sec, nsec)) in zip(vals, ctx.parser(**kw2).do_batch([v.encode() for v in vals])):
        rr, rdata, (rsec, rnsec) = ref.parser_do(rp2, v.encode())
        assert (r >= 0) == (rr >= 0) and data == rdata, v
        if rr >= 0:
            assert (sec, nsec) == (rsec, rnsec), v
    for v, (r, data, (sec, nsec)) in zip(vals, got):
        rr, rdata, (rsec, rnsec) = ref.parser_do(rp, v.encode())
        assert (r >= 0) == (rr >= 0), v
        assert data == rdata, v
        if r >= 0:
            assert (sec, nsec) == (rsec, rnsec), v


def test_apache_time_fast_path_hostsim(sim_lib, ref_available):
    check_apache_time_fast_path(sim_lib)


@pytest.mark.gpu
def test_apache_time_fast_path_gpu(gpu_lib, ref_available):
    check_apache_time_fast_path(gpu_lib)


def check_time_programs(lib):
    """formats with a compiled fixed-shape program (runtime.c:time_fast_compile) against the reference, on
    canonical values and on values that must fall back to the general path"""
    import random
    rng = random.Random(77)
    fmts = ["%Y-%m-%dT%H:%M:%S.%L%z", "%Y-%m-%dT%H:%M:%S.%LZ", "%Y-%m-%dT%H:%M:%S.%L", "%Y-%m-%d %H:%M:%S", "%Y-%m-%dT%T%z",
            "%d/%b/%Y:%H:%M:%S %z", "%Y/%m/%d %H:%M:%S.%L %z", "%b %d %Y %H:%M:%S", "%Y%m%d%H%M%S", "%H:%M:%S %d-%m-%Y", "%z %Y-%m-%d %H:%M:%S.%L"]
    base = ["2023-07-14T09:08:07.123456789+02:00", "2023-07-14T09:08:07.5Z", "2023-07-14T09:08:07.000001", "2023-07-14 09:08:07",
            "2023-07-14T23:59:60-0330", "14/Jul/2023:09:08:07 +0000", "2023/07/14 09:08:07.25 +0900", "Jul 14 2023 09:08:07",
            "20230714090807", "09:08:07 14-07-2023", "+0100 2023-07-14 09:08:07.75", "2023-7-14T9:8:7.1Z", "2023-07-14T09:08:07.1234567891Z",
            "2023-07-14T09:08:07.+02:00", "2023-07-14T09:08:07.12 +02:00", "2023-13-14T09:08:07.1Z", "2023-07-32T09:08:07.1Z",
            "2023-07-14T24:08:07.1Z", "2023-07-14T09:08:07.1z", "2023-07-14T09:08:07.1+2", "2023-07-14T09:08:07.1+02", "2023-07-14T09:08:07.1+02:",
            "2023-07-14T09:08:07.1+02:3", "2023-07-14T09:08:07.1GMT", "July 14 2023 09:08:07", "jul 14 2023 09:08:07", "Jul  14 2023 09:08:07",
            "2023-07-14  09:08:07", "2023-07-14 09:08:07 trailing", "0000-01-01 00:00:00", "9999-12-31 23:59:60", ""]
    for fi, fmt in enumerate(fmts):
        ref = util.Ref()
        ctx = pkg.Context(0, lib=lib)
        kw = dict(name="t%d" % fi, format="regex", regex=r"^(?<time>.*)$", time_fmt=fmt, time_key="time", time_keep=True, skip_empty=False)
        rp = ref.parser(**kw)
        vals = list(base)
        for _ in range(20):                        # mutations of the canonical values
            v = rng.choice(base[:11])
            if v:
                i = rng.randrange(len(v))
                v = v[:i] + rng.choice("0159:-+.TZ /a") + v[i + 1:]
            vals.append(v)
        got = ctx.parser(**kw).do_batch([v.encode() for v in vals])
        for v, (r, data, (sec, nsec)) in zip(vals, got):
            rr, rdata, (rsec, rnsec) = ref.parser_do(rp, v.encode())
            assert (r >= 0) == (rr >= 0), (fmt, v)
            assert data == rdata, (fmt, v)
            if r >= 0:
                assert (sec, nsec) == (rsec, rnsec), (fmt, v)


def test_time_programs_hostsim(sim_lib, ref_available):
    check_time_programs(sim_lib)


@pytest.mark.gpu
def test_time_programs_gpu(gpu_lib, ref_available):
    check_time_programs(gpu_lib)


def check_parser_do_sign_and_bytes(lib):
    """flbgpu_parser_do() / _do_batch() against flb_parser_do() over odd lines for every parser kind: the same
    value comes back -- -1, or the position inside the line the parser consumed it up to (end of the last named capture,
    end of the JSON document plus white space, where the LTSV / logfmt scan stopped) -- with the reference's map and time
    (64-bit seconds: times before 1970 and after 2106 included)."""
    import cases
    tf = "%Y-%m-%dT%H:%M:%S.%LZ"
    parsers = [cases.AP, cases.JS, cases.LF, cases.LT, dict(name="jst", format="json", time_key="time", time_fmt=tf),
               dict(name="jsk", format="json", time_key="time", time_fmt=tf, time_keep=True, time_strict=False),
               dict(name="lft", format="logfmt", time_key="time", time_fmt=tf, types="n:integer"), dict(name="lfb", format="logfmt", logfmt_no_bare_keys=True),
               dict(name="ltt", format="ltsv", time_key="time", time_fmt=tf, time_keep=True, types="n:hex"), dict(name="jse", format="json", skip_empty=False),
               dict(name="opt", format="regex", regex=r"^(?<a>x)?y(?<b>.*)$"), dict(name="none", format="regex", regex=r"^(?<a>x)?(?<b>q)?y"),
               dict(name="mid", format="regex", regex=r"(?<b>y+)(?<a>x)?"), dict(name="same", format="regex", regex=r"^(?<_>.*)$"),
               dict(name="old", format="regex", regex=r"^(?<time>[^ ]+) (?<m>.*)$", time_key="time", time_fmt="%Y-%m-%dT%H:%M:%S"),
               dict(name="jsold", format="json", time_key="time", time_fmt="%Y-%m-%dT%H:%M:%S")]
    lines = (util.apache_lines(20, seed=1) + util.json_lines(20, 2) + util.logfmt_lines(20, 3) + util.ltsv_lines(20) +
             [b"", b" ", b"\t", b"{}", b"{", b"}", b"[]", b"null", b'""', b"a=", b"=", b":", b"a:", b"\xff\xfe", b'{"a":"\\ud800"}', b'{"a":1e999}', b'{"a":-0}',
              b'{"a":1E+2}', b'{"a":12345678901234567890}', b'{"a":-9223372036854775809}', b' {"a":1} ', b'{"a":1}{"b":2}', b'{"a":1},', b'{"a":1} trailing',
              b'{"a":1} 5', b'{"time":"2023-05-06T07:08:09.123Z","n":1}', b'{"time":"nonsense"}', b"time=2023-05-06T07:08:09.5Z n=7 bare", b"time=bad n=x",
              b"time:2023-05-06T07:08:09.250Z\tn:1f\tempty:", b"a:1\n\tb:2", b"a=1\nb=2", b'a="x', b"y", b"xy tail", b"zzz", b"a\x01:1",
              b"1901-02-03T04:05:06 before the epoch", b"2200-01-02T03:04:05 after 2106", b'{"time":"1901-02-03T04:05:06","n":1}', b'{"time":"2200-01-02T03:04:05"}  ',
              b"_:x", b"_=x", b"ayyyxx", b'{"a":1}\n\n', b'\n {"a":[1,2]} \t x'])
    for kw in parsers:
        ctx, ref = pkg.Context(0, lib=lib), util.Ref()
        p, rp = ctx.parser(**kw), ref.parser(**kw)
        batch = p.do_batch(lines)
        for line, b in zip(lines, batch):
            r, data, t = p.do(line)
            rr, rdata, rt = ref.parser_do(rp, line)
            assert (r, data, t) == b, (kw["name"], line)
            assert r == rr, (kw["name"], line, r, rr)
            if rr >= 0:
                assert data == rdata and t == rt, (kw["name"], line)


def test_parser_do_sign_and_bytes_hostsim(sim_lib, ref_available):
    check_parser_do_sign_and_bytes(sim_lib)


@pytest.mark.gpu
def test_parser_do_sign_and_bytes_gpu(gpu_lib, ref_available):
    check_parser_do_sign_and_bytes(gpu_lib)
