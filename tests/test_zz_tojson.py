"""f4 (output side): a chunk as JSON text -- flb_pack_msgpack_to_json_format() (src/flb_pack.c:1320-1602) against the reference's
own function: the three layouts, the five date formats, both string writers, numbers, duplicate keys, nested values."""
import random
import struct

import pytest

import util

pkg = util.pkg
S = util.mp_str


def mp_f64(x):
    return b"\xcb" + struct.pack(">d", x)


def mp_int(v):
    if 0 <= v < 128:
        return bytes([v])
    if -32 <= v < 0:
        return bytes([v & 0xff])
    if 0 <= v < 2 ** 32:
        return b"\xce" + struct.pack(">I", v)
    if v >= 0:
        return b"\xcf" + struct.pack(">Q", v)
    return b"\xd3" + struct.pack(">q", v)


TEXT = [b"plain", b"", b"with \"quotes\" and \\ backslash", b"tab\there", b"nl\nhere\r", b"ctl\x01\x02\x1f\x7f", b"\x7f" * 3,
        b"0123456789abcde\x7f" + b"x" * 16, b"x" * 15 + b"\x7f" + b"y" * 16 + b"\x01", "héllo wörld".encode(), "日本語".encode(),
        "\U0001f600 smile".encode(), b"bad \xff\xfe bytes", b"trunc \xe6\x97", b"\xc3(", b"\xed\xa0\x80 surrogate", b"\xc0\x80 overlong",
        b"\xf4\x90\x80\x80 too big", b"\xf8\x88\x80\x80\x80 five", b"a" * 40, b"slash / and ' and <>&", b"\x08\x0c", b"\x00zero"]
REALS = [0.0, -0.0, 1.0, -1.5, 0.1, 3.141592653589793, 1e15, 1e16, 1.5e16, 123456789012345678.0, 1e-4, 1e-5, 1.2345e-7, 5e-324, 1.7976931348623157e308,
         2.0 ** 62, 2.0 ** 63, -(2.0 ** 63), 2.0 ** 64, 9007199254740993.0, 0.30000000000000004, 2.5, 1e21, 1e22, 123.456, float("inf"), float("-inf"),
         1700000000.123457, 4.35, 0.000123456789012345678, 99999999999999990.0, 9999999999999999.0, 0.5, 1e100, 1.0000000000000002]


def value(rng, depth=0):
    r = rng.random()
    if r < 0.30:
        return S(rng.choice(TEXT))
    if r < 0.40:
        return mp_int(rng.choice([0, 1, -1, 127, 128, 255, 65535, 2 ** 31, 2 ** 32, 2 ** 63, 2 ** 64 - 1, -32, -33, -2 ** 31, -2 ** 63, 42]))
    if r < 0.55:
        x = rng.choice(REALS) if rng.random() < 0.6 else struct.unpack(">d", struct.pack(">Q", rng.getrandbits(64)))[0]
        return mp_f64(x)
    if r < 0.60:
        return b"\xca" + struct.pack(">f", rng.choice([1.0, 0.1, 3.5, 1e-3, 16777217.0, -2.75]))
    if r < 0.68:
        return rng.choice([b"\xc0", b"\xc2", b"\xc3"])
    if r < 0.73:
        body = rng.choice(TEXT)[:20]
        return b"\xc4" + bytes([len(body)]) + body                            # bin
    if r < 0.77:
        return b"\xd6\x05\x01\x80\xff\x7f" if rng.random() < 0.5 else b"\xc7\x03\x02abc"     # ext
    if depth < 3 and r < 0.87:
        n = rng.randint(0, 3)
        return bytes([0x90 | n]) + b"".join(value(rng, depth + 1) for _ in range(n))
    if depth < 3:
        n = rng.randint(0, 3)
        keys = [rng.choice([b"k", b"k2", b"dup", "clé".encode()]) for _ in range(n)]
        return bytes([0x80 | n]) + b"".join(S(k) + value(rng, depth + 1) for k in keys)
    return S(b"deep")


def chunk(rng, n, with_meta=True):
    evs = []
    for i in range(n):
        nk = rng.randint(0, 6)
        fields = []
        for _ in range(nk):
            k = rng.choice([b"log", b"level", b"date", b"dup", b"dup", b"n", "clé".encode(), b"__internal__", b"q\"k"])
            fields.append((k, value(rng)))
        if rng.random() < 0.1:
            fields.append((b"\x05", b"\x07"))                                # an integer key (the caller packs key bytes itself below)
        body = util.mp_map_hdr(len(fields)) + b"".join((k if k == b"\x05" else S(k)) + v for k, v in fields)
        meta = b"\x80"
        if with_meta and rng.random() < 0.15:
            meta = b"\x82" + S(b"src") + S(b"tail") + S(b"n") + mp_int(rng.randint(0, 9))
        sec = rng.choice([0, 1, 1700000000, 1700000000 + i, 2 ** 31 - 1, 951782400, 1709164800])
        nsec = rng.choice([0, 1, 999, 1000, 123456789, 999999999, 500000000])
        if rng.random() < 0.08:                                                                      # halves the decoder reads as negative, nanoseconds past a second
            sec = rng.choice([2 ** 31 + 5, 2 ** 32 - 10, 2 ** 32 - 86400 * 400, sec])
            nsec = rng.choice([2 ** 31 + 1, 2 ** 32 - 16, 10 ** 9 + 5, 4 * 10 ** 9, nsec])
        r = rng.random()
        if r < 0.03:
            evs.append(b"\x92\x92\xd7\x00\xff\xff\xff\xfe\x00\x00\x00\x00\x80\x80")          # a group end marker (the decoder steps over it)
        elif r < 0.07:                                                                               # a group start marker with attributes
            evs.append(b"\x92\x92\xd7\x00\xff\xff\xff\xff\x00\x00\x00\x00\x80" + rng.choice([b"\x80", b"\x81" + S(b"res") + value(rng, 2)]))
        if r < 0.1:
            evs.append(b"\x92\xce" + struct.pack(">I", sec) + body)                              # legacy [ts, body]
        elif r < 0.15:
            evs.append(b"\x92\x92\xcb" + struct.pack(">d", rng.choice([sec + 0.25, sec + 0.999999999, -(sec % 2 ** 31) - 0.75, float(sec)])) + meta + body)            # float timestamp
        else:
            evs.append(b"\x92\x92\xd7\x00" + struct.pack(">II", sec, nsec) + meta + body)
    return b"".join(evs)


def _diff(lib, rounds, n):
    rng = random.Random(21)
    ref = util.Ref()
    ctx = pkg.Context(0, lib=lib)
    checked = undefined = 0
    for _ in range(rounds):
        c = chunk(rng, rng.choice([0, 1, 3, n]))
        for jf in (1, 2, 3):
            df = rng.randint(0, 4)
            key = rng.choice(["date", "@timestamp", None, "dup"])
            esc = rng.random() < 0.5
            got, und = ctx.to_json(c, jf, df, key, esc)
            if und:                       # the reference's text depends on memory behind the event here: nothing to compare with
                undefined += 1
                continue
            want = ref.to_json(c, jf, df, key, esc)
            assert got == want, (jf, df, key, esc, c)
            checked += 1
    assert checked > rounds


def _numbers(lib):
    """every real of the table and a few thousand random bit patterns, one event each: "%.1f" / "%.16g" as glibc prints them"""
    rng = random.Random(22)
    ref = util.Ref()
    ctx = pkg.Context(0, lib=lib)
    xs = list(REALS) + [struct.unpack(">d", struct.pack(">Q", rng.getrandbits(64)))[0] for _ in range(3000)]
    xs += [rng.random() * 10 ** rng.randint(-30, 30) for _ in range(2000)] + [float(rng.randint(-10 ** 18, 10 ** 18)) for _ in range(500)]
    xs += [round(rng.random() * 1000, rng.randint(0, 6)) for _ in range(1500)]
    c = b"".join(util.event(1700000000, 0, [(b"x", mp_f64(x))]) for x in xs)
    assert ctx.to_json(c, 3, 2, "t", True)[0] == ref.to_json(c, 3, 2, "t", True)


def _strings(lib):
    """both writers over every text of the table at every offset of a 16-byte window, a plain key behind it (so that what the
    reference reads past the string is the next key's header: defined)"""
    ref = util.Ref()
    ctx = pkg.Context(0, lib=lib)
    evs = []
    for t in TEXT:
        for pad in range(0, 34, 3):
            evs.append(util.event(1700000000, 0, [(b"s", S(b"p" * pad + t + b"q" * (pad % 5))), (b"tail", S(b"end of the event, long enough"))]))
    c = b"".join(evs)
    for esc in (True, False):
        got, und = ctx.to_json(c, 3, 1, "ts", esc)
        assert und == 0
        assert got == ref.to_json(c, 3, 1, "ts", esc)


def _edges(lib):
    ref = util.Ref()
    ctx = pkg.Context(0, lib=lib)
    assert ctx.to_json(b"", 3)[0] is None and ref.to_json(b"", 3) is None
    assert ctx.to_json(b"", 1)[0] == ref.to_json(b"", 1) == b"[]"
    only_skipped = b"\x92\x92\xd7\x00\xff\xff\xff\xfe\x00\x00\x00\x00\x80\x80"
    assert ctx.to_json(only_skipped, 2)[0] is None and ref.to_json(only_skipped, 2) is None
    e = util.event(1700000000, 5, [(b"a", b"\x01")])
    for jf in (1, 2, 3):
        for df in range(5):
            assert ctx.to_json(e * 3, jf, df, "d", True)[0] == ref.to_json(e * 3, jf, df, "d", True)
    # garbage behind the events: what decodes is converted
    assert ctx.to_json(e + e + b"\xc1\xc1", 3)[0] == ref.to_json(e + e + b"\xc1\xc1", 3)
    # group markers: the events between a start (seconds -1) and an end (-2) carry the start's body as group_attributes
    def marker(sec, body, meta=b"\x80"):
        return b"\x92\x92\xd7\x00" + struct.pack(">iI", sec, 0) + meta + body
    attrs = b"\x82" + S(b"resource") + b"\x81" + S(b"svc") + S(b"api") + S(b"n") + b"\x05"
    with_meta = b"\x92\x92\xd7\x00" + struct.pack(">II", 1700000001, 7) + b"\x81" + S(b"m") + b"\x01" + b"\x81" + S(b"a") + b"\x02"
    grouped = (e + marker(-1, attrs, b"\x81" + S(b"schema") + S(b"otlp")) + e + with_meta + marker(-2, b"\x80") + e + with_meta +
               marker(-1, b"\x80") + e + with_meta + marker(-1, attrs) + e + marker(-3, attrs) + e)
    for jf in (1, 2, 3):
        for esc in (True, False):
            got, und = ctx.to_json(grouped, jf, 1, "date", esc)
            assert und == 0 and got == ref.to_json(grouped, jf, 1, "date", esc)
    # timestamps spelled as float64 (out of range and NaN convert as the x86-64 build converts them) and ext timestamps with
    # negative halves: flb_time_to_millisec() divides the nanoseconds as a signed long
    def f64ev(d):
        return b"\x92\x92\xcb" + struct.pack(">d", d) + b"\x80\x81\xa1a\x01"
    def extev(sec, nsec):
        return b"\x92\x92\xd7\x00" + struct.pack(">iI", sec, nsec & 0xffffffff) + b"\x80\x81\xa1a\x01"
    odd_times = [f64ev(1.5) * 2, f64ev(-1.5) * 2, f64ev(-0.25), f64ev(-1700000000.75), f64ev(1700000000.123456), extev(5, 0x80000001),
                 extev(-7, 0xfffffff0), extev(-1, 999999999), extev(0x7fffffff, 0x7fffffff)]
    for c in odd_times:
        for df in range(5):
            assert ctx.to_json(c, 3, df, "date", True)[0] == ref.to_json(c, 3, df, "date", True), (c, df)
    # 64-bit seconds up to the last value gmtime_r() takes (the decoder steps over the ones behind it, up to 2^56)
    good = b"\x92\xcf" + struct.pack(">Q", 1700000000) + b"\x81\xa1a\x01"
    for sec in (253402300799, 253402300800, 2 ** 40, 2 ** 50, 2 ** 55, 67767976233532799, 67767976233532800, 2 ** 56 - 2 ** 20):
        forms = [b"\x92\xcf" + struct.pack(">Q", sec) + b"\x81\xa1a\x01", b"\x92\x92\xcf" + struct.pack(">Q", sec) + b"\x80\x81\xa1a\x01", f64ev(float(sec))]
        if sec <= 2 ** 55:              # (negative seconds are not stepped over: past gmtime_r()'s range the reference formats a stale struct tm)
            forms += [b"\x92\x92\xd3" + struct.pack(">q", -sec) + b"\x80\x81\xa1a\x01", f64ev(-float(sec) - 0.5)]
        for c in forms:
            for df in range(5):
                for key in ("date", "d" * 40):
                    assert ctx.to_json(good + c + good, 3, df, key, True)[0] == ref.to_json(good + c + good, 3, df, key, True), (sec, c, df)
    for d in (1e300, -1e300, float("nan"), 9.3e18, -9.3e18):
        for df in (0, 2, 4):              # (the calendar formats of such a time are gmtime_r()'s failure path)
            assert ctx.to_json(f64ev(d) * 2, 3, df, "date", True)[0] == ref.to_json(f64ev(d) * 2, 3, df, "date", True), (d, df)
    # a string as the event's last value, out of step after multi-byte characters: counted, not compared
    odd = util.event(1700000000, 0, [(b"m", S(("é" * 15).encode() + b"abc"))])
    assert ctx.to_json(odd, 3, 0, "date", False)[1] == 1


def _extremes(lib):
    """every header width of strings, bins, keys, maps, arrays and exts, containers nested past msgpack-c's limit, keys that are
    not strings, a body that is not a map: the text, or `nothing`, as the reference gives it"""
    ref = util.Ref()
    ctx = pkg.Context(0, lib=lib)

    def SS(b):
        n = len(b)
        return S(b) if n < 256 else (b"\xda" + struct.pack(">H", n) + b if n < 65536 else b"\xdb" + struct.pack(">I", n) + b)

    def ev(body):
        return b"\x92\x92\xd7\x00" + struct.pack(">II", 1700000000, 5) + b"\x80" + body

    def m(n):
        return bytes([0x80 | n]) if n < 16 else (b"\xde" + struct.pack(">H", n) if n < 65536 else b"\xdf" + struct.pack(">I", n))

    def a(n):
        return bytes([0x90 | n]) if n < 16 else (b"\xdc" + struct.pack(">H", n) if n < 65536 else b"\xdd" + struct.pack(">I", n))

    def binhdr(n):
        return b"\xc4" + bytes([n]) if n < 256 else (b"\xc5" + struct.pack(">H", n) if n < 65536 else b"\xc6" + struct.pack(">I", n))
    tail = S(b"z") + S(b"end of the event, long enough")
    cases = {}
    for n in (31, 32, 255, 256, 65535, 65536, 70000):
        cases["str%d" % n] = ev(m(2) + S(b"s") + SS(b"x" * n) + tail)
        cases["bin%d" % n] = ev(m(2) + S(b"s") + binhdr(n) + b"\x01" * n + tail)
        cases["key%d" % n] = ev(m(2) + SS(b"k" * n) + b"\x01" + tail)
    for n in (15, 16, 17, 300, 2000):                   # (the duplicate-key rule is quadratic in the members, there as here)
        cases["map%d" % n] = ev(m(n + 1) + b"".join(S(b"k%d" % i) + b"\x01" for i in range(n)) + tail)
        cases["arr%d" % n] = ev(m(2) + S(b"a") + a(n) + b"\x02" * n + tail)
        cases["dup%d" % n] = ev(m(n + 1) + b"".join(S(b"k%d" % (i % 7)) + bytes([i % 100]) for i in range(n)) + tail)
    cases["arr70000"] = ev(m(2) + S(b"a") + a(70000) + b"\x02" * 70000 + tail)
    # the wide headers on small containers (msgpack-c never writes them; the unpacker takes them)
    cases["map32_of_3"] = ev(b"\xdf" + struct.pack(">I", 3) + S(b"a") + b"\x01" + S(b"b") + b"\xdd" + struct.pack(">I", 2) + b"\x01\x02" + tail)
    cases["map16_of_2"] = ev(b"\xde" + struct.pack(">H", 2) + S(b"a") + b"\xdc" + struct.pack(">H", 1) + b"\x07" + tail)
    for d in (1, 5, 16, 31, 32, 33, 40, 100):
        cases["nestmap%d" % d] = ev(m(2) + S(b"n") + (m(1) + S(b"k")) * d + b"\x01" + tail)
        cases["nestarr%d" % d] = ev(m(2) + S(b"n") + a(1) * d + b"\x01" + tail)
    for t, nb in ((0xd4, 1), (0xd5, 2), (0xd6, 4), (0xd7, 8), (0xd8, 16)):
        cases["fixext%d" % nb] = ev(m(2) + S(b"e") + bytes([t, 5]) + bytes(range(0x78, 0x78 + nb)) + tail)
    for n in (0, 1, 17, 255):
        cases["ext8_%d" % n] = ev(m(2) + S(b"e") + b"\xc7" + bytes([n, 0x85]) + b"\x81" * n + tail)
    cases["ext16"] = ev(m(2) + S(b"e") + b"\xc8" + struct.pack(">H", 300) + b"\x01" + b"\xfe" * 300 + tail)
    cases["nil_true_false"] = ev(m(4) + S(b"a") + b"\xc0" + S(b"b") + b"\xc2" + S(b"c") + b"\xc3" + tail)
    ints = [b"\xcc\xff", b"\xcd\xff\xff", b"\xce\xff\xff\xff\xff", b"\xcf" + b"\xff" * 8, b"\xd0\x80", b"\xd1\x80\x00", b"\xd2\x80\x00\x00\x00",
            b"\xd3\x80" + b"\x00" * 7]
    cases["ints"] = ev(m(9) + b"".join(S(b"i%d" % i) + v for i, v in enumerate(ints)) + tail)
    cases["keys_not_strings"] = ev(m(6) + b"\x01\x02" + b"\xc0\x03" + b"\xc3\x04" + b"\xcb" + struct.pack(">d", 1.5) + b"\x05" + b"\x91\x01\x06" + tail)
    cases["body_not_a_map"] = b"\x92\x92\xd7\x00" + struct.pack(">II", 1, 1) + b"\x80" + b"\x93\x01\x02\x03"
    cases["empty_body"] = ev(m(0))
    texts = 0
    for name, c in cases.items():
        for jf in (1, 3):
            for esc in (True, False):
                got, und = ctx.to_json(c, jf, 1, "date", esc)
                assert und == 0 and got == ref.to_json(c, jf, 1, "date", esc), (name, jf, esc)
                texts += got is not None
    assert texts > 3 * len(cases)


def test_tojson_diff_hostsim(sim_lib, ref_available):
    _diff(sim_lib, 150, 40)


def test_tojson_numbers_hostsim(sim_lib, ref_available):
    _numbers(sim_lib)


def test_tojson_strings_hostsim(sim_lib, ref_available):
    _strings(sim_lib)


def test_tojson_edges_hostsim(sim_lib, ref_available):
    _edges(sim_lib)


def test_tojson_extremes_hostsim(sim_lib, ref_available):
    _extremes(sim_lib)


@pytest.mark.gpu
def test_tojson_diff_gpu(gpu_lib, ref_available):
    _diff(gpu_lib, 40, 40)


@pytest.mark.gpu
def test_tojson_numbers_gpu(gpu_lib, ref_available):
    _numbers(gpu_lib)


@pytest.mark.gpu
def test_tojson_strings_gpu(gpu_lib, ref_available):
    _strings(gpu_lib)


@pytest.mark.gpu
def test_tojson_edges_gpu(gpu_lib, ref_available):
    _edges(gpu_lib)


@pytest.mark.gpu
def test_tojson_extremes_gpu(gpu_lib, ref_available):
    _extremes(gpu_lib)
