"""Random records (duplicate keys, non-string keys, bin / ext / nil / nested values, empty maps, metadata)
through randomly configured grep / modify / record_modifier / parser filters and short chains of them:
CPU emulation of the device code vs the unmodified reference.  Refusals (loud) are counted, not failures.
usage: python tests/tools/filterfuzz.py SEED ROUNDS"""
import os
import random
import struct
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import util

pkg = util.pkg
S = util.mp_str
# no empty key: the reference's modify reads through a NULL key pointer when a second rule meets one (segfault here)
KEYS = [b"log", b"level", b"k1", b"k2", b"a3", b"msg", b"nest", b"arr", b"n", b"n", b"flag", b"Key", b"k 1", b"LOG"]
WORDS = [b"GET /a HTTP/1.1", b"error", b"warn", b"info", b"", b"sample1", b"z2", b"true", b"false", b"123", b"a b", b"\xc3\xa9", b"x\x00y",
         b'{"a":1,"b":"c"}', b"k=v j=2", b"a:1\tb:2",
         b'{"time":"2023-05-06T07:08:09.123Z","level":"warn","n":-5,"f":1.5e3,"nest":{"k1":"v"},"arr":[1,"two",null],"u":"\\u00e9\\ud83d\\ude00"}',
         b'{"time":"not a time","level":"info"}', b'{"time":"2023-05-06T07:08:09Z"}', b'{"a":1} trailing', b'[1,2]', b'{"a":{"b":{"c":{"d":1}}},"e":[]}', b'{"dup":1,"dup":2,"":3}',
         b'time=2023-05-06T07:08:09.5Z level=warn msg="quoted \\"text\\"" bare n=7', b"time=bad level=info", b"=x a==b c=\"\"",
         b"time:2023-05-06T07:08:09.250Z\tlevel:warn\tn:7\tempty:\t:nolabel", b"time:xx\tlevel:info",
         b"POST /p?q=1 took 2023-05-06T07:08:09.750Z 12ab 0x1F 3.25 TRUE", b"GET / took bad-time 7 ff 1e3 false"]


def value(rng, depth=0):
    r = rng.random()
    if r < 0.45:
        return S(rng.choice(WORDS))
    if r < 0.55:
        return rng.choice([b"\x00", b"\x7f", b"\xff", b"\xe0", b"\xcc\x80", b"\xcd\x01\x00", b"\xce\x00\x01\x00\x00", b"\xd0\x80", b"\xd1\xff\x00",
                           b"\xcf" + struct.pack(">Q", 2 ** 63 + 5), b"\xd3" + struct.pack(">q", -2 ** 40), b"\xcc\x05", b"\xd2\x00\x00\x00\x07"])
    if r < 0.62:
        return rng.choice([b"\xc2", b"\xc3", b"\xc0"])
    if r < 0.68:
        return rng.choice([b"\xcb" + struct.pack(">d", 0.5), b"\xca" + struct.pack(">f", 1.25), b"\xcb" + struct.pack(">d", -1e300)])
    if r < 0.74:
        w = rng.choice(WORDS)
        return bytes([0xc4, len(w)]) + w
    if r < 0.78:
        return b"\xd4\x05\x01" if rng.random() < 0.5 else b"\xc7\x03\x07abc"
    if r < 0.82:
        # spellings msgpack-c would not write (wider than needed): the reference re-packs them, raw copies keep them
        w = rng.choice(WORDS)[:200]
        return rng.choice([b"\xd9" + bytes([len(w)]) + w, b"\xda" + struct.pack(">H", len(w)) + w, b"\xdb" + struct.pack(">I", len(w)) + w,
                           b"\xcd\x00\x05", b"\xce\x00\x00\x00\x05", b"\xcf" + struct.pack(">Q", 5), b"\xd1\xff\xfb", b"\xd2\xff\xff\xff\xfb",
                           b"\xd3" + struct.pack(">q", -5), b"\xd3" + struct.pack(">q", 5), b"\xde\x00\x01" + S(b"k1") + b"\xcd\x00\x07",
                           b"\xdf\x00\x00\x00\x00", b"\xdc\x00\x02\x01\xd9\x01x", b"\xdd\x00\x00\x00\x00", b"\xc5\x00\x02ab", b"\xc6\x00\x00\x00\x01z",
                           b"\xc8\x00\x01\x09q", b"\xca" + struct.pack(">f", 0.1)])
    if depth < 2 and r < 0.9:
        n = rng.randrange(0, 4)
        return util.mp_map_hdr(n) + b"".join(key(rng) + value(rng, depth + 1) for _ in range(n))
    if depth < 2:
        n = rng.randrange(0, 4)
        return bytes([0x90 | n]) + b"".join(value(rng, depth + 1) for _ in range(n))
    return S(b"deep")


def key(rng):
    r = rng.random()
    if r < 0.85:
        return S(rng.choice(KEYS))
    if r < 0.9:
        k = rng.choice(KEYS)                       # a key spelled wider than needed
        return rng.choice([b"\xd9" + bytes([len(k)]) + k, b"\xda" + struct.pack(">H", len(k)) + k, b"\xc5" + struct.pack(">H", len(k)) + k])
    if r < 0.95:
        k = rng.choice(KEYS)
        return bytes([0xc4, len(k)]) + k
    return rng.choice([b"\x05", b"\xc3", b"\xc0"])


def record(rng, i):
    n = rng.choice([0, 1, 2, 3, 4, 5, 6, 8])
    body = util.mp_map_hdr(n) + b"".join(key(rng) + value(rng) for _ in range(n))
    meta = b"\x80" if rng.random() < 0.8 else util.mp_map_hdr(1) + S(b"m") + value(rng)
    return b"\x92\x92\xd7\x00" + struct.pack(">II", 1700000000 + i, i % 1000) + meta + body


def marker(rng):
    """group start / end and other events with a negative timestamp: stepped over by the event decoder"""
    sec = rng.choice([0xffffffff, 0xfffffffe, 0x80000001, 0xfffffffd])
    n = rng.choice([0, 1, 2])
    return b"\x92\x92\xd7\x00" + struct.pack(">II", sec, 0) + b"\x80" + util.mp_map_hdr(n) + b"".join(key(rng) + value(rng) for _ in range(n))


RA = ["log", "level", "k1", "$nest['k1']", "$nest['nest']['k2']", "$arr[1]", "$arr[0]['k1']", "$k1", "msg", "$nest", "$arr", "n", "flag", "$log['x']"]
RX = ["GET", "^(warn|error)$", ".", "^$", "sample[0-9]", "^[a-z][0-9]$", "true", "a b", "1", "\\d+", "^x", "/get/i", "(?i)WARN", "^(?<w>\\w+)\\s", "(a|b)\\1", "[[:digit:]]+",
      "\\bGET\\b", "^.{3,5}$", "(?:ab)*c", "é", "\\x41", "[^\\x00-\\x7f]", "^\\s*$", "\\A.*\\z", "T /a H", "\\.", "^(?!GET)"]


def grep_props(rng):
    props = []
    if rng.random() < 0.4:
        props.append(("Logical_Op", rng.choice(["and", "or", "legacy"])))
    kind = rng.choice(["Regex", "Exclude"])
    for _ in range(rng.randrange(1, 4)):
        k = kind if props and props[0][0] == "Logical_Op" and props[0][1] != "legacy" else rng.choice(["Regex", "Exclude"])
        props.append((k, "%s %s" % (rng.choice(RA), rng.choice(RX))))
    return props


PLAIN = ["log", "level", "k1", "k2", "a3", "msg", "nest", "arr", "n", "flag", "Key", "new", "k 1"]


def modify_props(rng):
    props = []
    for _ in range(rng.randrange(0, 3)):
        c = rng.choice(["Key_exists", "Key_does_not_exist", "A_key_matches", "No_key_matches", "Key_value_equals", "Key_value_does_not_equal",
                        "Key_value_matches", "Key_value_does_not_match", "Matching_keys_have_matching_values", "Matching_keys_do_not_have_matching_values"])
        if c in ("Key_exists", "Key_does_not_exist"):
            v = rng.choice(RA)
        elif c in ("A_key_matches", "No_key_matches"):
            v = rng.choice(RX)
        elif c in ("Key_value_equals", "Key_value_does_not_equal"):
            v = "%s %s" % (rng.choice(RA), rng.choice(["error", "sample1", "true", "123", "z2"]))
        elif c in ("Key_value_matches", "Key_value_does_not_match"):
            v = "%s %s" % (rng.choice(RA), rng.choice(RX))
        else:
            v = "%s %s" % (rng.choice(RX), rng.choice(RX))
        props.append(("Condition", "%s %s" % (c, v)))
    for _ in range(rng.randrange(1, 5)):
        r = rng.choice(["Set", "Add", "Remove", "Remove_wildcard", "Remove_regex", "Rename", "Hard_rename", "Copy", "Hard_copy", "Move_to_start", "Move_to_end"])
        if r in ("Remove", "Move_to_start", "Move_to_end"):
            v = rng.choice(PLAIN)
        elif r == "Remove_wildcard":
            v = rng.choice(["k", "l", "a", "ne", "K", "m"])
        elif r == "Remove_regex":
            v = rng.choice(RX)
        else:
            v = "%s %s" % (rng.choice(PLAIN if r not in ("Set", "Add") else ["new", "k1", "level", "x"]), rng.choice(["v", "k2", "new", "level", "value 1"]) if r in ("Set", "Add") else rng.choice(PLAIN))
            if "value 1" in v:
                v = v.replace("value 1", '"value 1"')
        props.append((r, v))
    return props


def recmod_props(rng):
    props = [("Record", "%s %s" % (rng.choice(["host", "k1", "new"]), rng.choice(["n1", "v"]))) for _ in range(rng.randrange(0, 3))]
    kind = rng.choice(["Remove_key", "Allowlist_key", "Whitelist_key", None])
    if kind:
        props += [(kind, rng.choice(["log", "LEVEL", "k*", "K1", "nest", "a*", "msg", "*"])) for _ in range(rng.randrange(1, 3))]
    return props or [("Record", "a b")]


TF = "%Y-%m-%dT%H:%M:%S.%LZ"
PARSERS = [dict(name="js", format="json"), dict(name="lf", format="logfmt"), dict(name="lt", format="ltsv"),
           dict(name="rx", format="regex", regex=r"^(?<m>[A-Z]+) (?<p>[^ ]+) (?<v>.*)$"), dict(name="kv", format="regex", regex=r"^(?<a>[a-z]+)(?<d>\d*)$", types="d:integer"),
           dict(name="jst", format="json", time_key="time", time_fmt=TF), dict(name="jsk", format="json", time_key="time", time_fmt=TF, time_keep=True, time_strict=False),
           dict(name="lft", format="logfmt", time_key="time", time_fmt=TF, types="n:integer"), dict(name="ltt", format="ltsv", time_key="time", time_fmt=TF, time_keep=True, types="n:hex"),
           dict(name="rxt", format="regex", regex=r"^(?<m>[A-Z]+) (?<p>[^ ]+) took (?<time>[^ ]+) (?<i>[^ ]+) (?<h>[^ ]+) (?<f>[^ ]+) (?<b>[^ ]+)$", time_key="time", time_fmt=TF,
                types="i:integer h:hex f:float b:bool"),
           dict(name="lfb", format="logfmt", logfmt_no_bare_keys=True), dict(name="jse", format="json", skip_empty=False)]


def parser_props(rng):
    props = [("Key_Name", rng.choice(["log", "msg", "k1", "$nest['k1']", "level"]))]
    props += [("Parser", p) for p in rng.sample([kw["name"] for kw in PARSERS], rng.randrange(1, 4))]
    if rng.random() < 0.5:
        props.append(("Reserve_Data", rng.choice(["On", "Off"])))
    if rng.random() < 0.5:
        props.append(("Preserve_Key", rng.choice(["On", "Off"])))
    return props


def l2m_props(rng):
    mode = rng.choice(["counter", "counter", "gauge", "histogram"])
    props = [("metric_mode", mode), ("metric_name", "m"), ("metric_description", "d"), ("tag", "t")]
    if mode != "counter":
        props.append(("value_field", rng.choice(["n", "$nest['n']", "k2", "flag"])))
    for _ in range(rng.randrange(0, 3)):
        props.append(rng.choice([("label_field", rng.choice(["level", "k1", "flag", "n"])), ("add_label", "x %s" % rng.choice(RA))]))
    if rng.random() < 0.4:
        props.append((rng.choice(["regex", "exclude"]), "%s %s" % (rng.choice(RA), rng.choice(RX))))
    if rng.random() < 0.2:
        props.append(("kubernetes_mode", "on"))
    if rng.random() < 0.2:
        props.append(("discard_logs", "on"))
    if mode == "histogram" and not any(k in ("label_field", "add_label", "kubernetes_mode") for k, _ in props):
        props.append(("label_field", "level"))     # the reference cannot print a label-less histogram that saw nothing (crash in cmt_encode_text)
    if mode == "histogram" and rng.random() < 0.5:
        props += [("bucket", rng.choice(["1", "0.5", "100", "7"])) for _ in range(rng.randrange(1, 4))]
    return props


MAKERS = {"grep": grep_props, "modify": modify_props, "record_modifier": recmod_props, "parser": parser_props, "log_to_metrics": l2m_props}


def l2m_ref_text(ref, f):
    import ctypes as C
    import re
    ref.L.flbref_l2m_cmt_text.restype = C.c_void_p
    ref.L.flbref_l2m_cmt_text.argtypes = [C.c_void_p]
    p = ref.L.flbref_l2m_cmt_text(f)
    t = C.string_at(p).decode(errors="replace")
    ref.L.flbref_cfree(C.c_void_p(p))
    return re.sub(r"^\S+Z ", "", t, flags=re.M)


def main(seed, rounds):
    rng = random.Random(seed)
    lib = pkg.load(util.HOSTSIM_SO)
    bad = refused = rejected = 0
    for rd in range(rounds):
        chunk = b"".join(record(rng, i) if rng.random() < 0.93 else marker(rng) for i in range(rng.choice([1, 5, 40])))
        filters = [(k, MAKERS[k](rng)) for k in (rng.choice(list(MAKERS)) for _ in range(rng.choice([1, 1, 2, 3])))]
        if sum(k == "log_to_metrics" for k, _ in filters) > 1:
            continue                                          # one metrics filter per fused chain
        ctx, ref = pkg.Context(0, lib=lib), util.Ref()
        for kw in PARSERS:
            ctx.parser(**kw); ref.parser(**kw)
        try:
            rfs = [ref.filter(p, props) for p, props in filters]
        except RuntimeError:
            rfs = None
        try:
            fs = [ctx.filter(p, props) for p, props in filters]
        except pkg.FlbGpuError as e:
            if rfs is not None and "malformed map" not in str(e):
                print("CONFIG refused here, accepted by the reference:", filters, str(e)[:100])
                rejected += 1
            continue
        if rfs is None:
            print("CONFIG accepted here, refused by the reference:", filters)
            bad += 1
            continue
        if os.environ.get("FILTERFUZZ_TRACE"):
            print("round", rd, filters, flush=True)
        want = ref.chain_do(chunk)
        try:
            got = ctx.chain(fs).do(chunk)
        except pkg.FlbGpuError as e:
            refused += 1
            if os.environ.get("FILTERFUZZ_TRACE"):
                print("refused:", str(e)[:160], filters, flush=True)
            continue
        for (k, _), f, rf in zip(filters, fs, rfs):
            if k == "log_to_metrics" and f.l2m_text() != l2m_ref_text(ref, rf):
                bad += 1
                print("METRICS MISMATCH seed=%d round=%d filters=%r\n got  %r\n want %r" % (seed, rd, filters, f.l2m_text()[:300], l2m_ref_text(ref, rf)[:300]))
        # the same filters called one by one, as the reference's flb_filter_do() would call the shim's cb_filter
        c3 = pkg.Context(0, lib=lib)
        for kw in PARSERS:
            c3.parser(**kw)
        cur, changed, seq = chunk, False, None
        try:
            for p_, props in filters:
                r3, o3 = c3.filter(p_, props).cb(cur)
                if r3 == 1:
                    changed = True
                    if not o3:
                        seq = (1, o3)                     # nothing left: the chain ends here
                        break
                    cur = o3
            if seq is None:
                seq = (1, cur) if changed else (2, None)
            if seq != want:
                bad += 1
                print("PER-FILTER MISMATCH seed=%d round=%d filters=%r" % (seed, rd, filters))
        except pkg.FlbGpuError:
            pass
        if got != want:
            bad += 1
            print("MISMATCH seed=%d round=%d filters=%r" % (seed, rd, filters))
            for i in range(0, 1):
                recs = util.split_records(chunk)
                for o, l in recs:
                    one = chunk[o:o + l]
                    c2, r2 = pkg.Context(0, lib=lib), util.Ref()
                    for kw in PARSERS:
                        c2.parser(**kw); r2.parser(**kw)
                    [r2.filter(p, props) for p, props in filters]
                    try:
                        g1 = c2.chain([c2.filter(p, props) for p, props in filters]).do(one)
                    except pkg.FlbGpuError:
                        continue
                    w1 = r2.chain_do(one)
                    if g1 != w1:
                        print("   record", one.hex()); print("   got ", g1[0], g1[1].hex() if g1[1] else None); print("   want", w1[0], w1[1].hex() if w1[1] else None)
                        break
    print("rounds", rounds, "mismatches", bad, "loud refusals", refused, "configs refused only here", rejected)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]), int(sys.argv[2])) else 0)
