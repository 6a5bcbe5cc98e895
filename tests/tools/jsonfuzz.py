"""Differential fuzz of the JSON parser path: random and mutated documents through
filter_parser(json) on the CPU emulation of the device code vs the unmodified reference.
usage: python tests/tools/jsonfuzz.py SEED NDOCS"""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import util

pkg = util.pkg
WORDS = ["a", "level", "msg", "x y", "", "tab\\t", "q\\\"q", "caf\\u00e9", "\\ud83d\\ude00", "\\u0000", "sl\\/ash", "é", "日本",
         "\\ud800", "back\\\\slash", "ctl\x01", "long" * 20]


def rnd_str(rng):
    return '"' + rng.choice(WORDS) + (rng.choice(WORDS) if rng.random() < 0.3 else "") + '"'


def rnd_num(rng):
    r = rng.random()
    if r < 0.4:
        return str(rng.randint(-10 ** rng.randint(1, 20), 10 ** rng.randint(1, 20)))
    if r < 0.7:
        return "%.*f" % (rng.randint(0, 12), rng.uniform(-1e6, 1e6))
    if r < 0.85:
        return "%de%d" % (rng.randint(-999, 999), rng.randint(-30, 30))
    return rng.choice(["0", "-0", "0.0", "-0.0", "1E5", "1e+5", "01", "1.", ".5", "-", "1e", "123456789012345678", "1234567890123456789",
                       "18446744073709551615", "18446744073709551616", "-9223372036854775808", "1e400", "0.1e-400"])


def rnd_val(rng, depth):
    r = rng.random()
    if r < 0.3:
        return rnd_str(rng)
    if r < 0.55:
        return rnd_num(rng)
    if r < 0.7:
        return rng.choice(["true", "false", "null"])
    if depth > 4:
        return "1"
    ws = rng.choice(["", "", " ", "\n", "\t "])
    if r < 0.85:
        n = rng.choice([0, 1, 2, 3, 17])
        return "{" + ws + ("," + ws).join(rnd_str(rng) + ws + ":" + ws + rnd_val(rng, depth + 1) for _ in range(n)) + ws + "}"
    n = rng.choice([0, 1, 2, 5, 16, 20])
    return "[" + ws + ",".join(rnd_val(rng, depth + 1) for _ in range(n)) + ws + "]"


def rnd_doc(rng):
    n = rng.choice([0, 1, 2, 4, 8, 17])
    ws = rng.choice(["", "", " ", "\r\n"])
    d = ws + "{" + ws + ("," + ws).join(rnd_str(rng) + ":" + ws + rnd_val(rng, 1) for _ in range(n)) + ws + "}" + ws
    r = rng.random()
    if r < 0.15 and d:                                   # mutate one byte
        i = rng.randrange(len(d))
        d = d[:i] + rng.choice(['"', "\\", ",", ":", "{", "}", "[", "]", " ", "x", "1"]) + d[i + 1:]
    elif r < 0.2:
        d = d + rng.choice(["x", "{}", " 1", ",", "]"])
    elif r < 0.23:
        d = d[:rng.randrange(len(d) + 1)]
    return d.encode("utf-8", "surrogatepass") if False else d.encode()


def main(seed, ndocs):
    rng = random.Random(seed)
    lib = pkg.load(util.HOSTSIM_SO)
    bad = 0
    for batch in range(0, ndocs, 500):
        docs = [rnd_doc(rng) for _ in range(min(500, ndocs - batch))]
        chunk = util.chunk_from_lines(docs)
        ctx = pkg.Context(0, lib=lib)
        ref = util.Ref()
        kw = dict(name="js", format="json", time_key="t", time_fmt="%s", time_keep=rng.random() < 0.5, skip_empty=rng.random() < 0.5,
                  time_strict=rng.random() < 0.5)
        ctx.parser(**kw); ref.parser(**kw)
        props = [("Key_Name", "log"), ("Parser", "js"), ("Reserve_Data", rng.choice(["On", "Off"])), ("Preserve_Key", rng.choice(["On", "Off"]))]
        # what follows the parser in a fused chain works on the parsed field list: keys of every kind of value
        tail = rng.choice([[], [("grep", [("Exclude", "a ^x")])], [("modify", [("Rename", "a b")])], [("modify", [("Remove_wildcard", "k")])],   # (one rule: a second one makes the reference read through a NULL key on "" keys)
                           [("record_modifier", [("Remove_key", "a"), ("Record", "h n")])], [("grep", [("Regex", "$a['b'] .")]), ("modify", [("Copy", "t t2")])]])
        fs = [ctx.filter("parser", props)] + [ctx.filter(p_, pr) for p_, pr in tail]
        ref.filter("parser", props)
        for p_, pr in tail:
            ref.filter(p_, pr)
        f = ctx.chain(fs)
        f.cb = f.do
        want = ref.chain_do(chunk)
        try:
            got = f.cb(chunk)
        except pkg.FlbGpuError as e:
            # loud refusals (floats outside the exact path) are allowed; find and report the count
            print("batch %d refused: %s" % (batch, e))
            continue
        if got != want:
            bad += 1
            for d in docs:
                c1 = util.chunk_from_lines([d])
                r2 = util.Ref(); r2.parser(**kw); r2.filter("parser", props)
                for p_, pr in tail:
                    r2.filter(p_, pr)
                c2 = pkg.Context(0, lib=lib); c2.parser(**kw)
                try:
                    g = c2.chain([c2.filter("parser", props)] + [c2.filter(p_, pr) for p_, pr in tail]).do(c1)
                except pkg.FlbGpuError:
                    continue
                if g != r2.chain_do(c1):
                    print("MISMATCH", d, kw, props, tail)
                    break
    print("docs", ndocs, "bad batches", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]), int(sys.argv[2])) else 0)
