"""Differential fuzz of logfmt values with backslash escapes (flb_unescape_string_utf8 + strlen) through
filter_parser(logfmt) on the CPU emulation of the device code vs the unmodified reference.
usage: python tests/tools/logfmtfuzz.py SEED NLINES"""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import util

pkg = util.pkg
BS = "\\"
PIECES = [BS + "n", BS + "t", BS + '"', BS + BS, BS + "/", BS + "b", BS + "f", BS + "r", BS + "v", BS + "a", BS + "x41", BS + "x4", BS + "xZ",
          BS + "u0041", BS + "u00e9", BS + "u20ac", BS + "ud83d" + BS + "ude00", BS + "ud83d", BS + "udc00", BS + "u12", BS + "u", BS + "uZZZZ",
          BS + "U0001F600", BS + "U41", BS + "U", BS + "101", BS + "7", BS + "0", BS + "400", BS + "q", BS + "é", "é", "日本", "plain", " ",
          "a=b", BS + "ud83d" + BS + "u0041", BS + "ud83d" + BS + "u", BS + "'", BS + BS + BS + BS, "x" + BS]


def main(seed, nlines):
    rng = random.Random(seed)
    lib = pkg.load(util.HOSTSIM_SO)
    lines = []
    for i in range(nlines):
        v = "".join(rng.choice(PIECES) for _ in range(rng.randint(1, 6)))
        lines.append(('k1=v1 msg="%s" n=%d other="%s" bare' % (v, i, rng.choice(PIECES))).encode("utf-8"))
    chunk = util.chunk_from_lines(lines)
    bad = 0
    for types in (None, "n:integer"):
        kw = dict(name="lf", format="logfmt", types=types)
        ctx = pkg.Context(0, lib=lib)
        ref = util.Ref()
        ctx.parser(**kw); ref.parser(**kw)
        props = [("Key_Name", "log"), ("Parser", "lf")]
        f = ctx.filter("parser", props)
        ref.filter("parser", props)
        if f.cb(chunk) != ref.chain_do(chunk):
            bad += 1
            for l in lines:
                c1 = util.chunk_from_lines([l])
                r2 = util.Ref(); r2.parser(**kw); r2.filter("parser", props)
                c2 = pkg.Context(0, lib=lib); c2.parser(**kw)
                if c2.filter("parser", props).cb(c1) != r2.chain_do(c1):
                    print("MISMATCH", l)
                    break
    print("lines", nlines, "bad", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]), int(sys.argv[2])) else 0)
