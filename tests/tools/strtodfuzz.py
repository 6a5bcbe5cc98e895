"""`Types v:float` (flb_parser_typecast -> atof -> glibc strtod) through a regex parser: the
rounding-boundary decimals of floatfuzz.py plus strtod's own syntax (white space, '+', bare dots,
dangling exponents, inf / nan, trailing text, over- and underflow, leading zeros, random decimals).
CPU emulation of the device code vs the unmodified reference.
usage: python tests/tools/strtodfuzz.py SEED N"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import floatfuzz
import util

pkg = util.pkg

import cases

SYNTAX = cases.STRTOD_TEXTS


def rand_decimal(rng):
    nd = rng.choice([1, 3, 8, 15, 16, 17, 18, 19, 20, 21, 25, 40, 80])
    digs = "".join(rng.choice("0123456789") for _ in range(nd))
    if rng.random() < 0.6:
        k = rng.randrange(0, nd + 1)
        digs = digs[:k] + "." + digs[k:]
    if rng.random() < 0.6:
        digs += rng.choice("eE") + rng.choice(["", "+", "-"]) + str(rng.choice([0, 1, 5, 22, 23, 30, 100, 290, 300, 307, 308, 310, 320, 324, 340]))
    return rng.choice(["", "", "-", "+", " "]) + digs


def main(seed, n):
    rng = random.Random(seed)
    lib = pkg.load(util.HOSTSIM_SO)
    texts = list(SYNTAX) + floatfuzz.cases(rng, n) + [rand_decimal(rng) for _ in range(4 * n)]
    lines = [t.encode() for t in texts if "\n" not in t]
    kw = dict(name="f", format="regex", regex=r"^(?<v>[^\n]*)$", types="v:float")
    props = [("Key_Name", "log"), ("Parser", "f")]
    bad = 0
    for b0 in range(0, len(lines), 400):
        part = lines[b0:b0 + 400]
        chunk = util.chunk_from_lines(part)
        ctx = pkg.Context(0, lib=lib)
        ref = util.Ref()
        ctx.parser(**kw); ref.parser(**kw)
        f = ctx.filter("parser", props)
        ref.filter("parser", props)
        try:
            got = f.cb(chunk)
        except pkg.FlbGpuError as e:
            print("refused:", e)
            bad += 1
            continue
        if got != ref.chain_do(chunk):
            bad += 1
            for d in part:
                c1 = util.chunk_from_lines([d])
                r2 = util.Ref(); r2.parser(**kw); r2.filter("parser", props)
                c2 = pkg.Context(0, lib=lib); c2.parser(**kw)
                a, b = c2.filter("parser", props).cb(c1), r2.chain_do(c1)
                if a != b:
                    print("MISMATCH", d[:120], a[-9:].hex() if a else a, b[-9:].hex() if b else b)
    print("texts", len(lines), "bad batches", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]), int(sys.argv[2])) else 0)
