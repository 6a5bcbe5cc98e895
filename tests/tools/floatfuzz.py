"""JSON numbers with more than 19 significant digits that sit on or next to a rounding boundary of
binary64: exact midpoints between adjacent doubles (ties), and the same with the last digit nudged.
Through filter_parser(json) on the CPU emulation of the device code vs the unmodified reference.
usage: python tests/tools/floatfuzz.py SEED N"""
import math
import os
import random
import struct
import sys
from fractions import Fraction

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import util

pkg = util.pkg


def dec(fr, extra=""):
    """exact decimal expansion of a Fraction whose denominator is a power of two (terminates)"""
    num, den = fr.numerator, fr.denominator
    ip = num // den
    rem = num - ip * den
    digs = []
    while rem:
        rem *= 10
        digs.append(str(rem // den))
        rem %= den
    return str(ip) + ("." + "".join(digs) if digs else "") + extra


def cases(rng, n):
    out = []
    for _ in range(n):
        r = rng.random()
        if r < 0.7:
            d = rng.uniform(1e-6, 1e15) if rng.random() < 0.8 else rng.uniform(1e15, 1e22)
        elif r < 0.85:
            d = rng.uniform(1e-300, 1e-290)
        else:
            d = rng.uniform(1e290, 1e300)
        bits = struct.unpack("<Q", struct.pack("<d", d))[0]
        up = struct.unpack("<d", struct.pack("<Q", bits + 1))[0]
        mid = (Fraction(d) + Fraction(up)) / 2
        s = dec(mid)
        if len(s) > 600:
            continue
        out.append(s)                                     # exact tie
        out.append(s + "1")                               # just above
        out.append(s + "0000000000000000000000001")
        if s[-1] != "0":
            out.append(s[:-1] + str(int(s[-1]) - 1) + "9999999999")   # just below
        out.append("-" + s)
        if rng.random() < 0.3 and "." not in s:
            out.append(s + "e0")
    # the edges: overflow boundary and the smallest subnormal
    big = Fraction(2) ** 1024 - Fraction(2) ** 970        # midpoint between DBL_MAX and 2^1024
    out += [dec(big), dec(big - 1), dec(big + 1)]
    tiny = Fraction(1, 2 ** 1075)
    t = dec(tiny)
    out += [t, t + "1", t[:-1] + "4"]
    return out


def main(seed, n):
    rng = random.Random(seed)
    lib = pkg.load(util.HOSTSIM_SO)
    nums = cases(rng, n)
    docs = [('{"v":%s,"w":[%s]}' % (x, x)).encode() for x in nums]
    bad = 0
    for b0 in range(0, len(docs), 400):
        part = docs[b0:b0 + 400]
        chunk = util.chunk_from_lines(part)
        kw = dict(name="js", format="json")
        ctx = pkg.Context(0, lib=lib)
        ref = util.Ref()
        ctx.parser(**kw); ref.parser(**kw)
        props = [("Key_Name", "log"), ("Parser", "js")]
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
                if c2.filter("parser", props).cb(c1) != r2.chain_do(c1):
                    print("MISMATCH", d[:120])
                    break
    print("numbers", len(nums), "bad batches", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]), int(sys.argv[2])) else 0)
