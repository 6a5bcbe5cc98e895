"""Differential fuzzer for filter_multiline (buffer off): random rule sets over a small alphabet of states and patterns, random
lines, several chunks per filter instance -- the CPU emulation of the device code against the reference's own plugin.
  python tests/tools/mlfuzz.py [rounds] [seed]"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
import test_zy_multiline as T

pkg = util.pkg
PATS = [r"/^A/", r"/^B/", r"/^\s+at/", r"/^$/", r"/x/", r"/^[a-c]+$/", r"/./", r"/^(E|F).*:$/", r"/\d+/", r"/^ /", r"/\n$/"]
STATES = ["s1", "s2", "s3", "start_state"]
WORDS = [b"A", b"B", b"A one", b"B two", b"  at f()", b"", b"x", b"abc", b"E:", b"F 12:", b"12", b" ", b"A\n", b"\n", b"zzz", b"  at g()\n", b"cab x"]


def rules(rng):
    n = rng.randint(1, 6)
    out = []
    for i in range(n):
        frm = ["start_state"] if i == 0 else rng.sample(STATES, rng.randint(1, 2))
        if i and rng.random() < 0.25 and "start_state" not in frm:
            frm.append("start_state")
        out.append([", ".join(frm), rng.choice(PATS), None])
    froms = sorted({s.strip() for r in out for s in r[0].split(",")})
    for r in out:
        r[2] = rng.choice(froms)
    return [tuple(r) for r in out]


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    lib = pkg.load(util.HOSTSIM_SO)
    bad = 0
    for k in range(rounds):
        rs = rules(rng)
        props = [("multiline.parser", "f"), ("multiline.key_content", "log"), ("buffer", "off")]
        wm = rng.choice([0.0, 0.2, 1.0])
        chunks = [T.make_chunk(rng, WORDS, rng.choice([1, 3, 20, 150]), 1700000000 + 1000 * c, with_meta=wm) for c in range(rng.randint(1, 4))]
        try:
            T.diff(lib, rs, props, chunks, name="f")
        except AssertionError as e:
            bad += 1
            print("MISMATCH round %d rules %r\n%s" % (k, rs, e))
            if bad > 3:
                break
    print("mlfuzz: %d rounds, %d mismatches" % (rounds, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
