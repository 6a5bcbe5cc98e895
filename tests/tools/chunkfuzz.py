"""Chunk-level fuzz: runs of events glued together with group markers and garbage in between, often cut at a
random byte, several slices per call (FLBGPU_SLICE_MB=1), streaming and classic form -- the north-star chain
and a grep on the CPU emulation vs the unmodified reference.  It found that a chunk cut short is still
"clean" for grep / modify when msgpack-c's parser happens to eat the whole tail (runtime.c:
msgpack_tail_runs_out_cleanly).
usage: python tests/tools/chunkfuzz.py SEED TRIALS"""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
os.environ["FLBGPU_SLICE_MB"] = "1"
import cases
import test_chain_parity as T
import util

pkg = util.pkg


def main(seed, trials):
    lib = pkg.load(util.HOSTSIM_SO)
    rng = random.Random(seed)
    block = util.chunk_from_lines(util.apache_lines(3000, seed=rng.randrange(1000)))
    recs = util.split_records(block)
    case = [c for c in cases.CASES if c[0] == "north_star_chain"][0]
    grp, gre = util.event(0xffffffff, 0, [(b"g", util.mp_str(b"s"))]), util.event(0xfffffffe, 0, [])
    bad = 0
    for trial in range(trials):
        parts = []
        for _ in range(rng.randrange(3, 40)):
            a = rng.randrange(len(recs))
            b = min(len(recs), a + rng.randrange(1, 2500))
            parts.append(block[recs[a][0]: recs[b - 1][0] + recs[b - 1][1]])
            r = rng.random()
            if r < 0.15:
                parts.append(grp)
            elif r < 0.3:
                parts.append(gre)
            elif r < 0.36:
                parts.append(rng.choice([b"\xc1", b"\x92\x01", b"\x92\x92\xd7", b"garbage"]))
        chunk = b"".join(parts)
        if rng.random() < 0.5:
            chunk = chunk[: rng.randrange(len(chunk))]
        for stream in ("1", "0"):
            os.environ["FLBGPU_STREAM"] = stream
            try:
                T.run_case(lib, case[1], case[2], chunk)
                T.run_case(lib, [], [("grep", [("Exclude", "log POST")])], chunk)
                T.run_case(lib, [], [("modify", [("Condition", "Key_value_matches log POST"), ("Add", "m 1")])], chunk)
            except AssertionError:
                bad += 1
                print("MISMATCH seed=%d trial=%d stream=%s bytes=%d" % (seed, trial, stream, len(chunk)))
    print("trials", trials, "mismatches", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]), int(sys.argv[2])) else 0)
