"""Structure fuzz of the two key/value line formats: random sequences of the bytes that steer ltsv_parser()
(src/flb_parser_ltsv.c:82-197: label and field byte classes, ':' and TAB, CR / LF / NUL) and logfmt_parser()
(src/flb_parser_logfmt.c:63-254: '=', blanks, quotes, backslashes, control bytes), with Types, a time key,
skip_empty off and logfmt_no_bare_keys -- filter_parser on the CPU emulation vs the unmodified reference,
line by line (parsed or not, map bytes, time).
usage: python tests/tools/kvfuzz.py SEED NLINES"""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import util

pkg = util.pkg
LTSV_TOK = [b"a", b"key", b"k_1", b"x.y-z", b"time", b"n", b":", b":", b"\t", b"\t", b" ", b"\n", b"\r", b"\x00", b"\x01", b"=", b'"', b"\xc3\xa9", b"1f", b"12",
            b"2023-05-06T07:08:09.5Z", b"::", b"\t\t", b"UP", b"/"]
LOGFMT_TOK = [b"a", b"key", b"k_1", b"time", b"n", b"=", b"=", b" ", b" ", b"  ", b'"', b'"', b"\\", b'\\"', b"\\n", b"\t", b"\n", b"\x00", b"\x01", b"\x7f", b"\xc3\xa9", b"12", b"-7",
              b"2023-05-06T07:08:09.5Z", b"==", b'""', b":", b"'", b"v w"]
TF = "%Y-%m-%dT%H:%M:%S.%LZ"
CONFIGS = [("ltsv", LTSV_TOK, [dict(), dict(types="n:integer a:hex"), dict(time_key="time", time_fmt=TF), dict(time_key="time", time_fmt=TF, time_keep=True, skip_empty=False)]),
           ("logfmt", LOGFMT_TOK, [dict(), dict(types="n:integer a:bool"), dict(time_key="time", time_fmt=TF, time_keep=True), dict(logfmt_no_bare_keys=True), dict(skip_empty=False)])]


def main(seed, nlines):
    rng = random.Random(seed)
    lib = pkg.load(util.HOSTSIM_SO)
    bad = 0
    for fmt, toks, variants in CONFIGS:
        lines = [b"".join(rng.choice(toks) for _ in range(rng.randint(0, 14))) for _ in range(nlines)]
        for extra in variants:
            kw = dict(name="p", format=fmt, **extra)
            ctx, ref = pkg.Context(0, lib=lib), util.Ref()
            p, rp = ctx.parser(**kw), ref.parser(**kw)
            for line, (r, data, t) in zip(lines, p.do_batch(lines)):
                rr, rdata, rt = ref.parser_do(rp, line)
                if (r < 0) != (rr < 0) or (rr >= 0 and (data != rdata or t != (rt[0] & 0xffffffff, rt[1]))):
                    bad += 1
                    if bad <= 20:
                        print("MISMATCH", fmt, extra, line, (r, data, t), (rr, rdata, rt))
            props = [("Key_Name", "log"), ("Parser", "p"), ("Reserve_Data", "On")]
            ctx.parser(**dict(kw, name="p")) if False else None
            f = ctx.filter("parser", props)
            ref.filter("parser", props)
            chunk = util.chunk_from_lines(lines)
            if f.cb(chunk) != ref.chain_do(chunk):
                bad += 1
                print("FILTER MISMATCH", fmt, extra)
    print("lines per format", nlines, "mismatches", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]), int(sys.argv[2])) else 0)
