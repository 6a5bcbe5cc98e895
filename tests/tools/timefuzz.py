"""Time strings for the Time_Format values of the reference's conf/parsers*.conf (and a few more
directives: %s %e %j %y %p %I %a %A %B %Z %%), valid, edge (leap days, 60th second, year 0 / 9999, 24:00),
and mutated (digits changed, truncated, doubled blanks, wrong case, trailing text): the parsed
(seconds, nanoseconds) -- or the failure -- of flbgpu_parser_do_batch() on the CPU emulation vs
flb_parser_do() of the unmodified reference, with and without Time_Offset / time_strict.
usage: python tests/tools/timefuzz.py SEED N"""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import util

pkg = util.pkg
FORMATS = ["%Y-%m-%dT%H:%M:%S.%L%z", "%Y-%m-%d %H:%M:%S.%L", "%d/%b/%Y:%H:%M:%S %z", "%Y-%m-%dT%H:%M:%S.%L", "%b %d %H:%M:%S", "%Y-%m-%dT%H:%M:%S",
           "%Y-%m-%d %H:%M:%S,%L", "%Y-%m-%d %H:%M:%S", "%d-%b-%Y::%H:%M:%S", "%Y-%m-%dT%H:%M:%SZ", "%Y-%m-%dT%H:%M:%S.%LZ", "%Y-%m-%d %H:%M:%S %z",
           "%s", "%s.%L", "%e %B %Y %I:%M:%S %p", "%a %b %d %H:%M:%S %Y", "%A, %d-%b-%y %H:%M:%S %Z", "%Y%m%d%H%M%S", "%j %Y %H%%%M", "%m/%d/%Y %H:%M:%S.%L %z",
           "%D %T", "%F %R", "%H:%M:%S"]
MON = ["Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"]
MONL = ["January", "February", "March", "April", "May", "June", "July", "August", "September", "October", "November", "December"]
DAY = ["Sun", "Mon", "Tue", "Wed", "Thu", "Fri", "Sat"]
DAYL = ["Sunday", "Monday", "Tuesday", "Wednesday", "Thursday", "Friday", "Saturday"]


def render(fmt, rng):
    y = rng.choice([1970, 1999, 2000, 2023, 2024, 2038, 2100, 1969, 1900, 9999, 0, 68, 69, rng.randrange(1, 3000)])
    mo, d = rng.randrange(1, 13), rng.choice([1, 9, 28, 29, 30, 31, rng.randrange(1, 32)])
    h, mi, s = rng.choice([0, 12, 23, 24, rng.randrange(0, 24)]), rng.choice([0, 59, 60, rng.randrange(0, 60)]), rng.choice([0, 59, 60, 61, rng.randrange(0, 60)])
    frac = rng.choice(["0", "5", "123", "123456", "123456789", "1234567890123", "000000001", "999999999", ""])
    tz = rng.choice(["+0000", "-0700", "+0530", "+05:30", "Z", "-1200", "+1400", "+9999", "UTC", "GMT", "+01", "-0", ""])
    out, i = [], 0
    while i < len(fmt):
        c = fmt[i]
        if c != "%":
            out.append(c); i += 1; continue
        k = fmt[i + 1]; i += 2
        out.append({"Y": "%04d" % y if rng.random() < 0.9 else str(y), "m": "%02d" % mo, "d": "%02d" % d, "e": "%2d" % d, "H": "%02d" % h, "M": "%02d" % mi, "S": "%02d" % s,
                    "L": frac, "z": tz, "Z": rng.choice(["UTC", "GMT", "EST", "Z", "PDT", ""]), "b": MON[mo - 1], "B": MONL[mo - 1], "a": DAY[d % 7], "A": DAYL[d % 7],
                    "y": "%02d" % (y % 100), "j": "%03d" % rng.randrange(1, 367), "I": "%02d" % ((h % 12) or 12), "p": rng.choice(["AM", "PM", "am", "pm"]),
                    "s": str(rng.choice([0, 1, 1700000000, 2147483647, 2147483648, 4294967295, 4294967296, 99999999999, rng.randrange(0, 2 ** 33)])), "%": "%",
                    "D": "%02d/%02d/%02d" % (mo, d, y % 100), "T": "%02d:%02d:%02d" % (h, mi, s), "F": "%04d-%02d-%02d" % (y, mo, d), "R": "%02d:%02d" % (h, mi)}[k])
    return "".join(out)


def mutate(t, rng):
    r = rng.random()
    if r < 0.55 or not t:
        return t
    i = rng.randrange(len(t))
    if r < 0.65:
        return t[:i] + rng.choice("0123456789") + t[i + 1:]
    if r < 0.72:
        return t[:i]
    if r < 0.79:
        return t[:i] + " " + t[i:]
    if r < 0.85:
        return t.swapcase()
    if r < 0.92:
        return t + rng.choice([" tail", "x", " ", "Z", ".5"])
    return t[:i] + rng.choice(["-", "/", ":", "T", "+", "a"]) + t[i + 1:]


def main(seed, n):
    rng = random.Random(seed)
    lib = pkg.load(util.HOSTSIM_SO)
    bad = 0
    for fmt in FORMATS:
        for strict in (True, False):
            for offset in (None, "+0530", "-0800"):
                if offset and ("%z" in fmt or rng.random() < 0.5):
                    continue
                kw = dict(name="t", format="regex", regex=r"^(?<time>.*)$", time_fmt=fmt, time_key="time", time_keep=True, time_strict=strict)
                if offset:
                    kw["time_offset"] = offset
                ctx, ref = pkg.Context(0, lib=lib), util.Ref()
                try:
                    rp = ref.parser(**kw)
                except RuntimeError:
                    continue
                p = ctx.parser(**kw)
                vals = [mutate(render(fmt, rng), rng).encode() for _ in range(n)]
                vals = [v for v in vals if b"\n" not in v]
                for v, (r, data, t) in zip(vals, p.do_batch(vals)):
                    rr, rdata, rt = ref.parser_do(rp, v)
                    if (r < 0) != (rr < 0) or (rr >= 0 and (data != rdata or t != (rt[0], rt[1]))):
                        bad += 1
                        if bad <= 25:
                            print("MISMATCH fmt=%r strict=%s offset=%s value=%r got=%s want=%s" % (fmt, strict, offset, v, (r, t), (rr, rt)))
    print("formats", len(FORMATS), "values per configuration", n, "mismatches", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]), int(sys.argv[2])) else 0)
