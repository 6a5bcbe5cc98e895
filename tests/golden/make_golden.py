"""Regenerates tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libflbref.so, built
from /root/reference by oracle/refshim/Makefile).  Run from the repo root:

    python tests/golden/make_golden.py

Three fixture files:
  time_vectors.json ... the reference's own time table (tests/internal/parser.c:61-105 with the
                        parser formats of tests/internal/data/parser/regex.conf) + extra formats,
                        each resolved by the reference's flb_parser_do()
  regex_vectors.json .. patterns x subjects with Onigmo's capture offsets (flb_regex_do)
  chain_vectors.json .. small chunks through the reference's filter chain (flb_filter_do), hex
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))

import cases   # noqa: E402
import rxdiff  # noqa: E402
import util    # noqa: E402

# (time format, time string, Time_Offset or None, expected epoch from the reference's table or None, frac)
TIME = [
    ("%b %d %H:%M:%S", "Feb 16 04:06:58", "-0600", None, 0),
    ("%b %d %H:%M:%S.%L", "Feb 16 04:06:58.1234", "-0600", None, 0.1234),
    ("%b %d %H:%M:%S,%L", "Feb 16 04:06:58,1234", "-0600", None, 0.1234),
    ("%b %d %H:%M:%S %z", "Feb 16 04:06:58 -0600", None, None, 0),
    ("%b %d %H:%M:%S.%L %z", "Feb 16 04:06:58.1234 -0600", None, None, 0.1234),
    ("%b %d %H:%M:%S,%L %z", "Feb 16 04:06:58,1234 -0600", None, None, 0.1234),
    ("%m/%d/%Y %H:%M:%S %z", "07/17/2017 20:17:03 +0000", None, 1500322623, 0),
    ("%m/%d/%Y %H:%M:%S %z", "07/18/2017 01:47:03 +0530", None, 1500322623, 0),
    ("%m/%d/%Y %H:%M:%S %z", "07/18/2017 05:17:03 +0900", None, 1500322623, 0),
    ("%m/%d/%Y %H:%M:%S %z", "07/17/2017 22:17:03 +0200", None, 1500322623, 0),
    ("%m/%d/%Y %H:%M:%S.%L %z", "07/17/2017 22:17:03.1 +0200", None, 1500322623, 0.1),
    ("%m/%d/%Y %H:%M:%S,%L %z", "07/17/2017 22:17:03,1 +0200", None, 1500322623, 0.1),
    ("%m/%d/%Y %H:%M:%S %z", "07/18/2017 01:47:03 +05:30", None, 1500322623, 0),
    ("%m/%d/%Y %H:%M:%S.%L %z", "07/17/2017 22:17:03.1 +02:00", None, 1500322623, 0.1),
    ("%m/%d/%Y %H:%M:%S,%L %z", "07/17/2017 22:17:03,1 +02:00", None, 1500322623, 0.1),
    ("%m/%d/%Y %H:%M:%S:%L %z", "07/17/2017 22:17:03:1 +02:00", None, 1500322623, 0.1),
    ("%m/%d/%Y %H:%M:%S", "07/18/2017 01:47:03", "+0530", 1500322623, 0),
    ("%m/%d/%Y %H:%M:%S", "07/18/2017 05:17:03", "+0900", 1500322623, 0),
    ("%m/%d/%Y %H:%M:%S", "07/17/2017 22:17:03", "+0200", 1500322623, 0),
    ("%m/%d/%Y %H:%M:%S.%L", "07/17/2017 22:17:03.1", "+0200", 1500322623, 0.1),
    ("%m/%d/%Y %H:%M:%S,%L", "07/17/2017 22:17:03,1", "+0200", 1500322623, 0.1),
    ("%m/%d/%Y %H:%M:%S", "07/17/2017 20:17:03", None, 1500322623, 0),
    ("%m/%d/%Y %H:%M:%SZ", "07/17/2017 20:17:03Z", None, 1500322623, 0),
    ("%m/%d/%Y %H:%M:%S.%LZ", "07/17/2017 20:17:03.1234Z", None, 1500322623, 0.1234),
    ("%m/%d/%Y %H:%M:%S,%LZ", "07/17/2017 20:17:03,1234Z", None, 1500322623, 0.1234),
    ("%a %b %d %H:%M:%S.%L %Y", "Fri Jul 17 20:17:03.1234 2017", None, 1500322623, 0.1234),
    # beyond the reference's table
    ("%d/%b/%Y:%H:%M:%S %z", "28/Jul/2006:10:27:10 -0300", None, 1154093230, 0),
    ("%Y-%m-%dT%H:%M:%S.%L", "2023-11-14T22:13:20.123456789", None, None, None),
    ("%Y-%m-%dT%H:%M:%S.%L%z", "2023-11-14T22:13:20.5+01:00", None, None, None),
    ("%Y-%m-%dT%H:%M:%S.%LZ", "2023-11-14T22:13:20.000000001Z", None, None, None),
    ("%s", "1700000000", None, 1700000000, 0),
    ("%s.%L", "1700000000.25", None, 1700000000, 0.25),
    ("%y%m%d %H%M%S", "231114 221320", None, None, None),
    ("%F %T", "2020-02-29 23:59:60", None, None, None),
    ("%D %R", "02/29/20 12:34", None, None, None),
    ("%b %e %Y %I:%M:%S %p", "Mar  5 2021 11:07:09 PM", None, None, None),
    ("%A, %d %B %Y %H:%M:%S %Z", "Friday, 17 July 2020 10:00:00 CEST", None, None, None),
    ("%Y-%m-%d %H:%M:%S %Z", "2020-07-17 10:00:00 PST", None, None, None),
    ("%Y %j %H", "2021 045 07", None, None, None),
    ("%d/%b/%Y:%H:%M:%S %z", "31/Feb/2021:10:27:10 +0000", None, None, None),
    ("%d/%b/%Y:%H:%M:%S %z", "28/Jul/2006:10:27:10", None, None, None),
    ("%d/%b/%Y:%H:%M:%S %z", "bogus", None, None, None),
    ("%d/%b/%Y:%H:%M:%S %z", "28/Jul/2006:10:27:10 -0300 trailing", None, None, None),
    ("%H:%M:%S", "25:00:00", None, None, None),
]

REGEX = [
    (util.APACHE_RX, None),
    (util.NGINX_RX, None),
    (r'^\[[^ ]* (?<time>[^\]]*)\] \[(?<level>[^\]]*)\](?: \[pid (?<pid>[^\]]*)\])?( \[client (?<client>[^\]]*)\])? (?<message>.*)$',
     [b"[Wed Oct 11 14:32:52 2000] [error] [pid 3] [client 127.0.0.1] client denied", b"[x y] [notice] hello", b"nope"]),
    (r'^(?<time>[^ ]+) (?<stream>stdout|stderr) (?<logtag>[^ ]*) (?<log>.*)$',
     [b"2023-01-01T00:00:00.000000000Z stdout F hello world", b"2023-01-01T00:00:00Z stderr P ", b"x stdin F y"]),
    (r'^\<(?<pri>[0-9]+)\>(?<time>[^ ]* {1,2}[^ ]* [^ ]*) (?<host>[^ ]*) (?<ident>[a-zA-Z0-9_\/\.\-]*)(?:\[(?<pid>[0-9]+)\])?(?:[^\:]*\:)? *(?<message>.*)$',
     [b"<34>Oct 11 22:14:15 mymachine su[230]: 'su root' failed", b"<13>Feb  5 17:32:18 10.0.0.99 app: hi", b"<>x"]),
    (r'/^(?<k>[a-z]+)=(?<v>\d+)$/i', [b"Abc=12", b"ABC=x", b"abc=7\n"]),
    (r'(?<a>a+?)(?<b>b*)(?<c>c|$)', [b"aaabbc", b"xxaab", b"b", b""]),
    (r'^(?<w>\w+)\s+(?<rest>(?:\S+\s*){1,3})', [b"alpha beta gamma delta epsilon", b"one two", b"solo"]),
    # tests/internal/regex.c: test_basic, test_uri ("/pattern/option" must not be misread), /i, /m, /x, /ix
    (r'/(?<str>[a-z]+) (?<num>\d+) (?<time>\d{4}/\d{2}/\d{2})/', [b"string 1234 2022/10/24"]),
    (r'/uri/(?<middle>[a-z]+)/hoge', [b"/uri/is/hoge"]),
    (r'/(?<str>[a-z]+)/i', [b"STRING"]),
    (r'/(?<full_str>.+)/m', [b"string\n1234\nstring"]),
    (r'/(?<pi>\d  \. 14)/x', [b"3.14"]),
    (r'/(?<full_str>\d  \. 14PI)/ix', [b"3.14pi"]),
]


def main():
    ref = util.Ref()
    out_time = []
    for i, (fmt, s, off, epoch, frac) in enumerate(TIME):
        p = ref.parser("t%d" % i, "regex", r"^(?<time>.+)$", time_fmt=fmt, time_key="time", time_offset=off, time_keep=True)
        r, data, (sec, nsec) = ref.parser_do(p, s.encode())
        no_year = not any(x in fmt for x in ("%Y", "%y", "%s", "%F", "%D"))
        out_time.append({"fmt": fmt, "str": s, "offset": off, "table_epoch": epoch, "table_frac": frac,
                         "ref_ret": r, "ref_sec": None if no_year else sec, "ref_nsec": nsec, "no_year": no_year,
                         "ref_map_hex": data.hex() if data else None})
    json.dump(out_time, open(os.path.join(HERE, "time_vectors.json"), "w"), indent=1)

    out_rx = []
    ap = util.apache_lines(40, seed=3) + [b"", b"x", b'1 2 3 [t] "GET" 200 1', b'1 2 3 [t] "GET  /a  b" 200 -\n']
    ng = util.apache_lines(20, seed=4, nginx=True)
    for pat, subs in REGEX:
        if subs is None:
            subs = ng if pat == util.NGINX_RX else ap
        re = rxdiff.REF.flbref_regex_create(pat.encode())
        rows = []
        for s in subs:
            m = rxdiff.ref_search(re, s)
            rows.append({"s": s.hex(), "m": None if m is None else [list(x) for x in m]})
        out_rx.append({"pattern": pat, "cases": rows})
    json.dump(out_rx, open(os.path.join(HERE, "regex_vectors.json"), "w"), indent=0)

    out_chain = []
    small = {"apache_chunk": lambda: util.chunk_from_lines(util.apache_lines(60, seed=11)),
             "nginx_chunk": lambda: util.chunk_from_lines(util.apache_lines(60, seed=12, nginx=True)),
             "tricky_ts_chunk": lambda: cases.tricky_ts_chunk()[:30000 * 0 + len(b"".join(cases.tricky_ts_chunk()[o:o + l] for o, l in util.split_records(cases.tricky_ts_chunk())[:60]))],
             "mixed_chunk": lambda: cases.mixed_chunk()[:20000]}
    small.update({"logfmt_escape_chunk": cases.logfmt_escape_chunk, "float_types_chunk": cases.float_types_chunk, "dup_key_chunk": cases.dup_key_chunk,
                  "docker_chunk": lambda: b"".join(cases.docker_chunk()[o:o + l] for o, l in util.split_records(cases.docker_chunk())[:40])})
    small.update({"wide_apache_chunk": lambda: b"".join(cases.wide_apache_chunk()[o:o + l] for o, l in util.split_records(cases.wide_apache_chunk())[:40]),
                  "wide_json_chunk": lambda: b"".join(cases.wide_json_chunk()[o:o + l] for o, l in util.split_records(cases.wide_json_chunk())[:40])})
    small.update({"json_chunk": lambda: cases.json_chunk(60), "json_edge_chunk": cases.json_edge_chunk,
                  "ltsv_chunk": lambda: util.chunk_from_lines(util.ltsv_lines(40) + [b"a:1\tb:\t:x\tc:3", b"nolabel"]),
                  "logfmt_chunk": lambda: util.chunk_from_lines(util.logfmt_lines(40) + [b"a=1 b c=", b'x="unterminated']),
                  "<lambda>": lambda: util.chunk_from_lines(util.json_lines(20, 3) + util.apache_lines(20, 4) + util.logfmt_lines(20, 5))})
    for name, parsers, filters, mk in cases.CASES:
        chunk = small[mk.__name__]()
        if mk.__name__ == "mixed_chunk":
            recs = util.split_records(cases.mixed_chunk())
            keep = recs[:40] + recs[-2:]
            full = cases.mixed_chunk()
            chunk = b"".join(full[o:o + l] for o, l in keep)
        r = util.Ref()
        for kw in parsers:
            r.parser(**kw)
        for p, props in filters:
            r.filter(p, props)
        ret, out = r.chain_do(chunk)
        out_chain.append({"name": name, "parsers": parsers, "filters": filters, "in_hex": chunk.hex(), "ret": ret,
                          "out_hex": None if out is None else out.hex()})
    json.dump(out_chain, open(os.path.join(HERE, "chain_vectors.json"), "w"))
    print("wrote %d time, %d regex patterns, %d chain vectors" % (len(out_time), len(out_rx), len(out_chain)))

    # filter_log_to_metrics: result of the chain plus the text dump of the plugin's cmetrics context
    import ctypes as C
    import re as _re
    import l2m_cases
    out_l2m = []
    for name, parsers, filters, mk, k in l2m_cases.L2M_CASES:
        full = mk()
        chunk = b"".join(full[o:o + l] for o, l in util.split_records(full)[:80])
        r = util.Ref()
        for kw in parsers:
            r.parser(**kw)
        fs = [r.filter(p, props) for p, props in filters]
        ret, out = r.chain_do(chunk)
        r.L.flbref_l2m_cmt_text.restype = C.c_void_p
        r.L.flbref_l2m_cmt_text.argtypes = [C.c_void_p]
        text = _re.sub(r"^\S+Z ", "", C.string_at(r.L.flbref_l2m_cmt_text(fs[k])).decode(errors="replace"), flags=_re.M)
        out_l2m.append({"name": name, "parsers": parsers, "filters": filters, "k": k, "in_hex": chunk.hex(), "ret": ret,
                        "out_hex": None if out is None else out.hex(), "text": text})
    json.dump(out_l2m, open(os.path.join(HERE, "l2m_vectors.json"), "w"))
    print("wrote %d log_to_metrics vectors" % len(out_l2m))


if __name__ == "__main__":
    main()
