"""Parity cases shared by the CPU-emulation tests and the GPU tests: (name, parsers, filters, chunk factory)."""
import util

AP = dict(name="apache", format="regex", regex=util.APACHE_RX, time_fmt=util.APACHE_TIME_FMT, time_key="time")
NG = dict(name="nginx", format="regex", regex=util.NGINX_RX, time_fmt=util.APACHE_TIME_FMT, time_key="time")
AP_TYPES = dict(AP, types="code:integer size:integer", time_keep=True)
P = ("parser", [("Key_Name", "log"), ("Parser", "apache")])
PN = ("parser", [("Key_Name", "log"), ("Parser", "nginx")])


def apache_chunk(n=1500, seed=7):
    return util.chunk_from_lines(util.apache_lines(n, seed=seed))


def nginx_chunk(n=1500, seed=9):
    return util.chunk_from_lines(util.apache_lines(n, seed=seed, nginx=True))


def mixed_chunk():
    """Events with several keys, nested values, non-str types, a legacy-format event and duplicates."""
    ev = []
    s = util.mp_str
    for i in range(300):
        items = [(b"log", s(b"line %d GET /x" % i)), (b"level", s([b"info", b"warn", b"error"][i % 3])),
                 (b"n", bytes([i % 128])), (b"flag", b"\xc3" if i % 2 else b"\xc2"),
                 (b"kube", b"\x82" + s(b"pod") + s(b"p-%d" % i) + s(b"labels") + b"\x81" + s(b"app") + s(b"web")),
                 (b"arr", b"\x93\x01" + s(b"two") + b"\xcb" + b"\x40\x09\x21\xfb\x54\x44\x2d\x18"),
                 (b"debug", s(b"x" * (i % 40)))]
        if i % 7 == 0:
            items.append((b"level", s(b"dup")))
        if i % 11 == 0:
            items.append((b"Agent-X", s(b"zz")))
        ev.append(util.event(1700000000 + i, i, items))
    ev.append(b"\x92\xce\x65\x53\xf1\x00" + b"\x81" + s(b"log") + s(b"legacy GET"))      # [ts, body]
    ev.append(util.event(1700000999, 5, [(b"log", s(b"with meta"))], meta=b"\x81" + s(b"m") + b"\x01"))
    return b"".join(ev)


JS = dict(name="json", format="json", time_fmt="%d/%b/%Y:%H:%M:%S %z", time_key="time")
JS_PLAIN = dict(name="jsonp", format="json")
LT = dict(name="ltsv", format="ltsv", time_fmt="%d/%b/%Y:%H:%M:%S %z", time_key="time", types="status:integer size:integer")
LF = dict(name="logfmt", format="logfmt", time_fmt="%Y-%m-%dT%H:%M:%SZ", time_key="ts", time_keep=True)
PJ = ("parser", [("Key_Name", "log"), ("Parser", "json")])

JSON_EDGE = [
    b'{"a":1,"b":"x","c":[1,2,{"d":null}],"e":true,"f":false,"g":-5,"h":1.5,"i":1e3,"time":"28/Jul/2006:10:27:10 -0300"}',
    b'  {"k":"v"}  ', b'{"a":1} trailing', b'{"a":1} {"b":2}', b'{"a":1}5', b'[1,2]', b'"str"', b'{}', b'{"a":}', b'{"a":1,}',
    b'{"a" 1}', b'{"big":18446744073709551615,"bigger":18446744073709551616,"neg":-9223372036854775808,"neg2":-9223372036854775809,"z":-0,"zf":-0.0}',
    b'{"s":"esc \\" \\\\ \\/ \\b \\f \\n \\r \\t \\u0041 \\u00e9 \\u20ac \\ud83d\\ude00 end"}',
    b'{"s":"bad \\ud83d x","t":"\\udc00","u":"\\u12","v":"\\ud83d\\u0041","w":"\\uZZ"}', b'{"s":"\\x41"}',
    b'{"s":"ctl\x01\x1f","u":"\xff\xfe ok \xc3\xa9"}',
    b'{"f":[0.1,0.2,0.3,3.14159,2.718281828459045,1e-7,1.7976931348623157e308,4.9e-324,5e-324,0.000001,1e22,1e23,9007199254740993,0.1e1]}',
    b'{"inf":1e999}', b'{"a":01}', b'{"a":1.}', b'{"a":.5}', b'{"a":1e}', b'{"a":+1}', b'{"a":tru}',
    b'{"nested":' + b'[' * 30 + b']' * 30 + b'}', b'{"nested":' + b'[' * 31 + b']' * 31 + b'}', b'{"nested":' + b'[' * 32 + b']' * 32 + b'}',
    b'{"time":123}', b'{"time":"bogus"}', b'{"a":"x","a":"y","time":"28/Jul/2006:10:27:10 -0300","time":"dup"}', b'', b'   ',
    b'{"a":"unterminated', b'{"a":{"b":{"c":{"d":[1,[2,[3]]]}}}}', b'\t{"tab":1}\n', b'{"a":1}\x00', b'{"a"\n:\r1 ,\t"b" : 2 }',
    b'{"u":"\\u0000x"}', b'{"e":"\\u00"}', b'{"k\\u0041":1,"\\n":2}', b'{"x":"\\ud3d\\ude00b\\u004\\":1}',
]


def json_chunk(n=1200, seed=5):
    return util.chunk_from_lines(util.json_lines(n, seed=seed))


def wide_apache_chunk():
    """apache lines next to 8 more fields: parsed (11) + reserved (8) fields exceed the 16-field list the
    evaluation pass caches for the emission pass, which then re-runs the chain for these records"""
    lines = util.apache_lines(300, seed=9)
    return b"".join(util.event(1700000000 + i, 0, [(b"log", util.mp_str(l))] + [(b"x%d" % j, util.mp_str(b"v%d" % j)) for j in range(8)])
                    for i, l in enumerate(lines))


def wide_json_chunk():
    """JSON documents with 20 members (same reason), some with nested values and escapes"""
    out = []
    for i in range(200):
        items = ['"k%02d":%s' % (j, ['"v%d"' % j, str(j * i), "true", "null", '{"a":[1,2,{"b":"c"}]}', '"esc\\n\"q\""', "%d.5" % j][(i + j) % 7]) for j in range(20)]
        items.insert(i % 20, '"level":"%s"' % ["warn", "info", "error"][i % 3])
        out.append(("{" + ",".join(items) + "}").encode())
    return util.chunk_from_lines(out)


def json_edge_chunk():
    return util.chunk_from_lines(JSON_EDGE)


def ltsv_chunk():
    return util.chunk_from_lines(util.ltsv_lines(400) + [b"a:1\tb:\t:x\tc:3", b"nolabel", b"k:v\r\nrest:1", b"time:bogus\ta:1", b"", b"a:1\t\tb:2"])


def logfmt_chunk():
    return util.chunk_from_lines(util.logfmt_lines(400) + [b"a=1 b c=", b'x="unterminated', b"=novalue k=v", b"  ", b'q="a b" r=s\nnext=1'])


def logfmt_escape_chunk():
    """quoted logfmt values with backslash escapes: decoded like flb_unescape_string_utf8() + strlen()"""
    bs = "\\"
    vals = [bs + "n", bs + "t tab", "say " + bs + '"hi' + bs + '"', bs + bs, "a" + bs + "/b", bs + "x41" + bs + "x4" + bs + "xZ", bs + "u00e9" + bs + "u20ac",
            bs + "ud83d" + bs + "ude00", bs + "ud83d alone", bs + "udc00", bs + "u12", bs + "u", bs + "U0001F600", bs + "101" + bs + "7", "cut" + bs + "0here",
            bs + "q" + bs + "é", "é 日本", "tail" + bs, bs + "v" + bs + "a" + bs + "b" + bs + "f" + bs + "r", bs + "400"]
    lines = [('level=info n=%d msg="%s" path=/x other="%s" flag' % (i, v, vals[(i * 7) % len(vals)])).encode("utf-8") for i, v in enumerate(vals * 5)]
    return util.chunk_from_lines(lines)


# what atof() / glibc strtod makes of a captured text (`Types v:float`, flb_parser.c: flb_parser_typecast)
STRTOD_TEXTS = ["", "-", "+", ".", "-.", "e5", ".e5", "5.", ".5", "+.5e1", "-5.e-1", "1e", "1e+", "1e-", "1ex", "1e+x", "1.5abc", "1..5",
                "  12.5", "\t-3.25", "\v\f\r 7", "00012.50", "0000.0001", "-0", "-0.0", "0e999999", "1e400", "-1e400", "1e-400", "-1e-400",
                "1e309", "1.7976931348623157e308", "1.7976931348623159e308", "4.9e-324", "2.4e-324", "2.5e-324", "2.2250738585072011e-308",
                "2.2250738585072014e-308", "inf", "-inf", "INF", "Infinity", "infinit", "+infinityx", "nan", "-nan", "NaN", "nanx", "in", "na",
                "1e99999999999", "1e-99999999999", "0." + "0" * 400 + "1e401", "1" + "0" * 400 + "e-400", "123456789012345678901234567890",
                "9007199254740993", "9007199254740992.5", "9007199254740993e0", "0.1", "0.3", "1e23", "8.5e22", "1e22", "1e-22", "5e-23",
                "3.14 15", "1,5", "1_000", "12e3.5", "- 5", "+-5", "\uff11"]


def float_types_chunk():
    """captures cast with `Types a:float b:float`: strtod's syntax, and decimals on / next to the midpoint of two doubles"""
    import os
    import random
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
    import floatfuzz
    texts = STRTOD_TEXTS + [t for t in floatfuzz.cases(random.Random(77), 30) if len(t) < 240]
    lines = [("%s|%s|%d" % (t, texts[(i * 7 + 3) % len(texts)], i)).encode("utf-8") for i, t in enumerate(texts)]
    return util.chunk_from_lines(lines)


def docker_chunk():
    """docker json-file lines whose `log` is itself JSON: the inner text only exists once the outer string is decoded"""
    import json as _json
    lines = []
    for i, l in enumerate(util.json_lines(300, 23)):
        inner = l.decode() if i % 3 else "plain text %d" % i
        lines.append(_json.dumps({"log": inner + "\n" if i % 5 == 0 else inner, "stream": "stdout" if i % 2 else "stderr", "time": "2023-05-06T07:08:09.%03dZ" % (i % 1000)}).encode())
    return util.chunk_from_lines(lines)


def dup_key_chunk():
    S = util.mp_str
    t1 = S(b"time=2023-05-06T07:08:09.5Z a=1")
    recs = [[(b"k1", t1), (b"k1", S(b"v"))], [(b"k1", S(b"v")), (b"k1", t1)], [(b"k1", t1), (b"k1", S(b"{"))],
            [(b"k1", t1), (b"x", S(b"y")), (b"k1", S(b'{"time":"2024-01-01T00:00:00.0Z","b":2}')), (b"k1", S(b"c=3"))],
            [(b"k1", S(b'{"time":"2024-01-01T00:00:00.25Z"}')), (b"k1", S(b'{"q":1}'))]] * 4
    return b"".join(util.event(1700000000 + i, 7, f) for i, f in enumerate(recs))


def tricky_ts_chunk():
    """Timestamps whose bytes frame as complete legacy events ([uint32, {}]) inside real records:
    the record index has to rule those candidates out (sec = 0x655492ce -> `92 ce 00 00 xx xx 80`)."""
    lines = util.apache_lines(400, seed=21)
    ev = []
    for i, l in enumerate(lines):
        sec = 0x655492ce if i % 5 == 0 else (0x6554cc92 if i % 7 == 0 else 1700000000 + i)
        ev.append(util.event(sec, i % 1000, [(b"log", util.mp_str(l)), (b"n", b"\x92\xcc\x05\x80") if i % 9 == 0 else (b"n", b"\x01")]))
    return b"".join(ev)


CASES = [
    ("apache_parser", [AP], [P], apache_chunk),
    ("apache_parser_types_keep", [AP_TYPES], [P], apache_chunk),
    ("north_star_chain", [AP], [P, ("grep", [("Regex", "method ^(GET|POST)$")]),
                                ("modify", [("Add", "env prod"), ("Rename", "code status"), ("Remove", "agent")])], apache_chunk),
    ("parser_modify_recmod", [AP], [P, ("modify", [("Add", "env prod"), ("Remove", "agent"), ("Rename", "code status")]),
                                    ("record_modifier", [("Record", "hostname node-1"), ("Remove_key", "referer")])], apache_chunk),
    ("nginx_recmod", [NG], [PN, ("record_modifier", [("Record", "hostname node-1"), ("Remove_key", "agent")])], nginx_chunk),
    ("parser_reserve_preserve", [AP], [("parser", [("Key_Name", "log"), ("Parser", "apache"), ("Reserve_Data", "On"), ("Preserve_Key", "On")])], apache_chunk),
    ("parser_preserve", [AP], [("parser", [("Key_Name", "log"), ("Parser", "apache"), ("Preserve_Key", "On")])], apache_chunk),
    ("parser_reserve", [AP], [("parser", [("Key_Name", "log"), ("Parser", "apache"), ("Reserve_Data", "On")])], mixed_chunk),
    ("parser_ra_key", [AP], [("parser", [("Key_Name", "$log"), ("Parser", "apache"), ("Reserve_Data", "On")])], apache_chunk),
    ("types_float_strtod", [dict(name="fl", format="regex", regex=r"^(?<a>[^|]*)\|(?<b>[^|]*)\|(?<n>\d+)$", types="a:float b:float n:integer")],
     [("parser", [("Key_Name", "log"), ("Parser", "fl")]), ("grep", [("Exclude", "n ^7$")])], float_types_chunk),
    # a parser whose input was made by an earlier filter of the same chain: the fused form refuses, the chain then runs filter by filter on the device
    ("docker_then_inner_json", [dict(name="docker", format="json", time_key="time", time_fmt="%Y-%m-%dT%H:%M:%S.%LZ"), JS],
     [("parser", [("Key_Name", "log"), ("Parser", "docker"), ("Reserve_Data", "On")]), ("parser", [("Key_Name", "log"), ("Parser", "json"), ("Reserve_Data", "On"), ("Preserve_Key", "On")]),
      ("grep", [("Exclude", "stream stderr")])], docker_chunk),
    ("set_then_parse", [AP], [("modify", [("Set", "log \"GET /x HTTP/1.1\"")]), ("parser", [("Key_Name", "log"), ("Parser", "apache"), ("Reserve_Data", "On")])], mixed_chunk),
    # several keys named like Key_Name: each is parsed in turn, the last one decides the body, the last non-zero time stays
    ("parser_duplicate_key_names", [dict(name="jst", format="json", time_key="time", time_fmt="%Y-%m-%dT%H:%M:%S.%LZ"),
                                    dict(name="lft", format="logfmt", time_key="time", time_fmt="%Y-%m-%dT%H:%M:%S.%LZ", types="n:integer")],
     [("parser", [("Key_Name", "k1"), ("Parser", "jst"), ("Parser", "lft")])],
     dup_key_chunk),
    ("tricky_timestamps_parser", [AP], [P], tricky_ts_chunk),
    ("tricky_timestamps_grep", [], [("grep", [("Regex", "log GET")])], tricky_ts_chunk),
    ("json_parser", [JS], [PJ], json_chunk),
    ("json_parser_edge", [JS], [PJ], json_edge_chunk),
    ("json_parser_edge_plain_reserve", [JS_PLAIN], [("parser", [("Key_Name", "log"), ("Parser", "jsonp"), ("Reserve_Data", "On")])], json_edge_chunk),
    ("json_chain_config1", [JS], [PJ, ("grep", [("Regex", "level ^(warn|error)$")]),
                                 ("modify", [("Add", "env prod"), ("Rename", "msg message"), ("Remove", "debug")])], json_chunk),
    ("json_chain_nested_grep", [JS], [PJ, ("grep", [("Regex", "$kubernetes['labels']['app'] .")]),
                                     ("record_modifier", [("Remove_key", "trace_id"), ("Record", "cluster c1")])], json_chunk),
    ("wide_records_apache", [AP], [("parser", [("Key_Name", "log"), ("Parser", "apache"), ("Reserve_Data", "On")]),
                                   ("modify", [("Add", "env prod"), ("Remove", "x3"), ("Rename", "x5 y5")])], wide_apache_chunk),
    ("wide_records_json", [JS], [PJ, ("grep", [("Regex", "level ^(warn|error)$")]), ("modify", [("Add", "env prod"), ("Remove", "k07")])],
     wide_json_chunk),
    ("ltsv_parser", [LT], [("parser", [("Key_Name", "log"), ("Parser", "ltsv")])], ltsv_chunk),
    ("logfmt_parser", [LF], [("parser", [("Key_Name", "log"), ("Parser", "logfmt")]), ("grep", [("Exclude", "level debug")])], logfmt_chunk),
    ("multi_parser_fallthrough", [AP, JS, LF], [("parser", [("Key_Name", "log"), ("Parser", "json"), ("Parser", "apache"), ("Parser", "logfmt")])], lambda: util.chunk_from_lines(util.json_lines(100, 3) + util.apache_lines(100, 4) + util.logfmt_lines(100, 5))),
    ("logfmt_escapes", [LF], [("parser", [("Key_Name", "log"), ("Parser", "logfmt")]), ("modify", [("Rename", "msg message")])], logfmt_escape_chunk),
    ("grep_regex", [], [("grep", [("Regex", "log GET")])], apache_chunk),
    ("grep_exclude", [], [("grep", [("Exclude", "log HTTP")])], apache_chunk),
    ("grep_keep_all_notouch", [], [("grep", [("Regex", "log .")])], apache_chunk),
    ("grep_drop_all", [], [("grep", [("Regex", "log ^nomatch$")])], apache_chunk),
    ("grep_and", [], [("grep", [("Logical_Op", "and"), ("Regex", "log GET"), ("Regex", "level ^(warn|error)$")])], mixed_chunk),
    ("grep_or", [], [("grep", [("Logical_Op", "or"), ("Exclude", "level info"), ("Exclude", "log 7")])], mixed_chunk),
    ("grep_legacy_mixed", [], [("grep", [("Exclude", "level dup"), ("Regex", "log 1"), ("Exclude", "level error")])], mixed_chunk),
    ("grep_nested", [], [("grep", [("Regex", "$kube['labels']['app'] ^web$"), ("Exclude", "$kube['pod'] p-1.$")])], mixed_chunk),
    ("grep_array", [], [("grep", [("Regex", "$arr[1] tw")])], mixed_chunk),
    ("modify_rules", [], [("modify", [("Set", "level fixed"), ("Copy", "log raw"), ("Hard_copy", "n flag"), ("Move_to_start", "ra"),
                                      ("Remove_wildcard", "de"), ("Remove_regex", "^Agent"), ("Hard_rename", "arr kube"),
                                      ("Move_to_end", "lo"), ("Add", "log no"), ("Add", "extra 1")])], mixed_chunk),
    ("modify_conditions", [], [("modify", [("Condition", "Key_exists debug"), ("Condition", "Key_value_matches level ^(warn|error)$"),
                                           ("Condition", "Key_does_not_exist nope"), ("Condition", "A_key_matches ^ku"),
                                           ("Condition", "Key_value_does_not_equal level info"),
                                           ("Condition", "Matching_keys_have_matching_values ^l [a-z]"),
                                           ("Add", "matched yes")])], mixed_chunk),
    ("modify_bool_condition", [], [("modify", [("Condition", "Key_value_matches flag true"), ("Rename", "flag FLAG")])], mixed_chunk),
    # three words: the reference leaves the calloc()ed rule type (RENAME) in place, first word -> last word
    ("modify_three_word_rules", [], [("modify", [("Copy", "level x renamed"), ("Remove", "log to raw"), ("Set", "n is count")])], mixed_chunk),
    ("modify_notouch", [], [("modify", [("Remove", "absent"), ("Rename", "nope x")])], mixed_chunk),
    ("recmod_remove_some", [], [("record_modifier", [("Remove_key", "agent-x"), ("Remove_key", "de*")])], mixed_chunk),
    ("recmod_allow", [], [("record_modifier", [("Allowlist_key", "LOG"), ("Whitelist_key", "lev*")])], mixed_chunk),
    ("recmod_notouch", [], [("record_modifier", [("Remove_key", "nothere")])], mixed_chunk),
    ("recmod_drop_all", [], [("record_modifier", [("Allowlist_key", "nothere")])], mixed_chunk),
    ("recmod_then_modify", [], [("record_modifier", [("Record", "a b")]), ("modify", [("Add", "c d")]),
                                ("grep", [("Exclude", "level info")])], mixed_chunk),
    ("recmod_noop_then_grep", [], [("record_modifier", [("Remove_key", "nothere")]), ("grep", [("Exclude", "level info")])], mixed_chunk),
]
