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
    ("tricky_timestamps_parser", [AP], [P], tricky_ts_chunk),
    ("tricky_timestamps_grep", [], [("grep", [("Regex", "log GET")])], tricky_ts_chunk),
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
    ("modify_notouch", [], [("modify", [("Remove", "absent"), ("Rename", "nope x")])], mixed_chunk),
    ("recmod_remove_some", [], [("record_modifier", [("Remove_key", "agent-x"), ("Remove_key", "de*")])], mixed_chunk),
    ("recmod_allow", [], [("record_modifier", [("Allowlist_key", "LOG"), ("Whitelist_key", "lev*")])], mixed_chunk),
    ("recmod_notouch", [], [("record_modifier", [("Remove_key", "nothere")])], mixed_chunk),
    ("recmod_drop_all", [], [("record_modifier", [("Allowlist_key", "nothere")])], mixed_chunk),
    ("recmod_then_modify", [], [("record_modifier", [("Record", "a b")]), ("modify", [("Add", "c d")]),
                                ("grep", [("Exclude", "level info")])], mixed_chunk),
    ("recmod_noop_then_grep", [], [("record_modifier", [("Remove_key", "nothere")]), ("grep", [("Exclude", "level info")])], mixed_chunk),
]
