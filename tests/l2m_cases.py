"""filter_log_to_metrics cases (plugins/filter_log_to_metrics/log_to_metrics.c; the reference's own
runtime tests are tests/runtime/filter_log_to_metrics.c: counter, histogram, labels, regex gate,
discard_logs).  Each case: (name, parsers, filters, chunk maker, index of the l2m filter)."""
import random
import struct

import util


def _mp(v):
    if isinstance(v, bool):
        return b"\xc3" if v else b"\xc2"
    if v is None:
        return b"\xc0"
    if isinstance(v, int):
        if 0 <= v < 128:
            return bytes([v])
        if -32 <= v < 0:
            return struct.pack("b", v)
        if 0 <= v < 1 << 32:
            return b"\xce" + struct.pack(">I", v)
        return b"\xd3" + struct.pack(">q", v)
    if isinstance(v, float):
        return b"\xcb" + struct.pack(">d", v)
    if isinstance(v, tuple) and v[0] == "f32":
        return b"\xca" + struct.pack(">f", v[1])
    if isinstance(v, bytes):
        return util.mp_str(v)
    if isinstance(v, dict):
        return util.mp_map_hdr(len(v)) + b"".join(util.mp_str(k) + _mp(x) for k, x in v.items())
    raise TypeError(v)


def events(n, seed, extra=None):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        rec = [
            (b"message", rng.choice([b"ok request", b"error: disk", b"ok cached", b"warn slow", b""])),
            (b"color", rng.choice([b"red", b"green", b"blue", b"a" * 300, b"nul\x00tail"])),
            (b"code", rng.choice([200, 404, 500, -7, 1 << 40])),
            (b"duration", rng.choice([0.001, 0.25, 0.5, 3, b"1.5", 12, 2.5, b"0.004", b"7"])),
            (b"kubernetes", {b"pod_name": rng.choice([b"web-1", b"web-2"]), b"labels": {b"app": b"shop"}}),
        ]
        if rng.random() < 0.1:
            rec = [kv for kv in rec if kv[0] != b"color"]                  # missing label -> ""
        if rng.random() < 0.1:
            rec = [kv for kv in rec if kv[0] != b"duration"]               # missing value -> no observation
        if rng.random() < 0.05:
            rec.append((b"code", True))                                    # unsupported type -> ""
        if extra:
            rec += extra(rng)
        out.append(util.event(1700000000 + i, i % 1000, [(k, _mp(v)) for k, v in rec]))
    return b"".join(out)


FLOATS = [0.5, 1 / 128, 3 / 128, 5 / 128, 2.5, 1e300, -1e-7, 5e-7, 4.9999995e-7, 5.0000005e-7, 1e22, 123456.789, -0.0, 0.0, float("inf"),
          float("-inf"), float("nan"), 1.7976931348623157e308, 5e-324, 0.1, 1e-6, 9.9999995e-7, 0.9999995, 0.99999949999, 1.0000005,
          9007199254740993.0, 4503599627370497.5, 2.0 ** 73, 2.0 ** -20, 2.0 ** -21, 3 * 2.0 ** -22, 1e15 + 0.3, 999999.9999995, 1e21,
          ("f32", 0.1), ("f32", 1 / 128), ("f32", -3.4028234663852886e38), ("f32", 1e-45)]


def float_label(rng):
    r = rng.random()
    if r < 0.5:
        return [(b"ratio", rng.choice(FLOATS))]
    if r < 0.8:
        return [(b"ratio", struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0])]     # any bit pattern
    return [(b"ratio", rng.randrange(-(1 << 30), 1 << 30) / (1 << rng.randrange(0, 40)))]      # dyadic: exact ties happen


def grouped_events(seed):
    rng = random.Random(seed)
    out = []
    for g in range(6):
        out.append(util.event(0xffffffff, 0, [(b"color", util.mp_str(b"group-%d" % (g % 2))), (b"duration", _mp(3))]))
        out.append(events(rng.randrange(5, 40), seed * 10 + g))
        if g % 3 == 0:
            out.append(util.event(0x80000005, 1, [(b"color", util.mp_str(b"negative")), (b"duration", _mp(0.25))]))
        out.append(util.event(0xfffffffe, 0, []))
    return b"".join(out)


BASE = [("metric_name", "reqs"), ("metric_description", "requests"), ("tag", "metrics")]

L2M_CASES = [
    ("counter_no_labels", [], [("log_to_metrics", BASE + [("metric_mode", "counter")])], lambda: events(500, 1), 0),
    ("counter_labels", [], [("log_to_metrics", BASE + [("metric_mode", "counter"), ("label_field", "color"),
                                                       ("add_label", "status $code"), ("add_label", "pod $kubernetes['pod_name']")])],
     lambda: events(800, 2), 0),
    ("counter_regex_gate", [], [("log_to_metrics", BASE + [("regex", "message ^ok"), ("exclude", "color blue"),
                                                           ("label_field", "color")])], lambda: events(800, 3), 0),
    ("counter_exclude_first", [], [("log_to_metrics", BASE + [("exclude", "message error"), ("regex", "message ^(ok|warn)"),
                                                              ("label_field", "message")])], lambda: events(600, 4), 0),
    ("counter_discard", [], [("log_to_metrics", BASE + [("discard_logs", "true"), ("label_field", "code")])],
     lambda: events(300, 5), 0),
    ("histogram_default_buckets", [], [("log_to_metrics", BASE + [("metric_mode", "histogram"), ("value_field", "duration"),
                                                                  ("label_field", "color")])], lambda: events(900, 6), 0),
    ("histogram_custom_buckets", [], [("log_to_metrics", BASE + [("metric_mode", "histogram"), ("value_field", "$duration"),
                                                                 ("bucket", "5"), ("bucket", "0.3"), ("bucket", "1"),
                                                                 ("metric_subsystem", "web"), ("metric_namespace", "shop"),
                                                                 ("add_label", "app $kubernetes['labels']['app']")])],
     lambda: events(900, 7), 0),
    ("histogram_regex", [], [("log_to_metrics", BASE + [("metric_mode", "histogram"), ("value_field", "duration"),
                                                        ("regex", "message ^ok")])], lambda: events(500, 8), 0),
    ("after_parser_and_grep",
     [dict(name="apache2", format="regex", regex=util.APACHE_RX, time_key="time", time_fmt="%d/%b/%Y:%H:%M:%S %z",
           types="code:integer size:integer")],
     [("parser", [("key_name", "log"), ("parser", "apache2"), ("reserve_data", "on")]),
      ("grep", [("exclude", "code ^5")]),
      ("log_to_metrics", BASE + [("label_field", "method"), ("label_field", "code")])],
     lambda: util.chunk_from_lines(util.apache_lines(1500, seed=11)), 2),
    ("discard_in_chain",
     [dict(name="apache2", format="regex", regex=util.APACHE_RX, time_key="time", time_fmt="%d/%b/%Y:%H:%M:%S %z",
           types="code:integer size:integer")],
     [("parser", [("key_name", "log"), ("parser", "apache2")]),
      ("log_to_metrics", BASE + [("metric_mode", "histogram"), ("value_field", "size"), ("label_field", "code"),
                                 ("bucket", "1000"), ("bucket", "10000"), ("discard_logs", "on")])],
     lambda: util.chunk_from_lines(util.apache_lines(1000, seed=12)), 1),
    ("counter_kubernetes_mode", [], [("log_to_metrics", BASE + [("kubernetes_mode", "on"), ("label_field", "color")])],
     lambda: events(400, 14, extra=lambda rng: [(b"kubernetes", {b"pod_name": rng.choice([b"web-1", b"web-2"]), b"namespace_name": b"prod",
                                                             b"container_name": rng.choice([b"app", b"sidecar"]), b"docker_id": b"abc",
                                                             b"pod_id": 17})] if rng.random() < 0.8 else []), 0),
    ("counter_float_labels", [], [("log_to_metrics", BASE + [("label_field", "ratio"), ("add_label", "c $color")])],
     lambda: events(700, 18, extra=float_label), 0),
    # cmetrics finds a metric by the hash of its label values run together: ("ab",""), ("","ab") and ("a","b") are ONE metric,
    # shown with the labels of the first record that had it (lib/cmetrics/src/cmt_map.c:208-224)
    ("counter_label_boundaries", [], [("log_to_metrics", BASE + [("label_field", "p"), ("label_field", "q")])],
     lambda: b"".join(util.event(1700000000 + i, 0, [(b"p", _mp(p)), (b"q", _mp(q))]) for i, (p, q) in enumerate(
         [(b"", b"ab"), (b"ab", b""), (b"a", b"b"), (b"x", b"y"), (b"xy", b""), (b"a", b"b"), (12, b"3"), (1, 23), (b"", b"")] * 7)), 0),
    ("gauge_label_boundaries", [], [("log_to_metrics", BASE + [("metric_mode", "gauge"), ("value_field", "v"), ("label_field", "p"), ("label_field", "q")])],
     lambda: b"".join(util.event(1700000000 + i, 0, [(b"p", _mp(p)), (b"q", _mp(q)), (b"v", _mp(i))]) for i, (p, q) in enumerate(
         [(b"", b"ab"), (b"ab", b""), (b"a", b"b"), (b"x", b"y"), (b"xy", b"")] * 5)), 0),
    # no label keys: cmetrics' static metric is there, at 0, before anything is counted
    ("counter_static_metric_at_zero", [], [("log_to_metrics", BASE + [("regex", "message ^nothing matches this$")])], lambda: events(50, 19), 0),
    ("gauge_static_metric_at_zero", [], [("log_to_metrics", BASE + [("metric_mode", "gauge"), ("value_field", "absent")])], lambda: events(50, 20), 0),
    # group markers and negative timestamps: skipped by the event decoder, but objects of the chunk for log_to_metrics
    # (msgpack_unpack_next, log_to_metrics.c:993) as long as no earlier filter rewrote the chunk
    ("counter_group_markers", [], [("log_to_metrics", BASE + [("label_field", "color")])], lambda: grouped_events(21), 0),
    ("counter_group_markers_after_notouch_grep", [], [("grep", [("Regex", "color .")]), ("log_to_metrics", BASE + [("label_field", "color")]),
                                                       ("modify", [("Add", "seen yes")])], lambda: grouped_events(22), 1),
    ("counter_group_markers_after_grep", [], [("grep", [("Exclude", "color blue")]), ("log_to_metrics", BASE + [("label_field", "color")])],
     lambda: grouped_events(23), 1),
    ("histogram_group_markers_discard", [], [("log_to_metrics", BASE + [("metric_mode", "histogram"), ("value_field", "duration"), ("label_field", "color"),
                                                                        ("discard_logs", "on")])], lambda: grouped_events(24), 0),
    ("counter_empty_namespace", [], [("log_to_metrics", BASE + [("metric_namespace", ""), ("label_field", "color")])], lambda: events(100, 25), 0),
    ("gauge_labels", [], [("log_to_metrics", BASE + [("metric_mode", "gauge"), ("value_field", "duration"), ("label_field", "color"),
                                                     ("add_label", "pod $kubernetes['pod_name']")])], lambda: events(900, 15), 0),
    ("gauge_regex_no_labels", [], [("log_to_metrics", BASE + [("metric_mode", "gauge"), ("value_field", "$code"), ("regex", "message ^ok"),
                                                              ("metric_subsystem", "last")])], lambda: events(700, 16), 0),
    ("gauge_after_parser",
     [dict(name="apache2", format="regex", regex=util.APACHE_RX, time_key="time", time_fmt="%d/%b/%Y:%H:%M:%S %z",
           types="code:integer size:integer")],
     [("parser", [("key_name", "log"), ("parser", "apache2")]),
      ("log_to_metrics", BASE + [("metric_mode", "gauge"), ("value_field", "size"), ("label_field", "method"), ("label_field", "code")])],
     lambda: util.chunk_from_lines(util.apache_lines(1200, seed=17)), 1),
    ("l2m_then_modify", [],
     [("log_to_metrics", BASE + [("label_field", "color")]), ("modify", [("add", "seen yes")])], lambda: events(200, 13), 0),
]
