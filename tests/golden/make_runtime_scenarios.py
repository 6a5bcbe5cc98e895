"""The reference's own runtime tests of the filters on the path, as data.

tests/runtime/filter_modify.c and filter_record_modifier.c configure one filter, push JSON events
through in_lib and look for substrings in out_lib's JSON; tests/runtime/filter_grep.c counts the
records that come out, or expects flb_start() to fail.  This script reads those files where they lie
(/root/reference, this container only), extracts per test: the filter properties, the pushed events,
and what the test asserts; runs the UNMODIFIED reference filter (oracle/_ref) on the resulting chunk;
and writes tests/golden/runtime_scenarios.json -- which travels, the reference does not.
tests/test_runtime_scenarios.py then asserts, on the CPU emulation and on the GPU, the reference
test's own expectation AND byte equality with the reference's output.

usage: python tests/golden/make_runtime_scenarios.py"""
import hashlib
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import util
from scenario_util import chunk_from_json_events, scenario_events

RT = "/root/reference/tests/runtime"
CSTR = r'"((?:[^"\\]|\\.)*)"'
REFUSED = re.compile(r"flb_start\([^)]*\);(?:\s|/\*.*?\*/)*TEST_CHECK\(ret != 0\)", re.S)


def unesc(s):
    return bytes(s, "utf-8").decode("unicode_escape").encode("latin-1").decode("utf-8")


def literals(text):
    """adjacent C string literals in `text`, concatenated"""
    return "".join(unesc(x) for x in re.findall(CSTR, text))


def call_args(body, fn):
    """argument lists (split at top-level commas) of every call of `fn` in `body`"""
    out = []
    for m in re.finditer(re.escape(fn) + r"\s*\(", body):
        i, depth, start, args, in_str = m.end(), 1, m.end(), [], False
        while depth:
            c = body[i]
            if in_str:
                if c == "\\":
                    i += 1
                elif c == '"':
                    in_str = False
            elif c == '"':
                in_str = True
            elif c == "(":
                depth += 1
            elif c == ")":
                depth -= 1
                if depth == 0:
                    args.append(body[start:i])
            elif c == "," and depth == 1:
                args.append(body[start:i])
                start = i + 1
            i += 1
        out.append([a.strip() for a in args])
    return out


def functions(src):
    parts = re.split(r"\n(?:static )?void (flb_\w+)\s*\(\s*(?:void)?\s*\)\s*\n\{", src)
    for i in range(1, len(parts), 2):
        yield parts[i], parts[i + 1].split("\n}\n")[0]


def filter_props(body):
    props = []
    for args in call_args(body, "flb_filter_set"):
        rest = [a for a in args[2:] if a != "NULL"]
        for k, v in zip(rest[0::2], rest[1::2]):
            k, v = literals(k), literals(v)
            if k.lower() != "match":
                props.append([k, v])
    return props


def pushed(body):
    return [literals(m.group(1)) for m in re.finditer(r"\bp\s*=\s*((?:\s*" + CSTR + r")+)\s*;", body)]


def extract_modify():
    src = open(os.path.join(RT, "filter_modify.c")).read()
    for name, body in functions(src):
        sc = dict(source="tests/runtime/filter_modify.c", test=name, filter="modify", props=filter_props(body), inputs=pushed(body),
                  gen=None, present=[], absent=[], count=None, init_error=False)
        m = re.search(r"cb_data\.data\s*=\s*((?:\s*" + CSTR + r")+)\s*;", body)
        if m:
            sc["present"].append(literals(m.group(1)))
        if REFUSED.search(body):                              # the configuration is refused
            sc.update(init_error=True, present=[])
        if name == "flb_test_not_drop_multi_event":          # callback_count: every event comes out, one is changed
            sc["count"] = len(sc["inputs"])
        if not sc["inputs"]:                                  # the event is a local array, not `p = "..."`
            m = re.search(r"char\s*\*?\s*\w+\s*(?:\[\])?\s*=\s*((?:\s*" + CSTR + r")+)\s*;", body)
            sc["inputs"] = [literals(m.group(1))] if m else []
        yield sc


def extract_record_modifier():
    src = open(os.path.join(RT, "filter_record_modifier.c")).read()
    for name, body in functions(src):
        props = filter_props(body)
        if not props or name in ("flb_uuid_key", "flb_test_json_long"):     # uuid_key: random by definition; json_long: checks in_lib
            continue
        sc = dict(source="tests/runtime/filter_record_modifier.c", test=name, filter="record_modifier", props=props, inputs=pushed(body),
                  gen=None, present=[], absent=[], count=None, init_error=False)
        if REFUSED.search(body):
            sc.update(init_error=True)
        for m in re.finditer(r"\{\s*(" + CSTR + r")\s*,\s*(FLB_TRUE|FLB_FALSE)\s*\}", body):
            (sc["present"] if m.group(3) == "FLB_TRUE" else sc["absent"]).append(literals(m.group(1)))
        yield sc


PARSER_ARGS = ["name", "format", "regex", "skip_empty", "time_fmt", "time_key", "time_offset", "time_keep", "time_strict",
               "time_system_timezone", "logfmt_no_bare_keys", "types", "types_len", "decoders", "config"]
TRUE, FALSE = ("FLB_TRUE", "MK_TRUE"), ("FLB_FALSE", "MK_FALSE")


def parser_defs(body):
    """flb_parser_create(...) calls of a test as keyword dicts; None when an argument is not a literal"""
    out = []
    for args in call_args(re.sub(r"//[^\n]*", "", body), "flb_parser_create"):
        kw = {}
        for k, a in zip(PARSER_ARGS, args):
            if k in ("config", "types_len"):
                continue
            if a == "NULL" or a == "0":
                continue
            if k == "time_system_timezone":
                if a not in FALSE:
                    return None                                 # depends on the machine's time zone
            elif a in TRUE or a in FALSE:
                kw[k] = a in TRUE
            elif a.startswith('"'):
                kw[k] = literals(a)
            else:
                return None
        out.append(kw)
    return out


def expectations(body):
    """strstr(output, X) != NULL / == NULL checks, X a literal or the last `expected = "..."`"""
    present, absent, last = [], [], None
    for m in re.finditer(r'expected\s*=\s*((?:\s*' + CSTR + r')+)\s*;|strstr\(output,\s*(expected|(?:' + CSTR + r'\s*)+)\)\s*(!=|==)\s*NULL', body):
        if m.group(1) is not None:
            last = literals(m.group(1))
            continue
        what = last if m.group(3) == "expected" else literals(m.group(3))
        (present if m.group(5) == "!=" else absent).append(what)
    if last is not None and re.search(r"strcmp\(output,\s*expected\)\s*==\s*0", body):
        present.append(last)                                    # the whole line; out_lib prints one record per line
    return present, absent


def extract_parser():
    """tests/runtime/filter_parser.c.  Left out: the tests that read the machine's time zone (use_system_timezone*),
    and one expectation that pins how out_lib prints an ext timestamp it cannot decode (fractional_timestamp)."""
    src = open(os.path.join(RT, "filter_parser.c")).read()
    helper = dict(functions(src.replace("static struct test_ctx *test_ctx_create(char *reserve_data, char *preserve_key)",
                                        "void flb_helper_test_ctx_create()")))
    for name, body in functions(src):
        if not name.startswith("flb_test_") or "system_timezone" in name:
            continue
        m = re.search(r'test_ctx_create\("(\w+)",\s*"(\w+)"\)', body)
        if m:                                                   # the four Reserve_Data x Preserve_Key tests share a helper
            hb = helper["flb_helper_test_ctx_create"]
            parsers = parser_defs(hb)
            props = [[k, {"reserve_data": m.group(1), "preserve_key": m.group(2)}.get(v, v)] for k, v in
                     [[k, v] for k, v in filter_props(hb.replace("reserve_data,", '"reserve_data",').replace("preserve_key,", '"preserve_key",'))]]
            inputs = [literals(a[2]) for a in call_args(body, "flb_lib_push")]
        else:
            parsers, props, inputs = parser_defs(body), filter_props(body), pushed(body)
        if not parsers or not props or not inputs:
            continue
        present, absent = expectations(body)
        present = [x for x in present if not x.startswith('["\\x')]
        yield dict(source="tests/runtime/filter_parser.c", test=name, filter="parser", props=props, parsers=parsers, inputs=inputs,
                   gen=None, present=present, absent=absent, count=None, init_error=False)


def l2m_table():
    """tests/runtime/filter_log_to_metrics.c.  The properties and the three JSON_MSG events are read from the file;
    which event is pushed how often is spelled out here and the assertion -- there a substring of the cmetrics JSON
    out_lib prints, e.g. `"value":5.0,"labels":["red","right"]` -- is restated on the text form of the same table."""
    src = open(os.path.join(RT, "filter_log_to_metrics.c")).read()
    msg = {}
    for m in re.finditer(r"#define (JSON_MSG\d)((?:[^\n]*\\\n)*[^\n]*\n)", src):
        msg[m.group(1)] = literals(m.group(2))
    k8s = 'namespace_name="k8s-dummy",pod_name="testpod",container_name="mycontainer",docker_id="abc123",pod_id="def456",'
    rows = [
        ("flb_test_log_to_metrics_counter_k8s", [("JSON_MSG1", 5)],
         ['log_metric_counter_test{' + k8s + 'color="red",direction="right"} = 5\n']),
        ("flb_test_log_to_metrics_counter", [("JSON_MSG1", 5)], ['myns_subsystem_test{color="red",direction="right"} = 5\n']),
        ("flb_test_log_to_metrics_counter_k8s_two_tuples", [("JSON_MSG1", 5), ("JSON_MSG2", 3)],
         ['{' + k8s + 'color="red",direction="right"} = 5\n', '{' + k8s + 'color="red",direction="left"} = 3\n']),
        ("flb_test_log_to_metrics_gauge", [("JSON_MSG1", 1)], ['log_metric_gauge_test{color="red",direction="right"} = 20\n']),
        ("flb_test_log_to_metrics_histogram", [("JSON_MSG1", 5)],
         ['log_metric_histogram_test{color="red",direction="right"} = { buckets = { 0.005=0, 0.01=0, 0.025=0, 0.05=0, 0.1=0, 0.25=0, '
          '0.5=0, 1=0, 2.5=0, 5=0, 10=0, +Inf=5 }, sum=100, count=5 }\n']),
        ("flb_test_log_to_metrics_reg", [("JSON_MSG1", 1), ("JSON_MSG3", 1)] * 3, ['{color="red",direction="left"} = 3\n']),
        ("flb_test_log_to_metrics_empty_label_keys_regex", [("JSON_MSG3", 3)], ['log_metric_counter_test = 3\n']),
        ("flb_test_log_to_metrics_label", [("JSON_MSG1", 2)], ['log_metric_counter_test{pod_name="testpod"} = 2\n']),
    ]
    bodies = dict(functions(src))
    for name, pushes, present in rows:
        props = [p for p in filter_props(bodies[name]) if p[0].lower() != "match"]
        yield dict(source="tests/runtime/filter_log_to_metrics.c", test=name, filter="log_to_metrics", props=props,
                   inputs=[msg[m] for m, n in pushes for _ in range(n)], gen=None, present=present, absent=[], count=None, init_error=False)


def grep_table():
    """tests/runtime/filter_grep.c prints its events in loops: `gen` = the format strings of one iteration
    (arguments i, i * i) and the iteration count, checked against the file below"""
    n = 256
    end = '[%d, {"val": "%d","END_KEY": "JSON_END"}]'
    dep = '[%d, {"val": "%d","log": "Using deprecated option"}]'
    opt = '[%d, {"val": "%d","log": "Using option"}]'
    rows = [
        ("flb_test_filter_grep_regex", [["Regex", "val 1"]], [end], None, False),
        ("flb_test_filter_grep_exclude", [["Exclude", "val 1"]], [end], None, False),
        ("flb_test_filter_grep_invalid", [["Regex", "val"], ["Exclude", "val"]], [], None, True),
        ("flb_test_filter_grep_multi_exclude", [["Exclude", "log deprecated"], ["Exclude", "log hoge"]], [dep, opt], n, False),
        ("flb_test_filter_grep_unknown_property", [["UNKNOWN_PROPERTY", "aaaaaa"]], [], None, True),
        ("flb_test_issue_5209", [["Exclude", "log /Using deprecated option/"]], [end, dep], n, False),
        ("flb_test_filter_grep_multi_regex", [["Regex", "log deprecated"], ["Regex", "log option"]], [dep, opt], n, False),
        ("flb_test_error_AND_regex_exclude", [["Regex", "val 1"], ["Exclude", "val2 3"], ["Logical_Op", "AND"]], [], None, True),
        ("flb_test_error_OR_regex_exclude", [["Regex", "val 1"], ["Exclude", "val2 3"], ["Logical_Op", "OR"]], [], None, True),
        ("flb_test_AND_regex", [["Regex", "log deprecated"], ["Regex", "log option"], ["Logical_Op", "AND"]], [dep, opt], n, False),
        ("flb_test_OR_regex", [["Regex", "log deprecated"], ["Regex", "log option"], ["Logical_Op", "OR"]], [dep, opt], 2 * n, False),
        ("flb_test_AND_exclude", [["Exclude", "log deprecated"], ["Exclude", "log option"], ["Logical_Op", "AND"]], [dep, opt], n, False),
        ("flb_test_OR_exclude", [["Exclude", "log deprecated"], ["Exclude", "log other"], ["Logical_Op", "OR"]], [dep, opt], n, False),
    ]
    src = open(os.path.join(RT, "filter_grep.c")).read()
    for name, props, pats, count, init_error in rows:
        body = dict(functions(src))[name]
        assert filter_props(body) == props, (name, filter_props(body))         # the table above is what the file configures
        printed = [literals(a[2]) for a in call_args(body, "snprintf")]
        assert init_error or printed == pats, (name, printed)
        yield dict(source="tests/runtime/filter_grep.c", test=name, filter="grep", props=props, inputs=None,
                   gen=dict(patterns=pats, n=n) if pats else None, present=[], absent=[], count=count, init_error=init_error)


def l2m_text(ref, f):
    import ctypes as C
    ref.L.flbref_l2m_cmt_text.restype = C.c_void_p
    ref.L.flbref_l2m_cmt_text.argtypes = [C.c_void_p]
    p = ref.L.flbref_l2m_cmt_text(f)
    t = C.string_at(p).decode(errors="replace")
    ref.L.flbref_cfree(C.c_void_p(p))
    return re.sub(r"^\S+Z ", "", t, flags=re.M)


def main():
    out = []
    for sc in list(extract_modify()) + list(extract_record_modifier()) + list(grep_table()) + list(extract_parser()) + list(l2m_table()):
        ref = util.Ref()
        if sc["init_error"]:
            try:
                ref.filter(sc["filter"], [tuple(p) for p in sc["props"]])
            except RuntimeError:
                out.append(sc)
                continue
            raise SystemExit("the reference accepted %s" % sc["test"])
        texts = scenario_events(sc)
        assert texts, sc["test"]
        chunk = chunk_from_json_events(texts)
        for kw in sc.get("parsers") or []:
            ref.parser(**kw)
        rf = ref.filter(sc["filter"], [tuple(p) for p in sc["props"]])
        ret, res = ref.chain_do(chunk)
        if sc["filter"] == "log_to_metrics":
            sc["text"] = l2m_text(ref, rf)
        sc.update(ret=ret, out_len=None if res is None else len(res))
        if res is not None and len(res) > 2048:              # long outputs travel as a digest
            sc.update(out_hex=None, out_sha256=hashlib.sha256(res).hexdigest())
        else:
            sc.update(out_hex=None if res is None else res.hex(), out_sha256=None)
        out.append(sc)
    json.dump(out, open(os.path.join(HERE, "runtime_scenarios.json"), "w"), indent=0)
    print("wrote %d scenarios: %s" % (len(out), {k: sum(s["filter"] == k for s in out) for k in sorted({s["filter"] for s in out})}))


if __name__ == "__main__":
    main()
