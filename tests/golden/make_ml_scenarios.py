"""The reference's internal multiline tests as data: the input / expected-output tables of tests/internal/multiline.c for the
rule-based parsers (java, ruby, python, go, the custom "elastic" and "endswith" parsers), read from the C file, plus what the
reference's filter_multiline (buffer off) makes of the inputs in one chunk.
    python tests/golden/make_ml_scenarios.py  ->  tests/golden/ml_scenarios.json"""
import base64
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util

SRC = open("/root/reference/tests/internal/multiline.c", encoding="utf-8").read()
ESC = {"n": "\n", "t": "\t", "r": "\r", "\\": "\\", '"': '"', "'": "'", "0": "\0"}


def c_strings(body):
    """the records of a `struct record_check x[] = { {"a" "b"}, {"c"} };` initialiser: adjacent literals concatenate"""
    recs, i, depth, cur = [], 0, 0, None
    while i < len(body):
        ch = body[i]
        if ch == "{":
            depth += 1
            if depth == 1:
                cur = ""
        elif ch == "}":
            if depth == 1 and cur is not None:
                recs.append(cur)
                cur = None
            depth -= 1
        elif ch == '"' and depth >= 1:
            i += 1
            out = []
            while body[i] != '"':
                if body[i] == "\\":
                    i += 1
                    out.append(ESC.get(body[i], body[i]))
                else:
                    out.append(body[i])
                i += 1
            cur = (cur or "") + "".join(out)
        elif body.startswith("/*", i):
            i = body.index("*/", i) + 1
        i += 1
    return recs


def table(name):
    m = re.search(r"struct record_check %s\[\] = \{" % name, SRC)
    start = m.end()
    depth, i = 1, start
    while depth:
        if SRC[i] == '"':
            i += 1
            while SRC[i] != '"':
                i += 2 if SRC[i] == "\\" else 1
        elif SRC[i] == "{":
            depth += 1
        elif SRC[i] == "}":
            depth -= 1
        i += 1
    return [s.encode() for s in c_strings(SRC[start:i - 1])]


def rules_of(func):
    """the flb_ml_rule_create(mlp, "from", "regex", "to", NULL) calls inside a test function"""
    a = SRC.index("static void %s()" % func)
    b = SRC.index("\n}\n", a)
    out = []
    for m in re.finditer(r'flb_ml_rule_create\(\s*\w+\s*,\s*"((?:[^"\\]|\\.)*)"\s*,\s*"((?:[^"\\]|\\.)*)"\s*,\s*"((?:[^"\\]|\\.)*)"', SRC[a:b]):
        out.append([re.sub(r"\\(.)", lambda k: ESC.get(k.group(1), "\\" + k.group(1)), g) for g in m.groups()])
    return out


CASES = [("java", "java", None), ("ruby", "ruby", None), ("python", "python", None), ("go", "go", None),
         ("elastic", "elastic", "test_parser_elastic"), ("endswith", "endswith", None)]

if __name__ == "__main__":
    B = lambda b: base64.b64encode(b).decode()
    out = []
    for name, tab, func in CASES:
        inp, exp = table(tab + "_input"), table(tab + "_output")
        ref = util.Ref()
        parser = name
        entry = {"name": name, "input": [B(x) for x in inp], "expected": [B(x) for x in exp]}
        if func:
            entry["rules"] = rules_of(func)
            ref.ml_parser("custom-" + name, rules=[tuple(r) for r in entry["rules"]])
            parser = "custom-" + name
        if name == "endswith":
            entry["type"], entry["match_string"], entry["negate"] = "endswith", "\\", True       # tests/internal/multiline.c test_endswith
            ref.ml_parser("custom-endswith", type="endswith", match_string="\\", negate=True)
            parser = "custom-endswith"
        entry["parser"] = parser
        props = [("multiline.parser", parser), ("multiline.key_content", "log"), ("buffer", "off")]
        rf = ref.filter("multiline", props)
        chunk = util.chunk_from_lines(inp)
        ret, res = ref.filter_cb(rf, chunk)
        entry["ret"], entry["out"] = ret, B(res)
        out.append(entry)
        vals = []
        for o, l in util.split_records(res):
            rec = res[o:o + l]
            k = rec.index(b"\xa3log") + 4
            hdr = rec[k]
            n, h = (hdr & 31, 1) if hdr < 0xc0 else (rec[k + 1], 2) if hdr == 0xd9 else (int.from_bytes(rec[k + 1:k + 3], "big"), 3)
            vals.append(rec[k + h:k + h + n])
        print(name, len(inp), "lines ->", len(vals), "messages; the test's table:", len(exp), "equal:", vals == exp)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ml_scenarios.json"), "w"), indent=0)
