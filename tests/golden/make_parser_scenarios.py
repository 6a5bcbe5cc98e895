"""The reference's internal parser tests (tests/internal/parser_{regex,json,ltsv,logfmt}.c:
test_basic, test_time_key, test_time_keep, test_types, ...) as data: parser definition, input line,
the (key, value-text) pairs compare_msgpack() looks for, the timestamp the test expects -- read
from the files where they lie -- plus what the UNMODIFIED reference (oracle/_ref) returns for the
line.  Writes tests/golden/parser_scenarios.json; tests/test_parser_scenarios.py replays it.
Left out: test_decode_field_json (Decode_Field is not built).

usage: python tests/golden/make_parser_scenarios.py"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import util
from make_runtime_scenarios import CSTR, PARSER_ARGS, TRUE, FALSE, call_args, literals

INT = "/root/reference/tests/internal"
TYPE_NAMES = {"INT": "integer", "FLOAT": "float", "BOOL": "bool", "STRING": "string", "HEX": "hex"}


def functions(src):
    parts = re.split(r"\n(?:static )?void (test_\w+)\s*\(\s*(?:void)?\s*\)\s*\n\{", src)
    for i in range(1, len(parts), 2):
        yield parts[i], parts[i + 1].split("\n}\n")[0]


def main():
    out = []
    for fmt in ("regex", "json", "ltsv", "logfmt"):
        src = open(os.path.join(INT, "parser_%s.c" % fmt)).read()
        for name, body in functions(src):
            if "decode_field" in name:
                continue
            var = {m.group(1): literals(m.group(2)) for m in re.finditer(r"char\s*\*\s*(\w+)\s*=\s*((?:\s*" + CSTR + r")+)\s*;", body)}
            types = " ".join("%s:%s" % (k, TYPE_NAMES[t]) for k, t in zip(re.findall(r'strcpy\(types->key,\s*"(\w+)"\)', body),
                                                                        re.findall(r"types->type\s*=\s*FLB_PARSER_TYPE_(\w+)", body)))
            calls = call_args(re.sub(r"/\*.*?\*/", "", body, flags=re.S), "flb_parser_create")
            assert len(calls) == 1, name
            kw = {}
            for k, a in zip(PARSER_ARGS, calls[0]):
                if k in ("config", "types_len", "decoders") or a in ("NULL", "0"):
                    continue
                if k == "types":
                    kw[k] = types
                elif k == "time_system_timezone":
                    assert a in FALSE
                elif a in TRUE or a in FALSE:
                    kw[k] = a in TRUE
                elif a.startswith('"'):
                    kw[k] = literals(a)
                else:
                    kw[k] = var[a]
            pairs = re.search(r"expected_strs\[\]\s*=\s*\{(.*?)\}\s*;", body, re.S)
            strs = [literals(x) for x in re.findall(r'"(?:[^"\\]|\\.)*"', re.sub(r"/\*.*?\*/", "", pairs.group(1)))]
            tm = re.search(r"tv_sec == (\d+) && out_time\.tm\.tv_nsec == (\d+)", body)
            ref = util.Ref()
            r, data, t = ref.parser_do(ref.parser(**kw), var["input"].encode())
            out.append(dict(source="tests/internal/parser_%s.c" % fmt, test=name, parser=kw, input=var["input"],
                            pairs=[[strs[i], strs[i + 1]] for i in range(0, len(strs), 2)],
                            time=[int(tm.group(1)), int(tm.group(2))] if tm else None,
                            ret=r, out_hex=None if data is None else data.hex(), out_time=list(t)))
    json.dump(out, open(os.path.join(HERE, "parser_scenarios.json"), "w"), indent=0)
    print("wrote %d parser scenarios" % len(out))


if __name__ == "__main__":
    main()
