"""JSON fixtures the reference's own tests hold, as lines for the JSON parser:
  tests/internal/data/pack/*.json  -- string documents (escapes, surrogate pairs) with the msgpack the
      reference's pack test expects beside them (*.mp): wrapped as {"v": <document>} so that they are maps,
      and the expected string must come out as the value; json_single_map_00{1,2}.json as they are;
  benchmarks/utf8_surrogate_bench_10k.ndjson -- 11 kinds of lines (valid / lone / boundary surrogates,
      NUL escapes, nested, arrays): the first two of each kind travel as golden vectors, all 10 000 are
      compared live against the reference when the tree is present (tests/test_json_scenarios.py).
Outputs are those of the UNMODIFIED reference's filter_parser (oracle/_ref).
usage: python tests/golden/make_json_scenarios.py"""
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import util

PACK = "/root/reference/tests/internal/data/pack"
NDJSON = "/root/reference/benchmarks/utf8_surrogate_bench_10k.ndjson"
KW = dict(name="js", format="json")
PROPS = [("Key_Name", "log"), ("Parser", "js")]


def reference(lines):
    ref = util.Ref()
    ref.parser(**KW)
    ref.filter("parser", PROPS)
    return ref.chain_do(util.chunk_from_lines(lines))


def main():
    out = []
    for path in sorted(glob.glob(os.path.join(PACK, "*.json"))):
        name = os.path.basename(path)[:-5]
        doc = open(path, "rb").read().strip()
        mp = path[:-5] + ".mp"
        if doc.startswith(b'"') and os.path.exists(mp):
            line = b'{"v": ' + doc + b"}"
            ret, res = reference([line])
            want = open(mp, "rb").read()
            assert want in res, name                          # the packed string the reference's pack test expects
            out.append(dict(name="pack/" + name, line_hex=line.hex(), ret=ret, out_hex=res.hex(), value_hex=want.hex()))
        elif name.startswith("json_single_map"):
            line = b" ".join(doc.split(b"\n"))
            ret, res = reference([line])
            out.append(dict(name="pack/" + name, line_hex=line.hex(), ret=ret, out_hex=res.hex(), value_hex=None))
    seen = {}
    for raw in open(NDJSON, "rb"):
        raw = raw.rstrip(b"\n")
        kind = json.loads(raw).get("test_case")
        if seen.get(kind, 0) < 2:
            seen[kind] = seen.get(kind, 0) + 1
            ret, res = reference([raw])
            out.append(dict(name="ndjson/%s/%d" % (kind, seen[kind]), line_hex=raw.hex(), ret=ret, out_hex=res.hex(), value_hex=None))
    json.dump(out, open(os.path.join(HERE, "json_scenarios.json"), "w"), indent=0)
    print("wrote %d JSON scenarios" % len(out))


if __name__ == "__main__":
    main()
