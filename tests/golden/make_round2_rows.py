"""Golden vectors of the rows added late in round 2, produced from the UNMODIFIED reference (oracle/_ref):
  multiline_vectors.json -- filter_multiline (buffer off): parser definition, filter properties, chunks in, (ret, chunk) out per call
  tojson_vectors.json ---- flb_pack_msgpack_to_json_format(): arguments, chunk in, text out
  lines_vectors.json ----- in_tail's line loop (restated in the harness) + the reference's encoder: text in, chunk / consumed / lines out
    python tests/golden/make_round2_rows.py
The tests that read them (tests/test_golden.py) need neither /root/reference nor oracle/_ref."""
import base64
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
import test_zy_multiline as TM
import test_zz_tojson as TJ

B = lambda b: base64.b64encode(b).decode()


def multiline():
    rng = random.Random(31)
    out = []
    cases = [(name, rules, "regex", None, False, VOC) for (name, rules), VOC in zip(TM.RULESETS.items(), [TM.VOCAB[k] for k in TM.RULESETS])]
    cases += [("ends", [], "endswith", "\n", False, [b"one, ", b"two\n", b"", b"\n", b"x"]), ("eqn", [], "equal", "end", True, [b"end", b"a", b"", b"end "])]
    for name, rules, typ, ms, neg, voc in cases:
        ref = util.Ref()
        ref.ml_parser(name, rules=rules, type=typ, match_string=ms, negate=neg)
        props = [("multiline.parser", name), ("multiline.key_content", "log"), ("buffer", "off")]
        rf = ref.filter("multiline", props)
        calls = []
        for k in range(3):
            c = TM.make_chunk(rng, voc, 60, 1700000000 + 1000 * k, with_meta=0.3 * (k % 2))
            ret, res = ref.filter_cb(rf, c)
            calls.append({"in": B(c), "ret": ret, "out": B(res) if res is not None else None})
        out.append({"name": name, "type": typ, "match_string": ms, "negate": neg, "rules": rules, "props": props, "calls": calls})
    for name, text in (("java", TM.JAVA), ("go", TM.GO), ("python", TM.PY), ("ruby", TM.RUBY)):
        ref = util.Ref()
        props = [("multiline.parser", name), ("multiline.key_content", "log"), ("buffer", "off")]
        rf = ref.filter("multiline", props)
        calls = []
        for k, lines in enumerate((list(text), list(text)[3:] + list(text)[:3])):
            c = util.chunk_from_lines(lines, t0=1700000000 + 1000 * k)
            ret, res = ref.filter_cb(rf, c)
            calls.append({"in": B(c), "ret": ret, "out": B(res) if res is not None else None})
        out.append({"name": name, "builtin": True, "props": props, "calls": calls})
    return out


def tojson():
    rng = random.Random(32)
    ref = util.Ref()
    out = []
    for _ in range(40):
        c = TJ.chunk(rng, rng.choice([1, 4, 12]))
        jf, df = rng.randint(1, 3), rng.randint(0, 4)
        key = rng.choice(["date", "@timestamp", None])
        esc = rng.random() < 0.5
        text = ref.to_json(c, jf, df, key, esc)
        out.append({"in": B(c), "json_format": jf, "date_format": df, "date_key": key, "escape_unicode": esc,
                    "out": B(text) if text is not None else None})
    return out


def lines():
    import test_zz_lines as TL
    rng = random.Random(33)
    ref = util.Ref()
    out = []
    for r in range(30):
        t = TL.text(rng, rng.choice([0, 1, 3, 12, 40]))
        kw = dict(key=rng.choice(["log", "message"]), skip_empty_lines=bool(r % 2), sec=1700000000 + r, nsec=r * 1000)
        if r % 3 == 0:
            kw.update(path_key="file", path="/var/log/x-%d.log" % r)
        if r % 4 == 0:
            kw.update(offset_key="offset", stream_offset=2 ** 31 * (r % 8))
        chunk, used, n = ref.lines_to_events(t, **kw)
        out.append({"text": B(t), "kw": kw, "out": B(chunk) if chunk is not None else None, "consumed": used, "lines": n})
    return out


if __name__ == "__main__":
    g = os.path.join(ROOT, "tests", "golden")
    json.dump(multiline(), open(os.path.join(g, "multiline_vectors.json"), "w"), indent=0)
    json.dump(tojson(), open(os.path.join(g, "tojson_vectors.json"), "w"), indent=0)
    json.dump(lines(), open(os.path.join(g, "lines_vectors.json"), "w"), indent=0)
    print("written")
