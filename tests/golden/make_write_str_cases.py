"""The reference's own cases for flb_utils_write_str() as data: every `struct write_str_case cases[]` table of
tests/internal/utils.c (input, input length, expected output), with the escape_unicode flag of the loop that runs it
(write_str_test_cases: on, write_raw_str_test_cases: off).
    python tests/golden/make_write_str_cases.py  ->  tests/golden/write_str_cases.json"""
import base64
import json
import os
import re

SRC = open("/root/reference/tests/internal/utils.c", "rb").read()
SIMPLE = {ord("n"): 10, ord("t"): 9, ord("r"): 13, ord("\\"): 92, ord('"'): 34, ord("'"): 39, ord("0"): 0, ord("b"): 8, ord("f"): 12, ord("a"): 7, ord("v"): 11}


def literal(b, i):
    """the C string literal that starts at b[i] == '"': (bytes, index behind it)"""
    out = bytearray()
    i += 1
    while b[i] != 0x22:
        if b[i] == 0x5c:
            c = b[i + 1]
            if c == ord("x"):
                j = i + 2
                while chr(b[j]) in "0123456789abcdefABCDEF":
                    j += 1
                out.append(int(b[i + 2:j], 16) & 0xff)
                i = j
                continue
            if chr(c) in "01234567":
                j = i + 1
                while j < i + 4 and chr(b[j]) in "01234567":
                    j += 1
                out.append(int(b[i + 1:j], 8) & 0xff)
                i = j
                continue
            out.append(SIMPLE.get(c, c))
            i += 2
            continue
        out.append(b[i])
        i += 1
    return bytes(out), i + 1


def cases_of(body):
    """[(input, length, output)] of one `cases[] = { {...}, ... }` initialiser"""
    out, i, depth = [], 0, 0
    fields, cur_lit, cur_txt = [], None, b""
    while i < len(body):
        c = body[i]
        if body.startswith(b"/*", i):
            i = body.index(b"*/", i) + 2
            continue
        if c == 0x22 and depth == 1:
            s, i = literal(body, i)
            cur_lit = (cur_lit or b"") + s
            continue
        if c == ord("{"):
            depth += 1
            if depth == 1:
                fields, cur_lit, cur_txt = [], None, b""
        elif c == ord("}"):
            if depth == 1:
                fields.append(cur_lit if cur_lit is not None else cur_txt.strip())
                if len(fields) >= 3 and isinstance(fields[0], bytes) and fields[1].strip().isdigit():
                    out.append((fields[0], int(fields[1]), fields[2]))
            depth -= 1
        elif c == ord(",") and depth == 1:
            fields.append(cur_lit if cur_lit is not None else cur_txt.strip())
            cur_lit, cur_txt = None, b""
        elif depth == 1:
            cur_txt += bytes([c])
        i += 1
    return out


if __name__ == "__main__":
    out = []
    for m in re.finditer(rb"void (test_write_[a-z_0-9]+)\(\)\s*\{", SRC):
        name = m.group(1).decode()
        if name == "test_write_str_buffer_overrun":        # what fits a 100-byte output buffer: no such limit on this side of the call
            continue
        end = SRC.index(b"\n}\n", m.end())
        fn = SRC[m.end():end]
        k = fn.find(b"cases[] = {")
        if k < 0:
            continue
        esc = b"write_raw_str_test_cases" not in fn
        a = k + len(b"cases[] = ")
        depth, j = 0, a
        while True:
            if fn[j] == 0x22:
                _, j = literal(fn, j)
                continue
            if fn.startswith(b"/*", j):
                j = fn.index(b"*/", j) + 2
                continue
            depth += fn[j] == ord("{")
            depth -= fn[j] == ord("}")
            j += 1
            if depth == 0:
                break
        for inp, n, exp in cases_of(fn[a + 1:j - 1]):
            out.append({"test": name, "escape_unicode": esc, "input": base64.b64encode(inp[:n]).decode(), "expected": base64.b64encode(exp).decode()})
    ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "write_str_cases.json"), "w"), indent=0)
    print(len(out), "cases from", sorted({c["test"] for c in out}))

    # tests/internal/pack.c test_utf8_to_json: data/pack/*.mp (a msgpack string each) -> the JSON text of the .json beside it
    import glob
    pk = []
    for mp in sorted(glob.glob("/root/reference/tests/internal/data/pack/*.mp")):
        js = open(mp[:-3] + ".json", "rb").read()
        pk.append({"name": os.path.basename(mp), "msgpack": base64.b64encode(open(mp, "rb").read()).decode(), "json": base64.b64encode(js.rstrip(b"\n")).decode()})
    json.dump(pk, open(os.path.join(ROOT, "tests", "golden", "pack_to_json_cases.json"), "w"), indent=0)
    print(len(pk), "pack fixtures")
