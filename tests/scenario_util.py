"""Helpers for the scenarios taken from the reference's runtime tests: in_lib style JSON events
(`[ts, {...}]`) to a chunk, and a chunk back to the compact JSON text out_lib's "format json" prints,
which is what those tests search with strstr()."""
import json
import re
import struct

import msgpack

import util


def mp_value(v):
    if v is True:
        return b"\xc3"
    if v is False:
        return b"\xc2"
    if v is None:
        return b"\xc0"
    if isinstance(v, int):
        return msgpack.packb(v)
    if isinstance(v, float):
        return b"\xcb" + struct.pack(">d", v)
    if isinstance(v, str):
        return util.mp_str(v.encode("utf-8"))
    if isinstance(v, Pairs):
        return util.mp_map_hdr(len(v)) + b"".join(util.mp_str(k.encode("utf-8")) + mp_value(x) for k, x in v)
    if isinstance(v, list):
        return _arr_hdr(len(v)) + b"".join(mp_value(x) for x in v)
    raise TypeError(v)


def _arr_hdr(n):
    return bytes([0x90 | n]) if n < 16 else b"\xdc" + struct.pack(">H", n)


class Pairs(list):
    """a JSON object as its (key, value) list: order and duplicate keys survive"""


def chunk_from_json_events(texts):
    out = []
    for t in texts:
        ts, body = json.loads(re.sub(r",\s*}", "}", t), object_pairs_hook=Pairs)     # jsmn tolerates a trailing comma
        sec = int(ts)
        nsec = int(round((ts - sec) * 1e9)) if isinstance(ts, float) else 0
        out.append(util.event(sec, nsec, [(k.encode("utf-8"), mp_value(v)) for k, v in body]))
    return b"".join(out)


def _json(v):
    if isinstance(v, Pairs):
        return "{" + ",".join(json.dumps(_s(k), ensure_ascii=False) + ":" + _json(x) for k, x in v) + "}"
    if isinstance(v, list):
        return "[" + ",".join(_json(x) for x in v) + "]"
    if isinstance(v, bytes):
        return json.dumps(_s(v), ensure_ascii=False)
    if isinstance(v, float):
        return repr(v)
    return json.dumps(v, ensure_ascii=False)


def _s(b):
    return b.decode("utf-8", "replace") if isinstance(b, bytes) else b


def records_as_json(chunk):
    """one compact JSON object text per record of a chunk (body only; duplicates and order kept)"""
    out = []
    if not chunk:
        return out
    u = msgpack.Unpacker(raw=True, strict_map_key=False, object_pairs_hook=Pairs)
    u.feed(chunk)
    for rec in u:
        out.append(_json(rec[1]))
    return out


def records_as_lib_lines(chunk):
    """`[<seconds with six decimals>,{...}]` per record: out_lib's "format json" line (flb_time_to_double, "%f")"""
    out = []
    if not chunk:
        return out
    u = msgpack.Unpacker(raw=True, strict_map_key=False, object_pairs_hook=Pairs)
    u.feed(chunk)
    for rec in u:
        ts = rec[0][0] if isinstance(rec[0], list) else rec[0]
        if isinstance(ts, msgpack.ExtType):
            sec, nsec = struct.unpack(">II", ts.data)
            ts = sec + nsec / 1e9
        out.append("[%f,%s]" % (float(ts), _json(rec[1])))
    return out


def scenario_events(sc):
    """the JSON events a scenario pushes: spelled out, or printed in a loop with (i, i * i)"""
    if sc.get("gen"):
        return [pat % (i, i * i) for i in range(sc["gen"]["n"]) for pat in sc["gen"]["patterns"]]
    return sc["inputs"] or []
