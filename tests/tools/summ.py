"""one-line summary of a bench.py JSON line (helper for the GPU experiment scripts)"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ns = d.get("north_star") or {}
print("value %.1f e2e %.1f ms %.1f kms %s frac %.4f traffic %.1f GB cpu %.1f | north-star %.1f / %.1f | launches %d clocks %s" % (
    d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], {k: round(v, 1) for k, v in d["kernel_ms_per_step"].items()},
    d["roofline"]["frac"], (d["roofline"]["traffic"] or 0) / 1e9, (d.get("cpu_baseline") or {}).get("value", 0) / 1e6,
    ns.get("value", 0) / 1e6, ns.get("e2e", 0) / 1e6, d["gpu_launches"], d["clocks"]))
