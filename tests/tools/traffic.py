"""profiles/traffic.json from ncu captures of the evaluation kernel (one launch over --lines events):
DRAM bytes (read + write) per event, which bench.py scales to the events of one step.
usage: python tests/tools/traffic.py gpurun_out/r01c_eval_json.ncu-rep json 1000000 [more rep wl lines ...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        tot += float(val[i]) * UNIT[units[i]]
    return tot


def main(argv):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        out = json.load(open(path))
    except Exception:
        out = {}
    for k in range(0, len(argv), 3):
        rep, wl, lines = argv[k], argv[k + 1], int(argv[k + 2])
        b = dram_bytes(rep)
        out["k_chain_eval_dram_bytes_per_event_" + wl] = b / lines
        out["source_" + wl] = "%s: %.1f MB DRAM read+write for one launch over %d events" % (os.path.basename(rep), b / 1e6, lines)
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1:])
