#!/usr/bin/env python
"""Per-source-line summary of one `ncu --set full --import-source on` capture:
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tests/tools/ncu_lines.py [top_n]
Keeps the rows ncu aggregates per CUDA-C line (samples, instructions, lanes per instruction, the main stall reasons) and
drops the per-SASS rows, so that the result is a few KB instead of tens of MB."""
import csv, sys, os
top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rows, fname, hdr, kernel = [], "?", None, "?"
for r in csv.reader(sys.stdin):
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        kernel = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or r[0] == "" or len(r) < len(hdr):
        continue
    d = {}
    for i, h in enumerate(hdr):
        d.setdefault(h, r[i])           # "Source" appears twice: the first is the CUDA-C text
    def num(k):
        try: return float(d.get(k, "0").replace(",", ""))
        except ValueError: return 0.0
    rows.append((num("Warp Stall Sampling (All Samples)"), fname, r[0], num("Instructions Executed"), num("Thread Instructions Executed"),
                 num("stall_long_sb"), num("stall_no_inst"), num("stall_wait"), num("stall_branch_resolving"), num("stall_short_sb"),
                 num("stall_not_selected") + num("stall_selected"), num("L2 Theoretical Sectors Local"), r[1].strip()[:110]))
tot_s = sum(r[0] for r in rows) or 1.0
tot_i = sum(r[3] for r in rows) or 1.0
tot_t = sum(r[4] for r in rows)
print("kernel %s: %d source lines with code, %.0f samples, %.3g warp instructions, %.1f lanes per instruction" % (kernel, len(rows), tot_s, tot_i, tot_t / tot_i))
byfile = {}
for r in rows:
    a = byfile.setdefault(r[1], [0.0, 0.0]); a[0] += r[0]; a[1] += r[3]
for f, (s, i) in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print("  %-22s %5.1f %% of samples  %5.1f %% of instructions" % (f, 100 * s / tot_s, 100 * i / tot_i))
tot_l = sum(r[11] for r in rows) or 1.0
print("local-memory sectors to L2 (theoretical): %.3g; the lines that make them:" % tot_l)
for r in sorted(rows, key=lambda x: -x[11])[:25]:
    if r[11] <= 0: break
    print("  %-26s %5.1f %%  lanes %4.1f | %s" % ("%s:%s" % (r[1], r[2]), 100 * r[11] / tot_l, r[4] / r[3] if r[3] else 0, r[12]))
print("%-26s %6s %6s %8s %5s | %5s %5s %5s %5s %5s %5s | %s" % ("line", "smp%", "cum%", "inst%", "lanes", "lsb", "noin", "wait", "brch", "ssb", "sel", "source"))
cum = 0.0
for r in sorted(rows, key=lambda x: -x[0])[:top_n]:
    cum += r[0]
    s = r[0] or 1.0
    print("%-26s %6.2f %6.1f %8.2f %5.1f | %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f | %s" % (
        "%s:%s" % (r[1], r[2]), 100 * r[0] / tot_s, 100 * cum / tot_s, 100 * r[3] / tot_i, r[4] / r[3] if r[3] else 0,
        100 * r[5] / s, 100 * r[6] / s, 100 * r[7] / s, 100 * r[8] / s, 100 * r[9] / s, 100 * r[10] / s, r[12]))
