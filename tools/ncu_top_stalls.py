#!/usr/bin/env python
"""Top stall locations of an `ncu --page source --csv` dump (SASS view): address, samples, dominant stall, instruction.
    ncu -i X.ncu-rep --page source --csv > x.csv; python tools/ncu_top_stalls.py x.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = next(r for r in rows if len(r) > 20 and r[0] == "Address")
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for s in stalls:
        agg[s] = agg.get(s, 0) + int(r[ix[s]] or 0)
print("by reason:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(1, tot)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    n = int(r[ix["# Samples"]])
    why = max(stalls, key=lambda s: int(r[ix[s]] or 0))
    print("%5d  %5.1f%%  %-14s #%-5d x%-8s %s" % (n, 100.0 * n / tot, why[6:], i, r[ix["Instructions Executed"]], r[ix["Source"]].strip()))
