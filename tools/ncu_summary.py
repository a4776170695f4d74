#!/usr/bin/env python
"""Join ncu's raw page (`ncu -i X.ncu-rep --page raw --csv`) of tools/profile_final.py with its launch manifest:
writes the per-kernel summary table (markdown) and profiles/traffic.json (dram bytes per launch, keyed by the
hash of the kernel sources the capture ran on).

    python tools/ncu_summary.py RAW.csv MANIFEST.json OUT.md [TRAFFIC.json]"""
import csv, json, sys

raw, manifest, out_md = sys.argv[1], json.load(open(sys.argv[2])), sys.argv[3]
traffic_out = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in data if "lmz_" in r[ix["Kernel Name"]]]
L = manifest["launches"]
assert len(data) == len(L), "ncu saw %d lmz launches, the manifest lists %d" % (len(data), len(L))


def num(r, name):
    v = r[ix[name]].replace(",", "")
    u = units[ix[name]]
    x = float(v)
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "ns": 1e-6,
             "Tbyte/s": 1e12, "Gbyte/s": 1e9, "Mbyte/s": 1e6}.get(u, 1.0)
    return x * scale


stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
              and "not_issued" not in h]
lines = ["| # | launch | kernel | grid x block | regs | time (ms) | DRAM read + write (GB) | DRAM TB/s | algorithmic GB | traffic / algorithmic | issue active | top stalls (warps per issue) |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {"csrc_hash": manifest["csrc_hash"],
           "what": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from an `ncu --set full` capture at `envs` envs "
                   "(tools/profile_final.py); bench.py scales it linearly to its own batch size and withholds it when "
                   "csrc_hash differs from the hash of the sources it was built from"}
for i, (r, m) in enumerate(zip(data, L)):
    name = r[ix["Kernel Name"]]
    assert m["kernel"] in name, (i, m, name)
    short = name.replace("void ", "").replace("lmz::", "").replace("(int)", "").replace("(KParams)", "")
    t = num(r, "gpu__time_duration.sum")
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    bw = num(r, "dram__bytes.sum.per_second") / 1e12
    issue = float(r[ix["sm__inst_issued.avg.pct_of_peak_sustained_active"]])
    st = sorted(((float(r[ix[c]] or 0), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for c in stall_cols), reverse=True)[:2]
    alg = m["envs"] * m["bytes_per_env_step"] / 1e9 if m.get("bytes_per_env_step") else None
    grid = r[ix["Grid Size"]].strip("() ").split(",")[0]
    block = r[ix["Block Size"]].strip("() ").split(",")[0]
    lines.append("| %d | %s | `%s` | %s x %s | %s | %.3f | %.3f + %.3f | %.2f | %s | %s | %.1f %% | %s |" % (
        i, m["tag"], short, grid, block, r[ix["launch__registers_per_thread"]], t, rd / 1e9, wr / 1e9, bw,
        "%.3f" % alg if alg else "-", "%.3f" % ((rd + wr) / 1e9 / alg) if alg else "-", issue,
        ", ".join("%s %.1f" % (n, v) for v, n in st)))
    if m.get("traffic_key"):
        traffic[m["traffic_key"]] = {"bytes": rd + wr, "envs": m["envs"], "source": "%s, launch %d (%s)" % (raw, i, m["tag"]),
                                     "kernel": short, "time_ms": t}
open(out_md, "w").write("\n".join(lines) + "\n")
if traffic_out:
    json.dump(traffic, open(traffic_out, "w"), indent=1)
print("\n".join(lines))
