#!/usr/bin/env python
"""Per-kernel opcode histogram of the built library (cuobjdump -sass / -res-usage; no GPU needed):
    python tools/sass_summary.py > profiles/sass_summary.txt
The first line carries the hash of the kernel sources (gym_lmaze_b200.build.source_hash); a CPU test fails when the
committed summary was made from other sources."""
import os, re, shutil, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gym_lmaze_b200 import build, _abi

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
CUFILT = shutil.which("cu++filt") or "/usr/local/cuda/bin/cu++filt"
OPS = ["UBLKCP.S.G", "UBLKCP.G.S", "SYNCS", "STG.E.EF.128", "STG.E.EF", "STG", "LDG", "LDS.128", "STS", "SHFL", "VOTE", "ATOM", "RED",
       "FADD", "FMUL", "FFMA2", "DADD", "DMUL", "DFMA", "HMMA", "UTCHMMA", "UTMALDG"]


def main():
    build.build()
    sass = subprocess.run([CUOBJDUMP, "-sass", _abi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    res = subprocess.run([CUOBJDUMP, "-res-usage", _abi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    regs = {}
    name = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
        if m and name:
            regs[name] = (int(m.group(1)), int(m.group(2)))
    funcs, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            funcs[name].append(line)
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print("csrc_hash %s   archs %s   kernels %d   (tools/sass_summary.py)" % (build.source_hash(), ",".join(archs), len(funcs)))
    names = sorted(funcs)
    dem = subprocess.run([CUFILT] + names, stdout=subprocess.PIPE, text=True).stdout.splitlines() if os.path.isfile(CUFILT) else names
    print("%-78s %5s %6s %6s " % ("kernel", "regs", "smem", "instr") + " ".join("%s" % o for o in OPS))
    for mangled, nice in sorted(zip(names, dem), key=lambda kv: kv[1]):
        body = funcs[mangled]
        text = "\n".join(body)
        counts = []
        for o in OPS:
            if o == "STG":
                counts.append(len(re.findall(r"\bSTG\.", text)))
            elif o == "STG.E.EF":
                counts.append(len(re.findall(r"\bSTG\.E\.EF\b(?!\.128)", text)) + len(re.findall(r"\bSTG\.E\.EF\.(?!128)", text)))
            else:
                counts.append(len(re.findall(r"\b" + re.escape(o), text)))
        nice = nice.replace("lmz::", "").replace("(int)", "").replace("void ", "").replace("(KParams)", "")
        r = regs.get(mangled, (0, 0))
        print("%-78s %5d %6d %6d " % (nice[:78], r[0], r[1], len(body)) + " ".join("%*d" % (len(o), c) for o, c in zip(OPS, counts)))


if __name__ == "__main__":
    main()
