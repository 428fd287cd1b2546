#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md), from
`cuobjdump -sass libfhvae_b200.so`.  Usage: python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_scalablefhvae_b200 import _lib

OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "REDG", "MUFU",
       "ACQBULK", "PREEXIT", "LDG", "LDG.nc", "STG", "LDS", "STS", "BAR", "SHFL"]
out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
cur, counts = None, {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = demangle(m.group(1))
        counts[cur] = dict.fromkeys(OPS, 0)
        counts[cur]["instructions"] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        counts[cur]["instructions"] += 1
        op = m.group(1).split(".")[0]
        if op in counts[cur]:
            counts[cur][op] += 1
        if op == "LDG" and ".CONSTANT" in m.group(1):
            counts[cur]["LDG.nc"] += 1
print(f"# cuobjdump -sass {os.path.relpath(_lib.LIB_PATH)} (sm_100a): opcode counts per kernel; UTCHMMA = tcgen05.mma kind::f16,")
print("# LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk, SYNCS = mbarrier,")
print("# REDG = red.global (split-K partials), MUFU = ex2/rcp (gate non-linearities), ACQBULK = griddepcontrol.wait and")
print("# PREEXIT = griddepcontrol.launch_dependents (programmatic dependent launch); LDG.nc = ld.global.nc (must be 0 in a kernel")
print("# with ACQBULK: such loads are not ordered behind the wait, DESIGN.md 3.7)")
cols = ["instructions"] + OPS
print(f"{'kernel':58s} " + " ".join(f"{c:>8s}" for c in cols))
tot = dict.fromkeys(cols, 0)
for k in sorted(counts, key=lambda k: -counts[k]["UTCHMMA"] * 100000 - counts[k]["instructions"]):
    print(f"{k[:58]:58s} " + " ".join(f"{counts[k][c]:8d}" for c in cols))
    for c in cols:
        tot[c] += counts[k][c]
print(f"{'TOTAL':58s} " + " ".join(f"{tot[c]:8d}" for c in cols))
