#!/usr/bin/env python
"""Split a kernel's ncu source page (SASS view) into regions at BAR.SYNC and print samples / instructions / top stalls per region.
usage: ncu_regions.py rep.ncu-rep [kernel-index]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; ksel = sys.argv[2] if len(sys.argv) > 2 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = []; cur = None
for line in raw.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = {"name": line, "rows": []}; blocks.append(cur); continue
    if cur is not None: cur["rows"].append(line)
blocks = [b for b in blocks if b["rows"]]
if ksel.isdigit(): b = blocks[int(ksel)]
else: b = [x for x in blocks if ksel in x["name"]][0]
print(b["name"][:260])
rows = list(csv.reader(b["rows"]))
hdr = rows[0]; rows = rows[1:]
iS = hdr.index("Source"); iN = hdr.index("# Samples"); iE = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
regions = []; reg = {"n": 0, "samples": 0, "exec": 0, "stalls": collections.Counter(), "ops": collections.Counter(), "start": 0}
def num(x):
    try: return float(x)
    except: return 0.0
for k, r in enumerate(rows):
    src = r[iS].strip()
    op = src.split()[0] if src else ""
    if op.startswith("@"): op = src.split()[1]
    reg["n"] += 1; reg["samples"] += num(r[iN]); reg["exec"] += num(r[iE])
    for i, h in stall_cols: reg["stalls"][h] += num(r[i])
    key = op.split(".")[0]
    reg["ops"][key] += num(r[iE])
    if op.startswith("BAR") or op.startswith("EXIT"):
        regions.append(reg); reg = {"n": 0, "samples": 0, "exec": 0, "stalls": collections.Counter(), "ops": collections.Counter(), "start": k + 1}
regions.append(reg)
tot = sum(r["samples"] for r in regions) or 1
for j, r in enumerate(regions):
    if r["exec"] == 0 and r["samples"] == 0: continue
    top = ", ".join(f"{h[6:]}={v / max(r['samples'], 1):.2f}" for h, v in r["stalls"].most_common(5))
    ops = ", ".join(f"{o}:{int(v)}" for o, v in r["ops"].most_common(8))
    print(f"region {j:2d} sass[{r['start']:5d}+{r['n']:5d}] samples {100 * r['samples'] / tot:5.1f}%  warp-instr {int(r['exec']):9d}  | {top}\n      {ops}")
