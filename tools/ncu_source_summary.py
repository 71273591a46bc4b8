#!/usr/bin/env python3
"""Source-page summary of an ncu --set full --import-source on capture: per kernel the opcode mix, the share of warp instructions and
stall samples by active-lane bucket, and the instructions that collect the most stall samples.
usage: python tools/ncu_source_summary.py gpurun_out/<tag>_trace.ncu-rep > profiles/<tag>_ncu_source.txt   (runs here, needs ncu only)"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:k_trace"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
print(f"# ncu source page of {rep}: SASS-level view of the traversal kernels (one launch each over the configs[1] streams)")
for k in kernels:
    if k["name"] in seen or not k["rows"]:
        continue
    seen.add(k["name"])
    h = k["hdr"]
    isrc, isamp, iex, iavg = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Avg. Threads Executed")
    tot = sum(int(r[isamp]) for r in k["rows"]) or 1
    totex = sum(int(r[iex]) for r in k["rows"]) or 1
    print(f"\n== {k['name']}\nSASS instructions {len(k['rows'])}, warp instructions executed {totex}, stall samples {tot}")
    b, bs = collections.Counter(), collections.Counter()
    op, ops = collections.Counter(), collections.Counter()
    for r in k["rows"]:
        a = float(r[iavg]) if r[iavg] not in ("", "-") else 0.0
        key = "<=4" if a <= 4 else "5-12" if a <= 12 else "13-19" if a <= 19 else "20-26" if a <= 26 else "27-32"
        b[key] += int(r[iex]); bs[key] += int(r[isamp])
        w = r[isrc].split()
        o = (w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "?")).split(".")[0]
        op[o] += int(r[iex]); ops[o] += int(r[isamp])
    print("active lanes per instruction : share of warp instructions / share of stall samples")
    for key in ("<=4", "5-12", "13-19", "20-26", "27-32"):
        print(f"  {key:6s} {100 * b[key] / totex:5.1f} % / {100 * bs[key] / tot:5.1f} %")
    print("opcode mix (share of warp instructions / of stall samples)")
    for o, c in op.most_common(16):
        print(f"  {o:8s} {100 * c / totex:5.1f} % / {100 * ops[o] / tot:5.1f} %")
    print("instructions with the most stall samples (share, avg active lanes, SASS)")
    for r in sorted(k["rows"], key=lambda r: -int(r[isamp]))[:14]:
        print(f"  {100 * int(r[isamp]) / tot:5.2f} %  {r[iavg]:>4s}  {r[isrc].strip()[:100]}")
