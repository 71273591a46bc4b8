#!/usr/bin/env python3
"""CUDA-source-line view of an ncu --set full --import-source on capture: stall samples and warp instructions per source line.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N] [launch index]"""
import collections, csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
which = int(sys.argv[4]) if len(sys.argv) > 4 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# the export repeats a header line per (kernel launch, source file); lines with a numeric first column and a '-' address are source-line aggregates
launches, cur = [], None
for r in rows:
    if r and r[0] == "Function Name":
        cur = {"name": r[1], "lines": []}; launches.append(cur)
    elif r and r[0] == "Line No":
        hdr = r
    elif cur is not None and len(r) > 10 and r[0].isdigit() and r[2] == "-":
        cur["lines"].append(r)
iS, iE, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
iL = hdr.index("stall_long_sb")
for n, L in enumerate(launches):
    if which is not None and n != which:
        continue
    tot = sum(int(r[iS]) for r in L["lines"]) or 1
    totE = sum(int(r[iE]) for r in L["lines"]) or 1
    if tot < 50:
        continue
    print(f"\n== launch {n}: {L['name'][:90]}\nstall samples {tot}, warp instructions {totE}")
    print("  samples%  inst%  lanes  long_sb%  line  source")
    for r in sorted(L["lines"], key=lambda r: -int(r[iS]))[:top]:
        e = int(r[iE]); lanes = int(r[iT]) / e if e else 0
        print(f"  {100*int(r[iS])/tot:6.2f}  {100*e/totE:6.2f}  {lanes:5.1f}  {100*int(r[iL])/max(int(r[iS]),1):6.1f}  {r[0]:>5s}  {r[1].strip()[:110]}")
