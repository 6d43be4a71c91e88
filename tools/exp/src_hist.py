"""Summarise the ncu source page of a report: executed warp instructions by SASS region.  usage: src_hist.py rep [chunk] [lo hi]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
IE, TE, SM = ix["Instructions Executed"], ix["Thread Instructions Executed"], ix["# Samples"]
tot = sum(int(r[IE]) for r in data); smp = sum(int(r[SM]) for r in data)
print("total warp instr", tot, "samples", smp)
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    for i in range(lo, min(hi, len(data))):
        r = data[i]; ie = int(r[IE]); te = int(r[TE])
        print(i, r[ix["Source"]][:64].ljust(64), f"{ie/1e6:8.1f}M thr {te/max(ie,1):5.1f} smp {r[SM]}")
else:
    for i in range(0, len(data), chunk):
        ch = data[i:i + chunk]
        ie = sum(int(r[IE]) for r in ch); te = sum(int(r[TE]) for r in ch); s = sum(int(r[SM]) for r in ch)
        if ie: print(f"{i:5d} {ie/tot*100:5.1f}% instr  {s/smp*100:5.1f}% samples  avg thr {te/max(ie,1):5.1f}")
