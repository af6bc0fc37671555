#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals from an ncu report (source page, cuda view).
  python tools/prof_lines.py <report.ncu-rep> [file-substring] [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, acc = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1]; hdr = None; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and cur and want in cur and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            inst = int(d.get("Instructions Executed", "0") or 0); smp = int(d.get("# Samples", "0") or 0)
        except ValueError:
            continue
        if inst or smp:
            acc.append((int(r[0]), inst, smp, d.get("Avg. Threads Executed", ""), r[1].strip()[:110], cur.split("/")[-1]))
ti = sum(a[1] for a in acc) or 1; ts = sum(a[2] for a in acc) or 1
print(f"total inst {ti}  samples {ts}")
print("-- by instructions")
for a in sorted(acc, key=lambda a: -a[1])[:top]:
    print(f"{a[5]}:{a[0]:5d} inst {100*a[1]/ti:5.1f}%  smp {100*a[2]/ts:5.1f}%  thr {a[3]:>5s} | {a[4]}")
print("-- by samples")
for a in sorted(acc, key=lambda a: -a[2])[:top]:
    print(f"{a[5]}:{a[0]:5d} inst {100*a[1]/ti:5.1f}%  smp {100*a[2]/ts:5.1f}%  thr {a[3]:>5s} | {a[4]}")
