#!/usr/bin/env python
"""Per-source-line stall samples of one kernel from an ncu report captured with
--import-source on (needs -lineinfo):  python profiles/srclines.py <rep> <kernel-regex> [top]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if r and r[0] == "Line No")
si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
cur_file = ""
data = []
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    if len(r) > si and r[0].isdigit():
        try:
            data.append((int(r[si]), int(r[ii]), cur_file, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1
print(f"total samples {tot}")
for n, ins, f, ln, src in sorted(data, reverse=True)[:top]:
    print(f"{n:7d} {100 * n / tot:5.1f}%  inst={ins:9d}  {f}:{ln}: {src}")
