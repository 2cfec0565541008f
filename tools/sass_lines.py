"""SASS of an ncu report in address order, each instruction with its CUDA source line, share of the dynamic warp
instructions and of the stall samples:  python tools/sass_lines.py rep.ncu-rep > out.txt"""
import csv, io, subprocess, sys
from pathlib import Path
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur, line, hd = None, None, None
rows = {}
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = Path(r[1]).name
        continue
    if r[0] == "Line No":
        hd = r
        iA, iS, iI, iN, iT = 2, 3, hd.index("Instructions Executed"), hd.index("# Samples"), hd.index("Thread Instructions Executed")
        continue
    if hd is None or len(r) <= iI:
        continue
    if r[0] != "":
        line = r[0]
        continue
    if r[iA].startswith("0x") and r[iI].isdigit():
        rows[int(r[iA], 16)] = (cur, line, r[iS].strip(), int(r[iI]), int(r[iN] or 0), int(r[iT] or 0))
tot = sum(v[3] for v in rows.values()) or 1
ts = sum(v[4] for v in rows.values()) or 1
base = min(rows)
for a in sorted(rows):
    f, l, s, n, sm, t = rows[a]
    print(f"{a - base:6x} {100 * n / tot:6.3f} {100 * sm / ts:6.3f} {t / max(n, 1):4.0f} {f}:{l:>4s}  {s}")
