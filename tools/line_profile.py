"""Source-line view of an ncu report: dynamic warp instructions and stall samples per CUDA source line.
    python tools/line_profile.py gpurun_out/prof.ncu-rep [top N]"""
import collections, csv, io, linecache, subprocess, sys
from pathlib import Path
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 80
ROOT = Path(__file__).resolve().parent.parent
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur, hd = None, None
inst, samp = collections.Counter(), collections.Counter()
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Line No":
        hd = r
        iI, iS = hd.index("Instructions Executed"), hd.index("# Samples")
        continue
    if hd is None or len(r) <= max(iI, iS) or r[2] == "":
        continue
    try:
        ln = int(r[0])
        inst[(cur, ln)] += int(r[iI] or 0)
        samp[(cur, ln)] += int(r[iS] or 0)
    except ValueError:
        continue
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print("total warp instructions", ti, "samples", ts)
for (fp, ln), v in inst.most_common(top):
    f = fp if Path(fp).exists() else str(ROOT / "raytracing_renderer_cuda_b200/csrc" / Path(fp).name)
    text = linecache.getline(f, ln).strip()
    print(f"{100 * v / ti:6.2f} {100 * samp[(fp, ln)] / ts:6.2f}  {Path(fp).name}:{ln}  {text[:120]}")
