"""Per-SASS-instruction view of an `ncu --set full --import-source on` report (works without a GPU):
    python tools/sass_profile.py gpurun_out/prof.ncu-rep [--regions N] [--dump]
Prints dynamic warp-instruction totals by opcode, and — mapped through the CUDA source correlation — by source
file:line, so an instruction diet can be planned and checked (VERDICT r01 task 3: warp instructions per ray)."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iA, iS, iI, iT, iN = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ins = []
for r in rows[2:]:
    if len(r) <= iT or not r[iI].isdigit():
        continue
    ins.append((r[iS].strip(), int(r[iI]), int(r[iT]), int(r[iN] or 0)))
tot = sum(x[1] for x in ins)
tots = sum(x[3] for x in ins)
print("static instructions", len(ins), "dynamic warp instructions", tot, "avg threads", sum(x[2] for x in ins) / max(tot, 1))
by = collections.Counter()
bys = collections.Counter()
for s, n, t, sm in ins:
    op = s.split()[0]
    if op.startswith("@"):
        op = s.split()[1]
    op = op.split(".")[0].rstrip(";")
    by[op] += n
    bys[op] += sm
print("\nopcode  warp-instr%  stall-samples%")
for op, n in by.most_common(40):
    print(f"{op:10s} {100 * n / tot:6.2f} {100 * bys[op] / max(tots, 1):6.2f}")
if "--dump" in sys.argv:
    for k, (s, n, t, sm) in enumerate(ins):
        print(f"{k:5d} {100 * n / tot:6.3f} {100 * sm / max(tots, 1):6.3f} {t / max(n, 1):5.1f}  {s}")
