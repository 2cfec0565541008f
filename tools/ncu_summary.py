"""Summarises an .ncu-rep (ncu --set full --import-source on) into markdown for profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/xyz.md
Reads the report with `ncu -i ... --page raw/source --csv` (works without a GPU)."""
import collections, csv, io, linecache, subprocess, sys
from pathlib import Path

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
ROOT = Path(__file__).resolve().parent.parent
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    return [r[hdr.index(name)] for r in data] if name in hdr else None


def f(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return float("nan")


print(f"# {title}\n")
print(f"Source: `{rep}` ({len(data)} launch(es) of `{col('Kernel Name')[0][:60]}`), captured with "
      "`ncu --set full --clock-control none --import-source on` (cold caches, serialised: use shares, not absolutes).\n")
keys = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
        ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy % (FP32-issue roofline)"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction (of 32)"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak")]
print("| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " | unit |\n|---|" + "---|" * (len(data) + 1))
for k, label in keys:
    c = col(k)
    if c:
        print(f"| {label} (`{k}`) | " + " | ".join(c) + f" | {units[hdr.index(k)]} |")
print("\n## Warp stall reasons (average warps stalled per issue-active cycle)\n")
st = []
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        st.append((h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), [f(r[i]) for r in data]))
st.sort(key=lambda x: -x[1][0])
print("| reason | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|" + "---|" * len(data))
for name, v in st:
    if max(v) >= 0.05:
        print(f"| {name} | " + " | ".join(f"{x:.2f}" for x in v) + " |")

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
byfile_i, byfile_s = collections.Counter(), collections.Counter()
for (fp, ln), v in inst.items():
    byfile_i[Path(fp).name] += v
    byfile_s[Path(fp).name] += samp[(fp, ln)]
print("\n## Where the instructions and the stall samples are (by source file)\n\n| file | instructions % | stall samples % |\n|---|---|---|")
for fn, v in byfile_i.most_common(8):
    print(f"| {fn} | {100 * v / ti:.1f} | {100 * byfile_s[fn] / ts:.1f} |")
print("\n## Top source lines by stall samples\n\n| line | stall % | instr % | source |\n|---|---|---|---|")
for (fp, ln), v in samp.most_common(24):
    text = linecache.getline(fp if Path(fp).exists() else str(ROOT / "raytracing_renderer_cuda_b200/csrc" / Path(fp).name), ln).strip().replace("|", "\\|")
    print(f"| {Path(fp).name}:{ln} | {100 * v / ts:.1f} | {100 * inst[(fp, ln)] / ti:.1f} | `{text[:110]}` |")
