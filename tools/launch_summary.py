"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into markdown.
    python tools/launch_summary.py gpurun_out/launches_r01.csv "command" > profiles/r01_launches.md"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [r for r in rows if "Kernel Name" in r][0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = [(r[ki], float(r[vi].replace(",", "")) / 1e3) for r in rows if len(r) == len(hdr) and r[0].isdigit()]  # us
tot = collections.defaultdict(lambda: [0, 0.0])
for k, v in d:
    name = k.split("(")[0].replace("void ", "").strip()
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v for _, v in d)
print(f"# launch list of `{sys.argv[2]}`\n")
print(f"`ncu --metrics gpu__time_duration.sum --clock-control none -c 900` (first {len(d)} launches; cold-cache, serialised: compare shares, "
      f"not absolutes).  Raw list: `profiles/{sys.argv[3] if len(sys.argv) > 3 else 'r01_launches_bench_c1.csv'}`.\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for name, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:70]}` | {n} | {v / 1e3:.3f} | {100 * v / total:.1f} % |")
step = [v for k, v in d if ("k_wf_step" in k or "k_wf_tail" in k)]
init = [i for i, (k, _) in enumerate(d) if "k_wf_init" in k]
if len(init) >= 2:
    frame = [v for k, v in d[init[0]:init[1]] if ("k_wf_step" in k or "k_wf_tail" in k)]
    print(f"\nOne frame = {len(frame)} `k_wf_step_*` / `k_wf_tail` launches, {sum(frame) / 1e3:.3f} ms under ncu.")
    print("Per-iteration durations of the first frame (us): " + str([round(v) for v in frame]))
