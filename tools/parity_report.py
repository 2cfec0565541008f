"""gpurun_out/parity.jsonl (written by tests/conftest.record_parity during `pytest -m gpu`) -> markdown for profiles/:
    python tools/parity_report.py gpurun_out/parity.jsonl "title" > profiles/r02_parity.md"""
import collections, json, sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
print(f"# {sys.argv[2] if len(sys.argv) > 2 else 'Measured values behind the parity gates'}\n")
print("Every row is a value MEASURED on a B200 by a test of `tests/` (the assertion only bounds it); written by "
      "`tests/conftest.record_parity`, collected with `tools/parity_report.py`.\n")
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r["test"], []).append(r)
for test, rs in by.items():
    keys = [k for k in rs[0] if k != "test"]
    print(f"## {test}\n\n| " + " | ".join(keys) + " |\n|" + "---|" * len(keys))
    for r in rs:
        print("| " + " | ".join(f"{r.get(k):.4g}" if isinstance(r.get(k), float) else str(r.get(k)) for k in keys) + " |")
    print()
