"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

usage: python profiles/summarize_launches.py gpurun_out/launches.csv [skip_first_n_launches] > profiles/<name>.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import collections
import csv
import re
import sys


def main(path, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], val * scale))
    rows = rows[skip:]
    agg = collections.OrderedDict()
    for name, ms in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    print(f"launches: {len(rows)}  total device time under ncu: {total:.1f} ms\n")
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:110]}` | {n} | {ms:.2f} | {100 * ms / total:.2f} % |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
