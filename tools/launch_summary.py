"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list (tools/profile_step.py) as a markdown table.
    python tools/launch_summary.py gpurun_out/launches_r02.csv [top_n] """
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row.get("Kernel Name")
    if not name:
        continue
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    unit = row.get("Metric Unit", "ns")
    v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"Total kernel time {tot / 1000:.2f} ms over {sum(a[0] for a in agg.values())} launches.\n")
print("| kernel | launches | total us | share | avg us |")
print("|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"| `{k[:80]}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.1f} |")
