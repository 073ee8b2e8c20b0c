"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one steady decode step (the window between
the last two accept-walk kernels) as a per-kernel table (markdown) + the raw rows of that window (csv)."""
import collections
import csv
import re
import sys

src, out_md, out_csv = sys.argv[1:4]
lines = [l for l in open(src) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
names = [re.sub(r"\(.*", "", re.sub(r"<.*", "", r["Kernel Name"])).replace("void ", "")[:60] for r in rows]
vals = [float(r["Metric Value"].replace(",", "")) * scale[r["Metric Unit"]] for r in rows]
acc = [i for i, n in enumerate(names) if "accept_" in n]
a0, a1 = acc[-2], acc[-1]
tot = collections.defaultdict(lambda: [0, 0.0])
for i in range(a0, a1):
    tot[names[i]][0] += 1
    tot[names[i]][1] += vals[i]
T = sum(v[1] for v in tot.values())
with open(out_md, "w") as f:
    f.write(f"One steady decode step (config c2, 68m->7B, tree 128): {a1 - a0} kernel launches, "
            f"sum of ncu per-launch durations {T:.0f} us (cold-cache, serialised: compare SHARES, not absolutes).\n\n")
    f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / T:.1f}% | {v[1] / v[0]:.2f} |\n")
with open(out_csv, "w") as f:
    f.write("index,kernel,duration_us\n")
    for i in range(a0, a1):
        f.write(f"{i},{names[i]},{vals[i]:.3f}\n")
print(open(out_md).read())
