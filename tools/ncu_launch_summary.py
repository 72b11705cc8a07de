"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the LAST
complete step (delimited by ce_fwd launches) and the share of each kernel.  usage: tools/ncu_launch_summary.py csv"""
import collections, csv, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
def us(x):
    v = float(x["Metric Value"].replace(",", "")); u = x["Metric Unit"]
    return v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
idx = [i for i, x in enumerate(rows) if "ce_fwd" in x["Kernel Name"]]
if len(idx) >= 3:
    step = rows[idx[-2]:idx[-1]]     # one full step, phase-shifted to start at a ce_fwd launch
else:
    step = rows
agg = collections.OrderedDict()
for x in step:
    n = x["Kernel Name"].split("(")[0][-60:]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us(x)
tot = sum(v[1] for v in agg.values())
print(f"one step: {len(step)} launches, {tot:.1f} us summed device time (cold-cache, serialised under ncu)")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  x{c:<3d} {t/c:8.1f} us/launch  {n}")
