"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time share of ONE bench step
(the launches between the last two radius-graph count kernels).  usage: python profiles/parse_launches.py file.csv"""
import collections
import csv
import sys

with open(sys.argv[1]) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
names = [x["Kernel Name"] for x in rows]
idx = [i for i, n in enumerate(names) if "radius_open_kernel<0>" in n or ("radius_open_kernel" in n and "false" in n)]
step = rows[idx[-2]:idx[-1]]
tot, cnt = collections.OrderedDict(), collections.Counter()
for x in step:
    n = x["Kernel Name"].split("(")[0]
    tot[n] = tot.get(n, 0) + float(x["Metric Value"].replace(",", "")) / 1e3
    cnt[n] += 1
s = sum(tot.values())
print("one step: %.1f us of kernel time over %d launches (cold-cache, serialised: compare shares)" % (s, len(step)))
for n, t in sorted(tot.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%9.1f us %5.1f%% x%3d %s" % (t, 100 * t / s, cnt[n], n[:100]))
