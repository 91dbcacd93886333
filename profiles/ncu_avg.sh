#!/bin/bash
# usage: profiles/ncu_avg.sh <kernel-regex> <command...>   -> average gpu__time_duration per kernel name (us)
re=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$re --csv --log-file /tmp/ncu_avg.csv "$@" > /tmp/ncu_avg.out 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.DictReader([l for l in open('/tmp/ncu_avg.csv') if not l.startswith('==')]))
t = collections.defaultdict(list)
for r in rows:
    t[r['Kernel Name'].split('(')[0]].append(float(r['Metric Value'].replace(',', '')) / 1e3)
for k, v in t.items():
    v = v[len(v) // 3:]
    print('%9.1f us  x%d  %s' % (sum(v) / len(v), len(v), k))
PY
