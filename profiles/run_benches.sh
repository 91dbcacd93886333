#!/bin/bash
# usage: run_benches.sh <tag> <workload> [workload...]   -- bench.py on one GPU, JSON lines under gpurun_out/
mkdir -p gpurun_out
tag=$1; shift
for wl in "$@"; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${tag}_bench_${wl}.json 2> gpurun_out/${tag}_bench_${wl}.err
  echo "$wl rc=$?"; grep -v Warning gpurun_out/${tag}_bench_${wl}.err | tail -c 1500
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${wl}.json").read().strip().splitlines()[-1])
    print("$wl", "ms/step", round(d["ms_per_step"],3), "atoms/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "step_frac", round(d["step_roofline"]["frac"],4), d["timing"])
    ks=d.get("kernel_shares") or {}
    print("  libhgb", ks.get("libhgb_share"), "aten", ks.get("aten_share"), ks.get("attribution"))
    for r in (ks.get("by_entry") or [])[:10]: print("  ", r)
    for r in (ks.get("by_kernel") or [])[:8]: print("  k ", r)
except Exception as ex: print("$wl parse failed", ex)
PY
done
