#!/bin/bash
# Runs bench.py for every BASELINE.json config on one GPU and stores the JSON lines under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
tag=${1:-r02}
for wl in qm9_painn md17_egnn oc20_mace gfm_pnaeq lj_egnn; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${tag}_bench_${wl}.json 2> gpurun_out/${tag}_bench_${wl}.err
  echo "$wl rc=$?"; tail -c 600 gpurun_out/${tag}_bench_${wl}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${wl}.json").read().strip().splitlines()[-1])
    print("$wl", "ms/step", round(d["ms_per_step"],3), "atoms/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "step_frac", round(d["step_roofline"]["frac"],4))
    ks=d.get("kernel_shares") or {}
    print("  libhgb", ks.get("libhgb_share"), "aten", ks.get("aten_share"))
    for r in (ks.get("by_entry") or ks.get("by_kernel") or [])[:8]: print("  ", r)
    print("  roofline", d.get("roofline"))
    print("  cpu", d.get("cpu_baseline"))
except Exception as ex: print("$wl parse failed", ex)
PY
done
