#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02m_bench_2gpu_qm9_painn.json 2> gpurun_out/r02m_bench_2gpu.err
echo "rc=$?"; tail -c 600 gpurun_out/r02m_bench_2gpu.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02m_bench_2gpu_qm9_painn.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches","clocks")}, d["e2e"], d["timing"], d.get("cpu_baseline"), (d.get("kernel_shares") or {}).get("libhgb_share"))
PY
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -3 | cut -c1-600
