python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -12
for wl in qm9_painn md17_egnn; do
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload $wl --steps 10 --warmup 3 --skip-cpu-baseline --skip-kernel-shares > gpurun_out/r02_bench_2gpu_$wl.json 2> gpurun_out/r02_bench_2gpu_$wl.err; echo "$wl rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_2gpu_$wl.json').read().strip().splitlines()[-1]); print('2gpu $wl', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['launch'], d['timing'])"
done
bash profiles/run_benches.sh r02e md17_egnn gfm_pnaeq lj_egnn 2>&1 | grep -v "^  k " | tail -45
