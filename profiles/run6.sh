python profiles/diag_r2.py wd 2>&1 | grep -v Warn | tail -14
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu-baseline --skip-kernel-shares > gpurun_out/r02_bench_2gpu_qm9.json 2> gpurun_out/r02_bench_2gpu_qm9.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_2gpu_qm9.err | grep -v Warn
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_2gpu_qm9.json').read().strip().splitlines()[-1]); print('2gpu qm9', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['launch'], d['timing'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload oc20_mace --steps 10 --warmup 3 --skip-cpu-baseline --skip-kernel-shares > gpurun_out/r02_bench_2gpu_oc20.json 2> gpurun_out/r02_bench_2gpu_oc20.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_2gpu_oc20.err | grep -v Warn
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_2gpu_oc20.json').read().strip().splitlines()[-1]); print('2gpu oc20', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['launch'], d['timing'])"
