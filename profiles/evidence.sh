#!/bin/bash
# Round-2 evidence run (one GPU): ncu launch lists of the bench command per config, ncu --set full of the dominant kernels,
# compute-sanitizer memcheck / racecheck over the new kernels' tests.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
for wl in qm9_painn md17_egnn oc20_mace gfm_pnaeq; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_${wl}.csv \
    python bench.py --workload $wl --steps 2 --warmup 1 --regions 1 --no-graph --skip-cpu-baseline --skip-kernel-shares > gpurun_out/r02_launches_${wl}.log 2>&1
  echo "launch list $wl rc=$?"
done
# dominant kernels, full sets (one launch each, late in the run so that shapes are the bench shapes)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_linear_kernel -s 40 -c 1 -o gpurun_out/r02_ncu_tc_linear \
  python bench.py --workload qm9_painn --steps 2 --warmup 1 --regions 1 --no-graph --skip-cpu-baseline --skip-kernel-shares > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:egnn_edge_fwd_kernel -s 6 -c 1 -o gpurun_out/r02_ncu_egnn_edge_fwd \
  python bench.py --workload md17_egnn --steps 2 --warmup 1 --regions 1 --no-graph --skip-cpu-baseline --skip-kernel-shares > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:egnn_edge_wgrad_kernel -s 6 -c 1 -o gpurun_out/r02_ncu_egnn_edge_wgrad \
  python bench.py --workload md17_egnn --steps 2 --warmup 1 --regions 1 --no-graph --skip-cpu-baseline --skip-kernel-shares > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mha_tc_bwd_kv_kernel -s 2 -c 1 -o gpurun_out/r02_ncu_mha_tc_bwd_kv \
  python bench.py --workload gfm_pnaeq --steps 2 --warmup 1 --regions 1 --no-graph --skip-cpu-baseline --skip-kernel-shares > /dev/null 2>&1
for f in tc_linear egnn_edge_fwd egnn_edge_wgrad mha_tc_bwd_kv; do
  ncu -i gpurun_out/r02_ncu_${f}.ncu-rep --page details --csv > gpurun_out/r02_ncu_${f}_details.csv 2>/dev/null
  ncu -i gpurun_out/r02_ncu_${f}.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
if len(rows)>2:
    h,v=rows[0],rows[-1]
    keep=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor','sm__inst_executed_pipe_tensor','sm__throughput','dram__throughput','sm__warps_active','smsp__inst_executed.sum','launch__registers_per_thread','sm__inst_executed_pipe_fma','l1tex__data_bank_conflicts']
    for k,x in zip(h,v):
        if any(s in k for s in keep): print(k, x)
" > gpurun_out/r02_ncu_${f}_key_metrics.txt
  rm -f gpurun_out/r02_ncu_${f}.ncu-rep
done
# compute-sanitizer over the kernels added in round 2
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_round2.py -q -x \
  -k "fused_egnn or edge_len or tensor_core_attention or grouped or known_edge or csr_build or pna_aggregate or train_fast" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_round2.py -q -x \
  -k "fused_egnn and md17_egnn-24 or tensor_core_attention and 300" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/r02_sanitizer_racecheck.log
