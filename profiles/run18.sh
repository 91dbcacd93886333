#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_round2.py -m gpu -q --tb=short -x -k "tc_ or zero_padded" 2>&1 | tail -15
timeout 200 python profiles/diag_wgrad.py 2>&1 | grep -v Warn | tail -8
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12
bash profiles/run_benches.sh r02o md17_egnn lj_egnn gfm_pnaeq oc20_mace qm9_painn
