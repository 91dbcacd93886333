#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_round2.py -m gpu -q --tb=short -x -k "tc_ or side_stream" 2>&1 | tail -15
timeout 200 python profiles/diag_wgrad.py 2>&1 | grep -v Warn | tail -8
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12
bash profiles/run_benches.sh r02n md17_egnn gfm_pnaeq lj_egnn qm9_painn
