timeout 600 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -25
bash profiles/run_benches.sh r02g md17_egnn gfm_pnaeq lj_egnn 2>&1 | grep -v "^  k " | tail -50
