timeout 700 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25
bash profiles/run_benches.sh r02g md17_egnn gfm_pnaeq 2>&1 | grep -v "^  k " | tail -40
