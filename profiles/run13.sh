timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12
bash profiles/run_benches.sh r02j qm9_painn gfm_pnaeq oc20_mace 2>&1 | grep -v "^  k " | tail -48
