timeout 700 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15
bash profiles/run_benches.sh r02h qm9_painn md17_egnn oc20_mace gfm_pnaeq lj_egnn 2>&1 | grep -v "^  k " | tail -90
