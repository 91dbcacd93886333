python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25
bash profiles/run_benches.sh r02f qm9_painn oc20_mace 2>&1 | grep -v "^  k " | tail -36
