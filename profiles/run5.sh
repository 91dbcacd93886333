python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -40
python profiles/diag_r2.py egnn attn 2>&1 | grep -v Warn | grep "edge_mlp.0.weight\|attention\|loss"
bash profiles/run_benches.sh r02d md17_egnn gfm_pnaeq lj_egnn 2>&1 | grep -v "^  k " | tail -60
