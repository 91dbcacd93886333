#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "Warning\|warn" | tail -40 | cut -c1-400
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
bash profiles/run_benches.sh r02p md17_egnn lj_egnn qm9_painn gfm_pnaeq oc20_mace
