#!/bin/bash
timeout 700 python -m pytest tests -m gpu -q --tb=short -x -k "side_stream or oc20_mace_shape_engine or gfm_pnaeq_gps or attention or tn or gemm" 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
bash profiles/run_benches.sh r02k qm9_painn gfm_pnaeq oc20_mace
