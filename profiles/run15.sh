#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12
bash profiles/run_benches.sh r02l qm9_painn gfm_pnaeq md17_egnn
