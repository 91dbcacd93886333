#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -x -k "fast_path and md17" 2>&1 | grep -v Warning | tail -40 | cut -c1-600
echo ==== EXACT_WGRAD=0
HGB_EXACT_WGRAD=0 timeout 200 python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=line -x -k "fast_path and md17" 2>&1 | tail -4 | cut -c1-400
