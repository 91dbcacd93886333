timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12
bash profiles/evidence.sh 2>&1 | tail -30
bash profiles/run_benches.sh r02i md17_egnn lj_egnn 2>&1 | grep -v "^  k " | tail -34
