#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2_c8}
mkdir -p "$OUT"
say() { echo "$@" | tee -a "$OUT/summary.txt"; }
timeout 120 python -m pytest tests -q -m gpu -x -k "tcgen05" > "$OUT/pytest_tc5.log" 2>&1; say "tc5 test rc=$?"; tail -30 "$OUT/pytest_tc5.log" | tee -a "$OUT/summary.txt"
timeout 600 python -m pytest tests -q -m gpu > "$OUT/pytest.log" 2>&1; say "pytest rc=$?"; tail -5 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"; say "bench rc=$?"; tail -3 "$OUT/bench.err" | tee -a "$OUT/summary.txt"
bash tools/r2_prof_mlp.sh ${1:-r2_c8}
say done
