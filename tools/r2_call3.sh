#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2_c3}
mkdir -p "$OUT"
say() { echo "$@" | tee -a "$OUT/summary.txt"; }
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest.log" 2>&1; say "pytest rc=$?"; tail -25 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"; say "bench rc=$?"; tail -c 1200 "$OUT/bench.json" | tee -a "$OUT/summary.txt"; tail -5 "$OUT/bench.err" | tee -a "$OUT/summary.txt"
say done
