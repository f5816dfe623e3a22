#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2_c4}
mkdir -p "$OUT"
say() { echo "$@" | tee -a "$OUT/summary.txt"; }
timeout 900 python -m pytest tests -q -m gpu -x ${2:-} > "$OUT/pytest.log" 2>&1; say "pytest rc=$?"; tail -40 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
say done
