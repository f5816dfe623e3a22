#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2_c17}
mkdir -p "$OUT"
say() { echo "$@" | tee -a "$OUT/summary.txt"; }
timeout 900 python tools/train_validation.py 500 > "$OUT/train_validation.json" 2> "$OUT/train_validation.err"; say "train validation rc=$?"; tail -3 "$OUT/train_validation.err" | tee -a "$OUT/summary.txt"
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_primary_matches_modular or fused_shade_kernel or valid_sample_lists or tcgen05 or fused_adam or secondary_marches" > "$OUT/sanitizer_$tool.log" 2>&1
  say "sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" "$OUT/sanitizer_$tool.log" | tail -4 | tee -a "$OUT/summary.txt"
done
say done
