#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2_c10}
mkdir -p "$OUT"
say() { echo "$@" | tee -a "$OUT/summary.txt"; }
timeout 900 python -m pytest tests -q -m gpu -x -k "baseline_shape or graph_replay_tracks" > "$OUT/pytest_new.log" 2>&1; say "new tests rc=$?"; tail -40 "$OUT/pytest_new.log" | tee -a "$OUT/summary.txt"
timeout 600 python tools/psnr_vs_reference.py 128 800 > "$OUT/psnr_800.json" 2> "$OUT/psnr.err"; say "psnr rc=$?"; tail -1 "$OUT/psnr_800.json" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/psnr.err" | tee -a "$OUT/summary.txt"
say done
