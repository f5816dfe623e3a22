#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2_c14}
mkdir -p "$OUT"
say() { echo "$@" | tee -a "$OUT/summary.txt"; }
timeout 900 python bench.py --no-cpu-baseline --no-torch-reference > "$OUT/bench.json" 2> "$OUT/bench.err"; say "bench rc=$?"; tail -2 "$OUT/bench.err" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --no-cpu-baseline --config 3 --steps 10 > "$OUT/bench_c3.json" 2> "$OUT/bench_c3.err"; say "c3 rc=$?"; tail -2 "$OUT/bench_c3.err" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --no-cpu-baseline --config 3 --envmap_h 16 --envmap_w 32 --steps 10 --no-torch-reference > "$OUT/bench_c3_512.json" 2> "$OUT/bench_c3_512.err"; say "c3_512 rc=$?"; tail -2 "$OUT/bench_c3_512.err" | tee -a "$OUT/summary.txt"
say done
