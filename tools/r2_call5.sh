#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2_c5}
mkdir -p "$OUT"
say() { echo "$@" | tee -a "$OUT/summary.txt"; }
timeout 900 python -m pytest tests -q -m gpu -x > "$OUT/pytest.log" 2>&1; say "pytest rc=$?"; tail -30 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"; say "bench rc=$?"; tail -3 "$OUT/bench.err" | tee -a "$OUT/summary.txt"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
  --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_bench.log" 2>&1
say "ncu rc=$?"
python tools/launch_phases.py "$OUT/launches.csv" > "$OUT/step_phase_attribution.txt" 2>&1
say done
