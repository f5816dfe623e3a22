#!/bin/bash
# Round-2 first GPU call: top of the applied patch stack; bisect over prefixes only if the parity suite fails.
set -u
OUT=gpurun_out/r2_c1
mkdir -p "$OUT"
say() { echo "$@" | tee -a "$OUT/summary.txt"; }
say "=== top of stack: parity suite"
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_top.log" 2>&1; RC=$?; say "rc=$RC"; tail -15 "$OUT/pytest_top.log" | tee -a "$OUT/summary.txt"
say "=== top of stack: bench"
timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench_top.json" 2> "$OUT/bench_top.err"; say "rc=$?"; tail -c 3000 "$OUT/bench_top.json" | tee -a "$OUT/summary.txt"; tail -5 "$OUT/bench_top.err" | tee -a "$OUT/summary.txt"
say "=== umma probe"
make -C experiments/umma_probe > "$OUT/umma_build.log" 2>&1
timeout 60 experiments/umma_probe/umma_probe > "$OUT/umma_probe.txt" 2>&1; say "rc=$?"
timeout 60 experiments/umma_probe/umma_probe --swap > "$OUT/umma_probe_swap.txt" 2>&1; say "rc(swap)=$?"
cat "$OUT/umma_probe.txt" "$OUT/umma_probe_swap.txt" | tee -a "$OUT/summary.txt"
if [ "$RC" != "0" ]; then
  say "=== bisect over prefixes"
  D=/tmp/tir_stack; rm -rf "$D"; mkdir -p "$D"; tar -xf .stack_base.tar -C "$D"
  for patch in "" experiments/stack/*.patch; do
    tag=base; [ -n "$patch" ] && tag=$(basename "$patch" .patch)
    if [ -n "$patch" ]; then ( cd "$D" && git apply "$OLDPWD/$patch" ) > "$OUT/${tag}_apply.log" 2>&1; fi
    ( cd "$D" && python -c "import __graft_entry__ as g; g.build(force=True)" ) > "$OUT/${tag}_build.log" 2>&1
    ( cd "$D" && timeout 700 python -m pytest tests -q -m gpu ) > "$OUT/${tag}_pytest.log" 2>&1
    say "$tag pytest rc=$?"; tail -4 "$OUT/${tag}_pytest.log" | tee -a "$OUT/summary.txt"
    ( cd "$D" && timeout 500 python bench.py --no-cpu-baseline ) > "$OUT/bench_${tag}.json" 2> "$OUT/bench_${tag}.err"
    say "$tag bench rc=$?"; tail -c 600 "$OUT/bench_${tag}.json" | tee -a "$OUT/summary.txt"
  done
fi
say "=== ncu launch list of the step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
  --log-file "$OUT/launches_top.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_bench.log" 2>&1
say "rc=$?"
python tools/launch_phases.py "$OUT/launches_top.csv" > "$OUT/step_phase_attribution.txt" 2>&1
say "done"
