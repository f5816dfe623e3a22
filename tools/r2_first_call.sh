#!/bin/bash
# First GPU call of round 2: everything round 1 prepared without being able to run it, in ONE gpurun invocation.
#   /usr/local/graft/bin/gpurun --timeout 3600 -- 'bash tools/r2_first_call.sh'
# Results land in gpurun_out/r2_first/.  Every step has its own timeout and never aborts the following ones.
set -u
OUT=gpurun_out/r2_first
mkdir -p "$OUT"
step() { echo "=== $1" | tee -a "$OUT/summary.txt"; }

step "1. GPU parity suite on main (includes the shadow-refresh fix and the PSNR assertion)"
timeout 700 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"

step "2. default bench on main (first number that contains the per-step shadow repack)"
timeout 500 python bench.py > "$OUT/bench_main.json" 2> "$OUT/bench_main.err"; echo "rc=$?" | tee -a "$OUT/summary.txt"

step "3. ncu launch list of the same command (2 steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
  --log-file "$OUT/launches_main.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_bench.log" 2>&1
echo "rc=$?" | tee -a "$OUT/summary.txt"
python tools/launch_phases.py "$OUT/launches_main.csv" > "$OUT/step_phase_attribution.txt" 2>&1
grep -c pack_channels_last "$OUT/launches_main.csv" | sed 's/^/pack_channels_last launches in the list: /' | tee -a "$OUT/summary.txt"

step "4. tcgen05 probe (descriptor convention, A from TMEM, MMA rate); bounded waits, own timeout"
make -C experiments/umma_probe > "$OUT/umma_build.log" 2>&1
timeout 60 experiments/umma_probe/umma_probe > "$OUT/umma_probe.txt" 2>&1; echo "rc=$?" | tee -a "$OUT/summary.txt"
timeout 60 experiments/umma_probe/umma_probe --swap > "$OUT/umma_probe_swap.txt" 2>&1; echo "rc(swap)=$?" | tee -a "$OUT/summary.txt"
cat "$OUT/umma_probe.txt" | tee -a "$OUT/summary.txt"

step "5. fused tail / epilogue kernels against autograd on the GPU"
timeout 300 python experiments/primary_tail/check_gpu.py > "$OUT/tail_gpu.txt" 2>&1; echo "rc=$?" | tee -a "$OUT/summary.txt"
timeout 300 python experiments/primary_epilogue/check_gpu.py > "$OUT/epilogue_gpu.txt" 2>&1; echo "rc=$?" | tee -a "$OUT/summary.txt"
cat "$OUT/tail_gpu.txt" "$OUT/epilogue_gpu.txt" | tee -a "$OUT/summary.txt"

step "6. channel-last VM parameters (experiments/channels_last_params/product.patch) in a scratch copy"
CL=/tmp/tir_channels_last
rm -rf "$CL"; mkdir -p "$CL"
tar --exclude=./gpurun_out --exclude=./.git -cf - . | tar -xf - -C "$CL"
( cd "$CL" && git apply experiments/channels_last_params/product.patch ) > "$OUT/cl_apply.log" 2>&1; echo "apply rc=$?" | tee -a "$OUT/summary.txt"
( cd "$CL" && timeout 700 python -m pytest tests -q -m gpu -x ) > "$OUT/cl_pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/cl_pytest_gpu.log" | tee -a "$OUT/summary.txt"
( cd "$CL" && timeout 500 python bench.py --no-cpu-baseline ) > "$OUT/bench_channels_last.json" 2> "$OUT/bench_channels_last.err"
echo "bench rc=$?" | tee -a "$OUT/summary.txt"

step "6b. product patches: term = exact early termination in the march; tail = fused primary tail/epilogue; both = tail + channel-last parameters; all = everything incl. batched heads backward"
for variant in term tail both all; do
  D=/tmp/tir_$variant
  rm -rf "$D"; mkdir -p "$D"
  tar --exclude=./gpurun_out --exclude=./.git -cf - . | tar -xf - -C "$D"
  ( cd "$D" && { [ $variant = both -o $variant = all ] && git apply experiments/channels_last_params/product.patch; true; } \
      && { [ $variant != term ] && git apply experiments/primary_tail/product.patch; true; } \
      && { [ $variant = all ] && git apply experiments/batched_heads/product.patch; true; } \
      && { [ $variant = all -o $variant = term ] && git apply experiments/exact_termination/product.patch; true; } \
      && python -c "import __graft_entry__ as g; g.build(force=True)" ) > "$OUT/${variant}_apply_build.log" 2>&1
  echo "$variant apply+build rc=$?" | tee -a "$OUT/summary.txt"
  ( cd "$D" && timeout 700 python -m pytest tests -q -m gpu -x ) > "$OUT/${variant}_pytest_gpu.log" 2>&1
  echo "$variant pytest rc=$?" | tee -a "$OUT/summary.txt"
  tail -3 "$OUT/${variant}_pytest_gpu.log" | tee -a "$OUT/summary.txt"
  ( cd "$D" && timeout 500 python bench.py --no-cpu-baseline ) > "$OUT/bench_$variant.json" 2> "$OUT/bench_$variant.err"
  echo "$variant bench rc=$?" | tee -a "$OUT/summary.txt"
done

step "7. PSNR of a 96x96 crop against the reference algorithm on the same GPU"
timeout 400 python tools/psnr_vs_reference.py 128 96 > "$OUT/psnr.json" 2> "$OUT/psnr.err"; echo "rc=$?" | tee -a "$OUT/summary.txt"
tail -1 "$OUT/psnr.json" | tee -a "$OUT/summary.txt"
echo "done" | tee -a "$OUT/summary.txt"
