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

step "6. staged product patches (experiments/stack/): cumulative prefixes 1 exact termination, 2 + lean march, 3 + channel-last parameters, 4 + fused tail/epilogue, 5 + batched heads backward"
D=/tmp/tir_stack
rm -rf "$D"; mkdir -p "$D"
tar --exclude=./gpurun_out --exclude=./.git -cf - . | tar -xf - -C "$D"
for patch in experiments/stack/*.patch; do
  tag=$(basename "$patch" .patch)
  ( cd "$D" && git apply "$patch" && python -c "import __graft_entry__ as g; g.build(force=True)" ) > "$OUT/${tag}_apply_build.log" 2>&1
  echo "$tag apply+build rc=$?" | tee -a "$OUT/summary.txt"
  ( cd "$D" && timeout 700 python -m pytest tests -q -m gpu -x ) > "$OUT/${tag}_pytest_gpu.log" 2>&1
  echo "$tag pytest rc=$?" | tee -a "$OUT/summary.txt"
  tail -3 "$OUT/${tag}_pytest_gpu.log" | tee -a "$OUT/summary.txt"
  ( cd "$D" && timeout 500 python bench.py --no-cpu-baseline ) > "$OUT/bench_${tag}.json" 2> "$OUT/bench_${tag}.err"
  echo "$tag bench rc=$?" | tee -a "$OUT/summary.txt"
  python - "$OUT/bench_${tag}.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print(f"   {d['ms_per_step']:.3f} ms/step, {d['value'] / 1e6:.1f} M rays/s, launches {d['gpu_launches']}, march {d['roofline']['ms_per_launch']:.3f} ms")
except Exception as e:
    print("   no bench line:", e)
PY
done

step "7. PSNR of a 96x96 crop against the reference algorithm on the same GPU"
timeout 400 python tools/psnr_vs_reference.py 128 96 > "$OUT/psnr.json" 2> "$OUT/psnr.err"; echo "rc=$?" | tee -a "$OUT/summary.txt"
tail -1 "$OUT/psnr.json" | tee -a "$OUT/summary.txt"
echo "done" | tee -a "$OUT/summary.txt"
