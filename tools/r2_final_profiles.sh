#!/bin/bash
# Round-2 evidence for profiles/: launch list of the bench step, ncu --set full of the two secondary kernels (march,
# tcgen05 MLP) and of the primary heads kernels, on the committed build.
set -u
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${1:-r2_final}_pytest.log 2>&1; echo "pytest rc=$?"
OUT=gpurun_out/${1:-r2_final}
mkdir -p "$OUT"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
  --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-reference > "$OUT/ncu_bench.log" 2>&1
python tools/launch_phases.py "$OUT/launches.csv" > "$OUT/step_phase_attribution.txt" 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'app_mlp_tc5|march_kernel' --launch-skip 4 -c 2 \
  -f -o "$OUT/sec" python tools/profile_target.py > "$OUT/ncu_sec.log" 2>&1
ncu -i "$OUT/sec.ncu-rep" --page raw --csv > "$OUT/sec_raw.csv" 2>/dev/null
ncu -i "$OUT/sec.ncu-rep" --page details > "$OUT/sec_details.txt" 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:'heads_|app_mlp_kernel' --launch-skip 6 -c 3 \
  -f -o "$OUT/prim" python tools/profile_primary.py > "$OUT/ncu_prim.log" 2>&1
ncu -i "$OUT/prim.ncu-rep" --page raw --csv > "$OUT/prim_raw.csv" 2>/dev/null
timeout 900 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference_arm.json" 2> "$OUT/bench_reference_arm.err"
echo done
