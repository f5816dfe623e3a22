#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2_prof}
mkdir -p "$OUT"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'heads_|app_mlp_kernel' --launch-skip 6 -c 3 \
  -f -o "$OUT/prim" python tools/profile_primary.py > "$OUT/ncu.log" 2>&1
echo "rc=$?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/ncu.log" | tee -a "$OUT/summary.txt"
ncu -i "$OUT/prim.ncu-rep" --page raw --csv > "$OUT/prim_raw.csv" 2>/dev/null
ls -la "$OUT" | tee -a "$OUT/summary.txt"
