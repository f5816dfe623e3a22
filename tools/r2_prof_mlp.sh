#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2_prof_mlp}
mkdir -p "$OUT"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'app_mlp_tc5|march_kernel' --launch-skip 4 -c 2 \
  -f -o "$OUT/sec" python tools/profile_target.py > "$OUT/ncu.log" 2>&1
echo "rc=$?" | tee -a "$OUT/summary.txt"
tail -4 "$OUT/ncu.log" | tee -a "$OUT/summary.txt"
ncu -i "$OUT/sec.ncu-rep" --page raw --csv > "$OUT/sec_raw.csv" 2>/dev/null
