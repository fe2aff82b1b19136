#!/bin/bash
# usage: tools/kern_time.sh "<nvcc extra flags>" : builds the variant and prints all per-frame kernel durations at frame 40 (hd1080)
DMF_NVCC_EXTRA="$1" python -m slamplay_b200.build --force 2>&1 | grep -E "Compiling entry|registers" | grep -A1 -E "advance_kernel|ncc_kernel" | grep registers | tr '\n' ' '
echo
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"advance_kernel|moments_kernel|ncc_kernel" -s 117 -c 3 --csv --log-file gpurun_out/t.csv python tools/profile_run.py hd_1920x1080 45 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/t.csv')) if len(r)>5]
h=rows[0]
print(' | '.join(f"{r[h.index('Kernel Name')].split('(')[0]}={float(r[h.index('Metric Value')])/1e3:.1f}us" for r in rows[1:]))
PY
