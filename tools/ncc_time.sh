#!/bin/bash
# usage: tools/ncc_time.sh "<nvcc extra flags>" : builds the variant and prints ncc_kernel durations at frames 39-41 (hd1080)
DMF_NVCC_EXTRA="$1" python -m slamplay_b200.build --force 2>&1 | grep -A2 "ncc_kernel" | grep -E "registers|spill" | tr '\n' ' '
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"ncc_kernel" -s 38 -c 2 --csv --log-file gpurun_out/t.csv python tools/profile_run.py hd_1920x1080 42 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/t.csv')) if len(r)>5]
h=rows[0]
print(' | '.join(f"{r[h.index('Metric Name')].split('.')[0][-14:]}={r[h.index('Metric Value')]}" for r in rows[1:]))
PY
