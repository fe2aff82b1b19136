#!/bin/bash
# Runs on the GPU box: captures the raw material of profiles/ (post-processed here by tools/summarize_profiles.py).
#   gpurun -- tools/make_profiles.sh
set -x
mkdir -p gpurun_out
# 1. launch list of 60 updates at 1080p (per-launch durations, cold-cache and serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_hd1080_60frames.csv \
    python tools/profile_run.py hd_1920x1080 61 > gpurun_out/launches.log 2>&1
# 2. full capture of ncc_kernel at frame 40 (launch index 39)
ncu --set full --clock-control none --import-source on -k regex:ncc_kernel -s 39 -c 1 -f -o gpurun_out/ncc_kernel_hd1080_frame40 \
    python tools/profile_run.py hd_1920x1080 42 > gpurun_out/ncu_ncc.log 2>&1
# 3. full capture of the other two per-update kernels at frame 40 (advance_kernel = fusion of the previous update + set-up)
ncu --set full --clock-control none --import-source on -k regex:'advance_kernel|moments_kernel' -s 77 -c 2 -f -o gpurun_out/aux_kernels_hd1080_frame40 \
    python tools/profile_run.py hd_1920x1080 42 > gpurun_out/ncu_aux.log 2>&1
tail -2 gpurun_out/ncu_ncc.log gpurun_out/ncu_aux.log
