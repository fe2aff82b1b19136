#!/bin/bash
# Runs on the GPU box: raw material of profiles/r02_* (post-processed by tools/summarize_profiles_r02.py).
#   gpurun -- tools/make_profiles_r02.sh
set -x
mkdir -p gpurun_out/r02
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_fma.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__registers_per_thread,sm__cycles_elapsed.avg
# 1. launch list of 60 updates at 1080p (per-launch durations are cold-cache and serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches_hd1080_60frames.csv \
    python tools/profile_run.py hd_1920x1080 61 > gpurun_out/r02/launches.log 2>&1
# 2. pipe counts of the ncc_kernel launch of update 40 at 1080p and 4K (+ the work counters of exactly that update)
for WL in hd_1920x1080 uhd_3840x2160; do
  ncu --metrics $M --clock-control none -k regex:ncc_kernel -s 39 -c 1 --csv --log-file gpurun_out/r02/ncc_pipes_$WL.csv \
      python tools/profile_run.py $WL 42 40 > gpurun_out/r02/ncc_pipes_$WL.log 2>&1
done
# 3. full captures: ncc_kernel at update 40 (1080p and 4K), advance / moments at 1080p
ncu --set full --clock-control none --import-source on -k regex:ncc_kernel -s 39 -c 1 -f -o gpurun_out/r02/ncc_kernel_hd1080_update40 \
    python tools/profile_run.py hd_1920x1080 42 > gpurun_out/r02/ncu_ncc_hd.log 2>&1
ncu --set full --clock-control none -k regex:ncc_kernel -s 39 -c 1 -f -o gpurun_out/r02/ncc_kernel_uhd2160_update40 \
    python tools/profile_run.py uhd_3840x2160 42 > gpurun_out/r02/ncu_ncc_uhd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'advance_kernel|moments_kernel' -s 77 -c 2 -f -o gpurun_out/r02/aux_kernels_hd1080_update40 \
    python tools/profile_run.py hd_1920x1080 42 > gpurun_out/r02/ncu_aux.log 2>&1
tail -2 gpurun_out/r02/*.log
