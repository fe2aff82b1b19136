#!/bin/bash
# Builds kernel variants for A/B runs on the GPU box (development tool):
#   tools/build_variants.sh name1 "<nvcc -D flags>" name2 "<flags>" ...
# -> ab/libdmf_<name>.so (git-ignored, travels with gpurun); run them with
#   python tools/ab_bench.py WORKLOAD FRAMES REPS ab/libdmf_a.so ab/libdmf_b.so[:ENV=VALUE,...]
# The flags reach nvcc through DMF_NVCC_EXTRA (slamplay_b200/build.py), e.g. -DDMF_CHUNK=24 -DDMF_NCC_THREADS=128
# -DDMF_NCC_MIN_BLOCKS=4 -DDMF_ADV_MIN_BLOCKS=5 -DDMF_GROUPED_DIV=0.  The default library is rebuilt at the end.
set -e
cd "$(dirname "$0")/.."
mkdir -p ab
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  DMF_NVCC_EXTRA="$flags" python -m slamplay_b200.build --force > /dev/null
  cp slamplay_b200/libdmf.so ab/libdmf_$name.so
  echo "$name [$flags]: $(grep -A2 'ncc_kernelILi1920ELb0' slamplay_b200/build/dmf_api.ptxas.log | grep -E 'registers|spill' | tr '\n' ' ' | sed 's/ptxas info    ://g; s/  */ /g')"
done
python -m slamplay_b200.build --force > /dev/null
