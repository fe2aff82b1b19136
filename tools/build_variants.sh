#!/bin/bash
# Builds kernel variants for A/B runs on the GPU box (development tool):
#   tools/build_variants.sh name1 "<nvcc -D flags>" name2 "<flags>" ...
# -> slamplay_b200/build/variants/libdmf_<name>.so  (build/ is git-ignored but travels with gpurun)
set -e
cd "$(dirname "$0")/.."
out=slamplay_b200/build/variants
mkdir -p $out
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -ccbin /usr/bin/g++"
[ -f slamplay_b200/build/synth.o ] || python -m slamplay_b200.build --force > /dev/null
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  $NV $flags -c slamplay_b200/csrc/dmf_api.cu -o $out/dmf_api_$name.o 2> $out/$name.ptxas.log
  $NV -shared -o $out/libdmf_$name.so $out/dmf_api_$name.o slamplay_b200/build/synth.o -lcudart
  rm -f $out/dmf_api_$name.o
  echo "$name [$flags]: $(grep -A2 'ncc_kernel' $out/$name.ptxas.log | grep -E 'registers|spill' | tr '\n' ' ' | sed 's/ptxas info    ://g; s/  */ /g')"
done
