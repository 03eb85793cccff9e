#!/bin/bash
# build_variant.sh <name> "<extra nvcc flags>" [file.cu ...]: rebuild the named sources (default gnn_tc.cu) with extra -D flags
# and link them with the objects of the regular build into ofdft_nflows_b200/variants/lib_<name>.so (development tool).
set -e
cd "$(dirname "$0")/../../ofdft_nflows_b200/csrc"
name=$1; flags=$2; shift 2 || true
files=${@:-gnn_tc.cu}
mkdir -p ../variants /tmp/abobj_$name
objs=""
for f in gnn.cu gnn_tc.cu gnn_general.cu functionals.cu mlp1d.cu mlp1d_tc.cu prior.cu tc_test.cu; do
  if [[ " $files " == *" $f "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $flags -c $f -o /tmp/abobj_$name/${f%.cu}.o 2>&1 | grep -A2 "bwd_tc_kernel\|error" | grep -E "spill|registers|error" || true
    objs="$objs /tmp/abobj_$name/${f%.cu}.o"
  else
    objs="$objs ${f%.cu}.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$name.so $objs
echo "built variants/lib_$name.so"
