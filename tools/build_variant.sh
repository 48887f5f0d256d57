#!/bin/bash
# build a tuning variant of libkgb200: tools/build_variant.sh <out.so> [-D... flags]
# (the same flags as keras_geometric_b200/_build.py plus the given defines; objects go to /tmp).
# Load it with KGB200_LIB=<out.so> (same ABI) to compare variants in one GPU call.
set -e
out=$1; shift
cd "$(dirname "$0")/.."
objs=""
for f in keras_geometric_b200/csrc/*.cu; do
  o=/tmp/kgbvar_$(basename $f .cu)_$$.o
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include \
       -I keras_geometric_b200/csrc "$@" -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -shared -o $out $objs
rm -f $objs
echo built $out
