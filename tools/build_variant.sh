#!/bin/bash
# build a tuning variant of libkgb200: tools/build_variant.sh <out.so> [-D... flags]
# (the same flags as keras_geometric_b200/_build.py plus the given defines; objects go to /tmp)
set -e
out=$1; shift
cd "$(dirname "$0")/.."
CUT=$(python -c "from keras_geometric_b200 import _build; print(_build._cutlass_include() or '')")
objs=""
for f in keras_geometric_b200/csrc/*.cu; do
  o=/tmp/kgbvar_$(basename $f .cu)_$$.o
  inc=""
  case "$f" in *dense_gemm_api.cu) ;; *dense_gemm_*) inc="--expt-relaxed-constexpr -I $CUT" ;; esac
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include \
       -I keras_geometric_b200/csrc $inc "$@" -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -shared -o $out $objs -lcuda
rm -f $objs
echo built $out
