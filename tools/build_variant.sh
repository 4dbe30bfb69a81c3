#!/bin/bash
# developer aid: build_ab/lib_<name>.so from the current csrc with extra -D flags (only sampler.cu is recompiled unless ALL=1)
# usage: tools/build_variant.sh <name> [nvcc flags...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd); SRC=$ROOT/mcmc-in-tonga_b200/csrc; OUT=$ROOT/build_ab; mkdir -p $OUT
name=$1; shift
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-O2"
cd $SRC
for f in ctx ingest evaluate; do
  if [ ! -f $OUT/$f.o ] || [ $f.cu -nt $OUT/$f.o ] || [ -n "$ALL" ]; then nvcc $F -c $f.cu -o $OUT/$f.o 2>/dev/null & fi
done
nvcc $F "$@" -Xptxas -v -c sampler.cu -o $OUT/sampler_$name.o 2>$OUT/ptxas_$name.log
wait
nvcc $F -shared -o $OUT/lib_$name.so $OUT/ctx.o $OUT/ingest.o $OUT/evaluate.o $OUT/sampler_$name.o
grep -A3 "Compiling entry function '_ZN2tg17tg_sampler_kernelILi3ELb0" $OUT/ptxas_$name.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores\|[0-9]* bytes stack frame" | tr '\n' ' '; echo " -> $OUT/lib_$name.so"
