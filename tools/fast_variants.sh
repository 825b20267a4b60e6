#!/bin/bash
# Builds variants of liborbfront_b200.so that differ only in -D switches of ONE source file (default csrc/fast.cu), into
# build/variants/ (travels with gpurun, git-ignored).  usage: tools/fast_variants.sh [-s file.cu] name1:"-DFV_A=1 -DFV_B=0" name2:"..." ...
set -e
cd "$(dirname "$0")/.."
SRC=fast.cu
if [ "$1" = "-s" ]; then SRC=$2; shift 2; fi
CS=adaptive-rgbd-localization-mappig_b200/csrc
OUT=build/variants; mkdir -p $OUT/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off"
for f in context pyramid fast quadtree describe match ransac kfdb comm adaptive projection c_abi; do
  [ "$f.cu" = "$SRC" ] && continue
  newest=$(ls -t $CS/*.h $CS/*.inc include/orbfront.h $CS/$f.cu | head -1)
  if [ ! -f $OUT/obj/$f.o ] || [ $newest -nt $OUT/obj/$f.o ]; then nvcc $FLAGS -c $CS/$f.cu -o $OUT/obj/$f.o & fi
done
wait
OBJS=""
for f in context pyramid fast quadtree describe match ransac kfdb comm adaptive projection c_abi; do [ "$f.cu" = "$SRC" ] || OBJS="$OBJS $OUT/obj/$f.o"; done
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  ( nvcc $FLAGS $defs -Xptxas -v -c $CS/$SRC -o $OUT/obj/var_$name.o 2> $OUT/var_$name.log; grep -A2 "Function properties" $OUT/var_$name.log | grep -E "registers|spill" | head -4 | tr '\n' ' '; echo " <- $name"
    nvcc $FLAGS -shared -o $OUT/lib_$name.so $OBJS $OUT/obj/var_$name.o -ldl ) &
done
wait
ls -la $OUT/*.so
