#!/bin/bash
# Builds an experimental variant of the library: recompiles the listed sources with extra -D flags and links them
# with the production objects.   usage: tools/build_variant.sh <name> "<-Dflags>" <src.cu> [src.cu ...]
name=$1; flags=$2; shift 2
d=veils_b200/build; mkdir -p $d/var_$name
objs=""
for o in $d/*.o; do
  b=$(basename $o .o); skip=0
  for s in "$@"; do [ "$b" == "$s" ] && skip=1; done
  [ $skip == 0 ] && objs="$objs $o"
done
for s in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3 $flags -c veils_b200/csrc/$s -o $d/var_$name/$s.o &
  objs="$objs $d/var_$name/$s.o"
done
wait
nvcc -shared -o $d/libveils_var_$name.so $objs -gencode arch=compute_100a,code=sm_100a && echo built $d/libveils_var_$name.so
