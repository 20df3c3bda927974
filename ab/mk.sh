#!/bin/bash
# A/B build of the tensor fast path: ab/mk.sh <tag> <extra nvcc flags...>  ->  ab/lib_<tag>.so
# (only mma_host.o and the distillation unit mma_10_24.o are rebuilt with the flags; everything else from csrc/build)
set -e
tag=$1; shift
R=/root/repo; C=$R/sihmpc_b200/csrc; O=/tmp/exp/$tag; mkdir -p $O
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
cd $C
nvcc $F "$@" -c -o $O/mma_host.o mma_units.cu &
nvcc $F "$@" -DSIH_MMA_N=10 -DSIH_MMA_NXP=24 -Xptxas -v -c -o $O/mma_10_24.o mma_units.cu 2> $O/ptxas.log &
wait
objs=""
for f in $C/build/*.o; do b=$(basename $f); if [ $b = mma_host.o ] || [ $b = mma_10_24.o ]; then objs="$objs $O/$b"; else objs="$objs $f"; fi; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $R/ab/lib_$tag.so $objs
grep -A2 "Li16EEELi[0-9]*ELi[0-9]*ELi0EE" $O/ptxas.log | grep -v "^--" | sed 's/ptxas info    ://' | cut -c1-200
