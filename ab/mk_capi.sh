#!/bin/bash
# A/B build of the kernels in capi.cu (general, log, quad): ab/mk_capi.sh <tag> <extra nvcc flags...>  ->  ab/lib_<tag>.so
set -e
tag=$1; shift
R=/root/repo; C=$R/sihmpc_b200/csrc; O=/tmp/exp/$tag; mkdir -p $O
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
cd $C
nvcc $F "$@" -Xptxas -v -c -o $O/capi.o capi.cu 2> $O/ptxas.log
objs=""
for f in $C/build/*.o; do b=$(basename $f); if [ $b = capi.o ]; then objs="$objs $O/$b"; else objs="$objs $f"; fi; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $R/ab/lib_$tag.so $objs
grep -A2 "ihmpc_log_kernelILi1ELi4" $O/ptxas.log | grep -v "^--" | sed 's/ptxas info    ://' | cut -c1-160
