#!/bin/bash
# Compiles gn_system.cu alone with extra flags and prints the bank-pressure model of the fast tracker kernel.
# Usage: tools/try_flags.sh <tag> "<extra nvcc flags>"
tag=$1; shift
out=/tmp/try_$tag.o
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --ftz=true --prec-div=false --prec-sqrt=false \
  -Xcompiler -fPIC -Xptxas -v $@ -c rgbid-slam_b200/csrc/gn_system.cu -o $out 2> /tmp/try_$tag.log || { tail -5 /tmp/try_$tag.log; exit 1; }
grep -A2 "Function properties for .*gn_build_fast_kernelILb1ELi0" /tmp/try_$tag.log | tail -2 | tr '\n' ' '; echo
python tools/sass_banks.py $out gn_build_fast_kernelILb1ELi0 ${PX:-8} | tail -2
