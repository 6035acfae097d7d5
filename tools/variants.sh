#!/bin/bash
# Builds kernel variants on the GPU box and times them (tools/bench_build.py).  Usage: tools/variants.sh "<flags1>" "<flags2>" ...
# Cheaper in GPU minutes: build the variant HERE next to the product library and only run it on the box --
#   RGBID_BUILD_TAG=mlp RGBID_EXTRA_NVCC_FLAGS="-DRGBID_SCALE_MLP=1" python rgbid-slam_b200/build.py
#   gpurun -- 'RGBID_LIB=$PWD/rgbid-slam_b200/lib/librgbid_b200_mlp.so python tools/bench_build.py 32; \
#              RGBID_LIB=$PWD/rgbid-slam_b200/lib/librgbid_b200_mlp.so python -m pytest tests -m gpu -x -q' 
for v in "$@"; do
  echo "== variant: [$v]"
  touch rgbid-slam_b200/csrc/gn_system.cu
  RGBID_EXTRA_NVCC_FLAGS="$v" python rgbid-slam_b200/build.py > /dev/null 2>&1 || { echo build failed; continue; }
  grep -A2 "gn_build_kernelILi4ELb0ELb1" rgbid-slam_b200/build/gn_system.cu.o.log | grep -E "Used|spill" | tr '\n' ' '; echo
  python tools/bench_build.py 32
  python tools/bench_build.py 1
done
