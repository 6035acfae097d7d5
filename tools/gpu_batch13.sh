#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
for i in 1 2; do python tools/bench_build.py 32 >> gpurun_out/b13_build.txt 2>&1; RGBID_LIB=$L/librgbid_b200_nowarm.so python tools/bench_build.py 32 >> gpurun_out/b13_build_nowarm.txt 2>&1; done
RGBID_NO_PDL=1 python tools/bench_build.py 32 >> gpurun_out/b13_build_nopdl.txt 2>&1
RGBID_NO_PDL=1 RGBID_LIB=$L/librgbid_b200_probe.so timeout 300 python tools/scale_round_probe.py 2>&1 | grep -E "level 0|level 2" | tail -14 > gpurun_out/b13_tail_probe.txt
RGBID_NO_PDL=1 RGBID_LIB=$L/librgbid_b200_probenw.so timeout 300 python tools/scale_round_probe.py 2>&1 | grep -E "level 0|level 2" | tail -8 > gpurun_out/b13_tail_probe_nowarm.txt
cat gpurun_out/b13_build.txt gpurun_out/b13_build_nowarm.txt gpurun_out/b13_build_nopdl.txt gpurun_out/b13_tail_probe.txt gpurun_out/b13_tail_probe_nowarm.txt
