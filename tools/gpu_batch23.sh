#!/bin/bash
mkdir -p gpurun_out
RGBID_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:visibility4 -s 2 -c 1 -o gpurun_out/r02e_vis4 -f python tools/profile_step.py 32 4 > gpurun_out/b23_ncu1.log 2>&1
RGBID_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_build_fast -s 12 -c 1 -o gpurun_out/r02e_fast -f python tools/profile_step.py 32 2 > gpurun_out/b23_ncu2.log 2>&1
tail -2 gpurun_out/b23_ncu1.log gpurun_out/b23_ncu2.log
