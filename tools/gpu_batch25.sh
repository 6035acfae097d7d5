#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
for v in "" _v45 _v46 _v26 _v28; do
RGBID_LIB=$L/librgbid_b200$v.so RGBID_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:visibility4 -s 1 -c 3 --csv --log-file gpurun_out/b25_vis4$v.csv python tools/profile_step.py 32 5 > /dev/null 2>&1
echo "variant [$v]" $(grep visibility4 gpurun_out/b25_vis4$v.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')
done
