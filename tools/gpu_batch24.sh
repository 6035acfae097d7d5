#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
for v in "" _vis2 _vis4; do
RGBID_LIB=$L/librgbid_b200$v.so RGBID_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:visibility4 -s 1 -c 4 --csv --log-file gpurun_out/b24_vis4$v.csv python tools/profile_step.py 32 6 > /dev/null 2>&1
echo "variant [$v]"; grep visibility4 gpurun_out/b24_vis4$v.csv | awk -F'","' '{print $NF}'
done
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b24_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b24_bench.json 2> gpurun_out/b24_bench.err
tail -3 gpurun_out/b24_pytest.txt; cut -c1-300 gpurun_out/b24_bench.json
