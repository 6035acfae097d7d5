#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
for i in 1 2; do
python tools/bench_build.py 32 >> gpurun_out/b11_build.txt 2>&1
RGBID_LIB=$L/librgbid_b200_wtex.so python tools/bench_build.py 32 >> gpurun_out/b11_build_wtex.txt 2>&1
done
python tools/bench_build.py 32 align >> gpurun_out/b11_build.txt 2>&1
RGBID_LIB=$L/librgbid_b200_wtex.so python tools/bench_build.py 32 align >> gpurun_out/b11_build_wtex.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b11_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b11_bench.json 2> gpurun_out/b11_bench.err
RGBID_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_build_fast -s 12 -c 1 -o gpurun_out/r02a_fast -f python tools/profile_step.py 32 2 > gpurun_out/b11_ncu.log 2>&1
cat gpurun_out/b11_build.txt gpurun_out/b11_build_wtex.txt; tail -n 5 gpurun_out/b11_pytest.txt; cut -c1-330 gpurun_out/b11_bench.json; tail -2 gpurun_out/b11_ncu.log
