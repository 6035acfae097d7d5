#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
./tools/ubench/texpat > gpurun_out/b10_texpat.txt 2>&1
for i in 1 2; do
python tools/bench_build.py 32 >> gpurun_out/b10_build.txt 2>&1
RGBID_LIB=$L/librgbid_b200_nohoist.so python tools/bench_build.py 32 >> gpurun_out/b10_build_nohoist.txt 2>&1
done
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/b10_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b10_bench.json 2> gpurun_out/b10_bench.err
cat gpurun_out/b10_texpat.txt gpurun_out/b10_build.txt gpurun_out/b10_build_nohoist.txt; tail -n 5 gpurun_out/b10_pytest.txt; cut -c1-330 gpurun_out/b10_bench.json
