#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
for i in 1 2; do python tools/bench_build.py 32 >> gpurun_out/b12_build.txt 2>&1; done
python tools/bench_build.py 1 >> gpurun_out/b12_build.txt 2>&1
python tools/bench_build.py 32 align >> gpurun_out/b12_build.txt 2>&1
RGBID_LIB=$L/librgbid_b200_probe.so timeout 300 python tools/scale_round_probe.py 2>&1 | tail -22 > gpurun_out/b12_tail_probe.txt
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b12_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b12_bench.json 2> gpurun_out/b12_bench.err
cat gpurun_out/b12_build.txt; tail -8 gpurun_out/b12_tail_probe.txt; tail -n 5 gpurun_out/b12_pytest.txt; cut -c1-330 gpurun_out/b12_bench.json
