#!/bin/bash
mkdir -p gpurun_out
python tools/bench_build.py 32 > gpurun_out/b2_bench_build.txt 2>&1
RGBID_LIB=$PWD/rgbid-slam_b200/lib/librgbid_b200_probe.so timeout 300 python tools/scale_round_probe.py 2>&1 | tail -22 > gpurun_out/b2_tail_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b2_pytest.txt 2>&1
timeout 600 python bench.py > gpurun_out/b2_bench.json 2> gpurun_out/b2_bench.err
tail -n 5 gpurun_out/b2_pytest.txt; cat gpurun_out/b2_bench_build.txt gpurun_out/b2_tail_probe.txt gpurun_out/b2_bench.json
