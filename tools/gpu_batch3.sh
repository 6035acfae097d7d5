#!/bin/bash
mkdir -p gpurun_out
python tools/bench_build.py 32 > gpurun_out/b3_bench_build.txt 2>&1
RGBID_LIB=$PWD/rgbid-slam_b200/lib/librgbid_b200_probe.so timeout 300 python tools/scale_round_probe.py 2>&1 | tail -22 > gpurun_out/b3_tail_probe.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=15 > gpurun_out/b3_pytest.txt 2>&1
timeout 600 python bench.py > gpurun_out/b3_bench.json 2> gpurun_out/b3_bench.err
RGBID_CHAINS=2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b3_bench_chains.json 2> gpurun_out/b3_bench_chains.err
tail -n 30 gpurun_out/b3_pytest.txt; cat gpurun_out/b3_bench_build.txt; tail -12 gpurun_out/b3_tail_probe.txt; cut -c1-400 gpurun_out/b3_bench.json; cut -c1-300 gpurun_out/b3_bench_chains.json; tail -3 gpurun_out/b3_bench_chains.err
