#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
RGBID_NO_PDL=1 RGBID_LIB=$L/librgbid_b200_probe.so timeout 300 python tools/scale_round_probe.py 2>&1 | grep -A2 "tail probe level" | tail -30 > gpurun_out/b20_tail_stages.txt
for i in 1 2; do python tools/bench_build.py 32 >> gpurun_out/b20_build.txt 2>&1; done
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b20_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b20_bench.json 2> gpurun_out/b20_bench.err
cat gpurun_out/b20_tail_stages.txt gpurun_out/b20_build.txt; tail -3 gpurun_out/b20_pytest.txt; cut -c1-300 gpurun_out/b20_bench.json
