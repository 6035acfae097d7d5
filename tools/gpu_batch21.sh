#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
RGBID_NO_PDL=1 RGBID_LIB=$L/librgbid_b200_probe.so timeout 300 python tools/scale_round_probe.py 2>&1 | grep -A1 "tail probe level" | tail -16 > gpurun_out/b21_tail_stages.txt
RGBID_NO_PDL=1 RGBID_LIB=$L/librgbid_b200_probenw.so timeout 300 python tools/scale_round_probe.py 2>&1 | grep -A1 "tail probe level" | tail -8 > gpurun_out/b21_tail_stages_nowarm.txt
for i in 1 2; do python tools/bench_build.py 32 >> gpurun_out/b21_build.txt 2>&1; RGBID_LIB=$L/librgbid_b200_nowarm.so python tools/bench_build.py 32 >> gpurun_out/b21_build_nowarm.txt 2>&1; done
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b21_bench.json 2> gpurun_out/b21_bench.err
RGBID_LIB=$L/librgbid_b200_nowarm.so timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b21_bench_nowarm.json 2> gpurun_out/b21_bench_nowarm.err
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b21_pytest.txt 2>&1
cat gpurun_out/b21_tail_stages.txt gpurun_out/b21_tail_stages_nowarm.txt gpurun_out/b21_build.txt gpurun_out/b21_build_nowarm.txt; tail -3 gpurun_out/b21_pytest.txt; cut -c1-300 gpurun_out/b21_bench.json; cut -c1-300 gpurun_out/b21_bench_nowarm.json
