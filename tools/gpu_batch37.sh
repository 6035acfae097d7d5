#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
python tools/bench_build.py 32 > gpurun_out/b37_build.txt 2>&1
RGBID_LIB=$L/librgbid_b200_c16.so python tools/bench_build.py 32 > gpurun_out/b37_build_c16.txt 2>&1
RGBID_LIB=$L/librgbid_b200_c16.so timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b37_bench_c16.json 2> gpurun_out/b37_bench_c16.err
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b37_bench.json 2> gpurun_out/b37_bench.err
RGBID_LIB=$L/librgbid_b200_c16.so timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b37_pytest_c16.txt 2>&1
cat gpurun_out/b37_build.txt gpurun_out/b37_build_c16.txt; tail -3 gpurun_out/b37_pytest_c16.txt; for f in bench bench_c16; do python -c "
import json;d=json.load(open('gpurun_out/b37_$f.json'));print('$f',round(d['value']),round(d['ms_per_step'],4))"; done; tail -2 gpurun_out/b37_bench_c16.err
