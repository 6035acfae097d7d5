#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
python tools/bench_build.py 32 > gpurun_out/b48_build.txt 2>&1
RGBID_LIB=$L/librgbid_b200_flog.so python tools/bench_build.py 32 > gpurun_out/b48_build_flog.txt 2>&1
RGBID_LIB=$L/librgbid_b200_flog.so timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b48_bench_flog.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b48_bench.json 2>/dev/null
RGBID_LIB=$L/librgbid_b200_flog.so timeout 600 python tools/stress_parity.py 777 2 2>&1 | tail -3 > gpurun_out/b48_stress_flog.txt
cat gpurun_out/b48_build.txt gpurun_out/b48_build_flog.txt gpurun_out/b48_stress_flog.txt; for f in bench bench_flog; do python -c "
import json;d=json.load(open('gpurun_out/b48_$f.json'));print('$f',round(d['value']),round(d['ms_per_step'],4))"; done
