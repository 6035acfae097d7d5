#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
for i in 1 2; do python tools/bench_build.py 32 >> gpurun_out/b38_build.txt 2>&1; RGBID_LIB=$L/librgbid_b200_inl.so python tools/bench_build.py 32 >> gpurun_out/b38_build_inl.txt 2>&1; done
RGBID_LIB=$L/librgbid_b200_inl.so timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b38_bench_inl.json 2> gpurun_out/b38_bench_inl.err
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b38_bench.json 2> gpurun_out/b38_bench.err
cat gpurun_out/b38_build.txt gpurun_out/b38_build_inl.txt; for f in bench bench_inl; do python -c "
import json;d=json.load(open('gpurun_out/b38_$f.json'));print('$f',round(d['value']),round(d['ms_per_step'],4))"; done
