#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
python tools/bench_build.py 32 > gpurun_out/b52_build.txt 2>&1
RGBID_LIB=$L/librgbid_b200_shfl.so python tools/bench_build.py 32 > gpurun_out/b52_build_shfl.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b52_bench.json 2>/dev/null
RGBID_LIB=$L/librgbid_b200_shfl.so timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b52_bench_shfl.json 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/b52_pytest.txt 2>&1
cat gpurun_out/b52_build.txt gpurun_out/b52_build_shfl.txt; tail -n 3 gpurun_out/b52_pytest.txt; for f in bench bench_shfl; do python -c "
import json;d=json.load(open('gpurun_out/b52_$f.json'));print('$f',round(d['value']),round(d['ms_per_step'],4))"; done
