#!/bin/bash
mkdir -p gpurun_out
python tools/bench_build.py 32 > gpurun_out/b5_bench_build.txt 2>&1
RGBID_NO_PDL=1 python tools/bench_build.py 32 > gpurun_out/b5_bench_build_nopdl.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b5_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b5_bench.json 2> gpurun_out/b5_bench.err
RGBID_NO_PDL=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b5_bench_nopdl.json 2> gpurun_out/b5_bench_nopdl.err
tail -n 5 gpurun_out/b5_pytest.txt; cat gpurun_out/b5_bench_build.txt gpurun_out/b5_bench_build_nopdl.txt; cut -c1-330 gpurun_out/b5_bench.json; tail -2 gpurun_out/b5_bench.err; cut -c1-330 gpurun_out/b5_bench_nopdl.json
