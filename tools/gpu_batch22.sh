#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b22_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-extras --warp-order warpFirst > gpurun_out/b22_bench_warpfirst.json 2> gpurun_out/b22_bench_warpfirst.err
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b22_bench.json 2> gpurun_out/b22_bench.err
tail -3 gpurun_out/b22_pytest.txt; cut -c1-300 gpurun_out/b22_bench_warpfirst.json; cut -c1-300 gpurun_out/b22_bench.json
