#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b8_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b8_bench.json 2> gpurun_out/b8_bench.err
tail -n 6 gpurun_out/b8_pytest.txt; cut -c1-420 gpurun_out/b8_bench.json; tail -3 gpurun_out/b8_bench.err
