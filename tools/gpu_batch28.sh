#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b28_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b28_bench.json 2> gpurun_out/b28_bench.err
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b28_bench2.json 2> gpurun_out/b28_bench2.err
tail -3 gpurun_out/b28_pytest.txt; cut -c1-300 gpurun_out/b28_bench.json; cut -c1-300 gpurun_out/b28_bench2.json
