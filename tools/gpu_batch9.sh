#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/b9_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b9_bench.json 2> gpurun_out/b9_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/b9_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b9_ncu.log 2>&1
tail -n 8 gpurun_out/b9_pytest.txt; cut -c1-600 gpurun_out/b9_bench.json; tail -3 gpurun_out/b9_bench.err
