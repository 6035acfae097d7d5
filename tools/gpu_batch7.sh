#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/b7_bench.json 2> gpurun_out/b7_bench.err
timeout 300 python bench.py --mode align --steps 10 > gpurun_out/b7_bench_align.json 2> gpurun_out/b7_bench_align.err
tail -5 gpurun_out/b7_bench.err; cat gpurun_out/b7_bench.json; tail -3 gpurun_out/b7_bench_align.err; cat gpurun_out/b7_bench_align.json
