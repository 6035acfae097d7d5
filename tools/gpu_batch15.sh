#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --no-extras --warp-order warpFirst > gpurun_out/b15_bench_warpfirst.json 2> gpurun_out/b15_bench_warpfirst.err
timeout 600 python bench.py --no-cpu-baseline --no-extras --mode align > gpurun_out/b15_bench_align.json 2> gpurun_out/b15_bench_align.err
timeout 900 python bench.py > gpurun_out/b15_bench_full.json 2> gpurun_out/b15_bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b15_bench_ref.json 2> gpurun_out/b15_bench_ref.err
RGBID_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gn_|pyr_down|warp_" -c 400 --csv --log-file gpurun_out/r02c_launches_warpfirst.csv python tools/profile_step.py 32 3 warpFirst > gpurun_out/b15_ncu.log 2>&1
for f in warpfirst align full ref; do cut -c1-400 gpurun_out/b15_bench_$f.json; tail -2 gpurun_out/b15_bench_$f.err; done
