#!/bin/bash
mkdir -p gpurun_out
# launch list of steady-state tracker steps (un-graphed so that ncu sees every kernel); only this library's kernels
RGBID_NO_GRAPH=1 timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gn_|pyr_down|ingest|visibility|warp_|vmap|nmap|bilateral|gradient|copy2|control_upload|fill_" -s 330 -c 160 --csv --log-file gpurun_out/b6_launches.csv python tools/profile_step.py 32 8 > gpurun_out/b6_ncu.log 2>&1
RGBID_NO_GRAPH=1 timeout 200 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/b6_bench_nograph.json 2>/dev/null
RGBID_NO_GRAPH=1 RGBID_NO_PDL=1 timeout 200 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/b6_bench_nograph_nopdl.json 2>/dev/null
cut -c1-330 gpurun_out/b6_bench_nograph.json; cut -c1-330 gpurun_out/b6_bench_nograph_nopdl.json; tail -3 gpurun_out/b6_ncu.log; wc -l gpurun_out/b6_launches.csv
