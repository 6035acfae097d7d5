#!/bin/bash
mkdir -p gpurun_out
RGBID_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gn_|pyr_down|ingest|visibility|warp_|vmap|nmap|bilateral|gradient|copy2|control_upload|fill_|export" -c 700 --csv --log-file gpurun_out/r02b_launches_all.csv python tools/profile_step.py 32 6 > gpurun_out/b14_ncu.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b14_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b14_bench.json 2> gpurun_out/b14_bench.err
tail -2 gpurun_out/b14_ncu.log; wc -l gpurun_out/r02b_launches_all.csv; tail -n 3 gpurun_out/b14_pytest.txt; cut -c1-330 gpurun_out/b14_bench.json
