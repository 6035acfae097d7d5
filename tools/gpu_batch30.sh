#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/b30_pytest.txt 2>&1
timeout 900 python bench.py > gpurun_out/b30_bench_full.json 2> gpurun_out/b30_bench_full.err
timeout 600 python bench.py --no-cpu-baseline --no-extras --warp-order warpFirst > gpurun_out/b30_bench_warpfirst.json 2> gpurun_out/b30_bench_warpfirst.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b30_smoke.txt 2>&1
RGBID_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gn_|pyr_down|ingest|visibility|warp_|vmap|nmap|bilateral|gradient|copy|control_upload|fill_|export|keyframe_maps" -c 700 --csv --log-file gpurun_out/r02g_launches_all.csv python tools/profile_step.py 32 6 > gpurun_out/b30_ncu.log 2>&1
tail -3 gpurun_out/b30_pytest.txt; cut -c1-330 gpurun_out/b30_bench_full.json; cut -c1-250 gpurun_out/b30_bench_warpfirst.json; tail -1 gpurun_out/b30_smoke.txt
