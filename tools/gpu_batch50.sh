#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/b50_pytest.txt 2>&1
timeout 900 python bench.py > gpurun_out/b50_bench_full.json 2> gpurun_out/b50_bench_full.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b50_smoke.txt 2>&1
RGBID_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_scale -s 5 -c 1 -o gpurun_out/r02j_scale -f python tools/profile_step.py 32 3 > gpurun_out/b50_ncu.log 2>&1
tail -n 3 gpurun_out/b50_pytest.txt; cut -c1-330 gpurun_out/b50_bench_full.json; tail -n 1 gpurun_out/b50_smoke.txt; tail -n 1 gpurun_out/b50_ncu.log
