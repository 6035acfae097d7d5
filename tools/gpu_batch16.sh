#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rgbid-slam_b200/lib
for i in 1 2; do python tools/bench_build.py 32 >> gpurun_out/b16_build.txt 2>&1; RGBID_LIB=$L/librgbid_b200_noef.so python tools/bench_build.py 32 >> gpurun_out/b16_build_noef.txt 2>&1; done
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b16_bench.json 2> gpurun_out/b16_bench.err
RGBID_LIB=$L/librgbid_b200_noef.so timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b16_bench_noef.json 2> gpurun_out/b16_bench_noef.err
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b16_bench2.json 2> gpurun_out/b16_bench2.err
RGBID_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"gn_build_fast|gn_scale" -s 40 -c 40 --csv --log-file gpurun_out/r02d_launches_gn_evictfirst.csv python tools/profile_step.py 32 3 > gpurun_out/b16_ncu.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b16_pytest.txt 2>&1
cat gpurun_out/b16_build.txt gpurun_out/b16_build_noef.txt; for f in bench bench_noef bench2; do cut -c1-300 gpurun_out/b16_$f.json; done; tail -3 gpurun_out/b16_pytest.txt
