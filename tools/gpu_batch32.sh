#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/b32_pytest.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b32_bench.json 2> gpurun_out/b32_bench.err
RGBID_NO_SPLIT_COV=1 timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b32_bench_nosplit.json 2> gpurun_out/b32_bench_nosplit.err
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/b32_bench2.json 2> gpurun_out/b32_bench2.err
tail -3 gpurun_out/b32_pytest.txt; for f in bench bench_nosplit bench2; do python -c "
import json;d=json.load(open('gpurun_out/b32_$f.json'));print('$f',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['wall_ms_per_step'],4))"; done
