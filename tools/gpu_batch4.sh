#!/bin/bash
mkdir -p gpurun_out
RGBID_LIB=$PWD/rgbid-slam_b200/lib/librgbid_b200_probe.so timeout 300 python tools/scale_round_probe.py 2>&1 | tail -22 > gpurun_out/b4_tail_probe.txt
timeout 1800 python -m pytest tests -m gpu -q -s --durations=15 > gpurun_out/b4_pytest.txt 2>&1
tail -n 60 gpurun_out/b4_pytest.txt; tail -12 gpurun_out/b4_tail_probe.txt
