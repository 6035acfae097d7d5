#!/bin/bash
# round-2 GPU batch 1: calibration microbenchmark, tail probe, sigma/nu MLP variant, ragged-size aligner test
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/b1_smi.txt 2>&1
./tools/ubench/rank1 > gpurun_out/b1_rank1.txt 2>&1
python tools/bench_build.py 32 > gpurun_out/b1_bench_build_base.txt 2>&1
RGBID_LIB=$PWD/rgbid-slam_b200/lib/librgbid_b200_mlp.so python tools/bench_build.py 32 > gpurun_out/b1_bench_build_mlp.txt 2>&1
RGBID_LIB=$PWD/rgbid-slam_b200/lib/librgbid_b200_probe.so timeout 300 python tools/scale_round_probe.py > gpurun_out/b1_tail_probe.txt 2>&1
RGBID_PREPARED_TESTS=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/b1_pytest.txt 2>&1
RGBID_LIB=$PWD/rgbid-slam_b200/lib/librgbid_b200_mlp.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/b1_pytest_mlp.txt 2>&1
tail -3 gpurun_out/b1_pytest.txt gpurun_out/b1_pytest_mlp.txt
cat gpurun_out/b1_rank1.txt gpurun_out/b1_bench_build_base.txt gpurun_out/b1_bench_build_mlp.txt
