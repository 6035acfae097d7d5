#!/bin/bash
mkdir -p gpurun_out
RGBID_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gn_|pyr_down|ingest|visibility|warp_|vmap|nmap|bilateral|gradient|copy|control_upload|fill_|export|keyframe_maps" -c 700 --csv --log-file gpurun_out/r02k_launches_all.csv python tools/profile_step.py 32 6 > gpurun_out/b51_ncu2.log 2>&1
tail -n 1 gpurun_out/b44_ncu1.log; tail -n 1 gpurun_out/b51_ncu2.log
