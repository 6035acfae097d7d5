#!/bin/bash
# Writes a 640x480 synthetic TUM-format sequence and runs the RGBID_SLAMapp-compatible driver on it (single stream) with the
# [VISODO] section of the reference's shipped config_data/visodoRGBDconfig.ini (WARP_ORDER = pyrFirst).
set -e
N=${1:-40}
OUT=${2:-/tmp/rgbd_dataset_synth640}
python - <<PY
import sys; sys.path.insert(0, '.')
import rgbid_slam_b200
from rgbid_slam_b200 import synth
seq = synth.make_sequence(seed=20261018, n_frames=$N, rows=480, cols=640, noise=True)
synth.write_tum_sequence(seq, "$OUT")
i = seq["intr"]
open("$OUT/visodo.ini", "w").write("[VISODO]\nM_ESTIMATOR = Student\nSIGMA_ESTIMATOR = sigmaML\nWARP_ORDER = pyrFirst\nIMAGE_FILTERING = none\n")
open("$OUT/calibration.ini", "w").write("[CALIBRATION]\nfx=%r\nfy=%r\ncx=%r\ncy=%r\n" % (i["fx"], i["fy"], i["cx"], i["cy"]))
PY
./apps/rgbid_slam_app -eval $OUT/ -match_file matches.txt -config $OUT/visodo.ini -calib $OUT/calibration.ini -o $OUT/poses.txt | tail -2
head -3 $OUT/poses.txt
