"""Workload for the clock64 probes (diagnostic builds, never bench values): one alignment of 32 streams, un-graphed.

  RGBID_BUILD_TAG=probe RGBID_EXTRA_NVCC_FLAGS="-DRGBID_SCALE_PROBE=1 -DRGBID_TAIL_PROBE=1" python rgbid-slam_b200/build.py
  gpurun -- 'RGBID_LIB=$PWD/rgbid-slam_b200/lib/librgbid_b200_probe.so python tools/scale_round_probe.py'

RGBID_SCALE_PROBE: gn_scale_kernel prints the per-segment clock counts of its rounds from CTA 0 / 131
(profiles/r01m_scale_round_probe.txt); RGBID_TAIL_PROBE: gn_build_fast_kernel prints pixel loop / reduction / serial tail
of the last CTA of stream 0."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RGBID_NO_GRAPH"] = "1"
from rgbid_slam_b200 import capi, host, synth  # noqa: E402

S = 32
rows, cols, levels = 480, 640, 4
p = synth.make_pair(seed=1, rows=rows, cols=cols, device="cuda", noise=True)
ctx = host.Context(0)
cfg = host.make_align_config(rows, cols, levels, capi.MODE_TRACKER, batch=S, **p["intr"])
al = host.Aligner(ctx, cfg)
WA, IA = ctx.convert_depth_to_invdepth(p["depth_a"]), ctx.compute_intensity(p["rgb_a"])
for b in range(S):
    al.set_keyframe(b, WA, IA)
    al.set_current_rgbd(b, p["depth_b"], p["rgb_b"])
al.run()
torch.cuda.synchronize()
print("--- second run (warm)")
sys.stdout.flush()
al.run()
torch.cuda.synchronize()
