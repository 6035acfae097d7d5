"""Prints the per-segment clock counts of gn_scale_kernel's rounds.  Needs a library built with
RGBID_EXTRA_NVCC_FLAGS=-DRGBID_SCALE_PROBE=1 (the kernel then prints from CTA 0 / 131); one alignment of 32 streams."""
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
