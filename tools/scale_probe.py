"""IRLS / bisection round counts per Gauss-Newton iteration (diagnostic for the scale kernel)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgbid_slam_b200 import capi, host, synth
rows, cols, levels = 480, 640, 4
p = synth.make_pair(seed=1, rows=rows, cols=cols, device="cuda", noise=True)
ctx = host.Context(0)
cfg = host.make_align_config(rows, cols, levels, capi.MODE_TRACKER, batch=1, **p["intr"])
al = host.Aligner(ctx, cfg)
WA, IA = ctx.convert_depth_to_invdepth(p["depth_a"]), ctx.compute_intensity(p["rgb_a"])
al.set_keyframe(0, WA, IA)
al.set_current_rgbd(0, p["depth_b"], p["rgb_b"])
out = al.run(want_trace=True)
for t in out["trace"][0]:
    print(t["level"], t["iter"], "irls", t["irls_iters_int"], t["irls_iters_depthinv"], "nu", t["nu_int"], t["nu_depthinv"], "sigma", round(t["sigma_int"], 3), round(t["sigma_depthinv"], 5))
