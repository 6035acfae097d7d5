"""Micro-benchmark of the two Gauss-Newton kernels per level (CUDA events via the ABI's measurement hook).
Usage: python tools/bench_build.py [streams] [mode: tracker|align]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgbid_slam_b200 import capi, host, synth  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    mode = capi.MODE_ALIGN if (len(sys.argv) > 2 and sys.argv[2] == "align") else capi.MODE_TRACKER
    rows, cols, levels = 480, 640, 4
    p = synth.make_pair(seed=1, rows=rows, cols=cols, device="cuda", noise=True)
    ctx = host.Context(0)
    cfg = host.make_align_config(rows, cols, levels, mode, batch=S, **p["intr"])
    al = host.Aligner(ctx, cfg)
    WA, IA = ctx.convert_depth_to_invdepth(p["depth_a"]), ctx.compute_intensity(p["rgb_a"])
    for b in range(S):
        al.set_keyframe(b, WA, IA)
        al.set_current_rgbd(b, p["depth_b"], p["rgb_b"])
    al.run()
    peak = 6539.2
    out = []
    for lvl in range(3):
        ms = al.time_build(lvl, 20)
        gb = 32.0 * (rows >> lvl) * (cols >> lvl) * S / 1e9
        sc = al.time_build(lvl, 20, scale=True)
        out.append("L%d build %.1f us (%.0f GB/s, %.1f%% of %.0f) scale %.1f us" % (lvl, ms * 1e3, gb / (ms * 1e-3), 100 * gb / (ms * 1e-3) / peak, peak, sc * 1e3))
    print("S=%d | " % S + " | ".join(out))


if __name__ == "__main__":
    main()
