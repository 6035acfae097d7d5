"""Workload for ncu captures: a few tracker steps of the bench configuration (see /opt/skills/guides/B200_PROFILING.md).
Usage: python tools/profile_step.py [streams] [frames] [pyrFirst|warpFirst]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rgbid_slam_b200 import capi, host  # noqa: E402


class A:
    rows, cols, levels = 480, 640, 4


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    depth, rgb, intr = bench.make_frames(A, list(range(S)), n, "cuda")
    ctx = host.Context(0)
    its = host.default_iterations(A.levels, capi.MODE_TRACKER)
    warp_first = 1 if (len(sys.argv) > 3 and sys.argv[3] == "warpFirst") else 0
    acfg = host.make_align_config(A.rows, A.cols, A.levels, capi.MODE_TRACKER, batch=S, iterations=its, warp_first=warp_first,
                                  **intr)
    trk = host.Tracker(ctx, host.make_tracker_config(acfg))
    for k in range(n):
        trk.track(depth[k], rgb[k])
    torch.cuda.synchronize()
    print("launches", ctx.launches)


if __name__ == "__main__":
    main()
