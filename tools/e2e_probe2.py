"""Per-step wall time of the bench's two arms (diagnostic)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from rgbid_slam_b200 import capi, host

class A: rows, cols, levels = 480, 640, 4
S, W, K = 32, 3, 12
n = 1 + W + K
depth, rgb, intr = bench.make_frames(A, list(range(S)), n, "cuda")
ctx = host.Context(0)
its = host.default_iterations(A.levels, capi.MODE_TRACKER)
acfg = host.make_align_config(A.rows, A.cols, A.levels, capi.MODE_TRACKER, batch=S, iterations=its, **intr)
trk = host.Tracker(ctx, host.make_tracker_config(acfg))
hd, hc = depth.cpu().pin_memory(), rgb.cpu().pin_memory()

def run(fd, fc, host_path, prefetch):
    trk.reset()
    for k in range(1 + W):
        trk.track(fd[k], fc[k])
    torch.cuda.synchronize()
    out = []
    for k in range(1 + W, n):
        t0 = time.perf_counter()
        if prefetch and k + 1 < n:
            trk.prefetch(fd[k + 1], fc[k + 1])
        r = trk.track(fd[k], fc[k])
        out.append((time.perf_counter() - t0) * 1e3)
    return out

for name, args in (("device", (depth, rgb, False, False)), ("host inline", (hd, hc, True, False)), ("host prefetch", (hd, hc, True, True))):
    t = run(*args)
    print("%-14s mean %.3f ms | %s" % (name, sum(t) / len(t), " ".join("%.2f" % v for v in t)))
