"""Where does the end-to-end step go?  Times the upload alone, tracking alone and both (tools/, diagnostic)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from rgbid_slam_b200 import capi, host

class A: rows, cols, levels = 480, 640, 4
S, n = 32, 12
depth, rgb, intr = bench.make_frames(A, list(range(S)), n, "cuda")
ctx = host.Context(0)
its = host.default_iterations(A.levels, capi.MODE_TRACKER)
acfg = host.make_align_config(A.rows, A.cols, A.levels, capi.MODE_TRACKER, batch=S, iterations=its, **intr)
trk = host.Tracker(ctx, host.make_tracker_config(acfg))
hd, hc = depth.cpu().pin_memory(), rgb.cpu().pin_memory()
for k in range(3):
    trk.track(hd[k], hc[k])
torch.cuda.synchronize()
def wall(f, reps=4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
k = [3]
def only_prefetch():
    trk.prefetch(hd[4], hc[4]); torch.cuda.synchronize()
def only_track_dev():
    trk.track(depth[k[0]], rgb[k[0]]); k[0] += 1
def track_host():
    trk.track(hd[k[0]], hc[k[0]]); k[0] += 1
def track_prefetched():
    trk.prefetch(hd[k[0] + 1], hc[k[0] + 1]); trk.track(hd[k[0]], hc[k[0]]); k[0] += 1
print("prefetch alone   %.3f ms" % wall(only_prefetch))
print("track device     %.3f ms" % wall(only_track_dev, 2)); 
print("track host       %.3f ms" % wall(track_host, 2))
trk.prefetch(hd[k[0]], hc[k[0]])
print("track prefetched %.3f ms" % wall(track_prefetched, 3))

# does an unrelated, concurrent host->device copy slow the tracker down?
side = torch.cuda.Stream()
dst_d, dst_c = torch.empty_like(depth[0]), torch.empty_like(rgb[0])
def track_dev_with_unrelated_copy():
    with torch.cuda.stream(side):
        dst_d.copy_(hd[5], non_blocking=True); dst_c.copy_(hc[5], non_blocking=True)
    trk.track(depth[k[0]], rgb[k[0]]); k[0] += 1
k[0] = 3
print("track device + unrelated H2D on another stream %.3f ms" % wall(track_dev_with_unrelated_copy, 3))
k[0] = 3
print("track device (again)   %.3f ms" % wall(only_track_dev, 3))
