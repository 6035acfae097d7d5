import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from util import pair_maps, cuda
import oracle as orc
from oracle import ref as refk
from rgbid_slam_b200 import host
ctx = host.Context(0)
P = pair_maps(seed=20261018, rows=480, cols=640, noise=True)
src = P["IA"]
mine = ctx.bilateral_filter(cuda(src), 3.0).cpu().numpy()
r = refk.bilateral(cuda(src), 3.0).cpu().numpy()
d = np.abs(mine - r)
ys, xs = np.where(d > 1e-4)
print("bilateral diff count", len(ys), "rows", np.unique(ys)[:10], "cols", np.unique(xs)[:10], "max", d.max())
