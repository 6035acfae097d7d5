import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from util import pair_maps, cuda
import oracle as orc
from oracle import ref as refk
P = pair_maps(seed=20261018, rows=480, cols=640, noise=False)
i = P["intr"]
L = 2
Rp, tp = orc.projective_pose(np.eye(3), np.zeros(3), i["fx"]/4, i["fy"]/4, i["cx"]/4, i["cy"]/4, inverse=True)
kf = orc.prepare_keyframe(P["WA"], P["IA"], 3, True); cur = orc.prepare_current(P["WB"], P["IB"], 3)
kfr = refk.prepare_keyframe(cuda(P["WA"]), cuda(P["IA"]), 3, True); curr = refk.prepare_current(cuda(P["WB"]), cuda(P["IB"]), 3)
def cmp(name, a, b):
    a = a.cpu().numpy() if hasattr(a, 'cpu') else a
    m = ~(np.isnan(a) | np.isnan(b))
    d = np.abs(a[m]-b[m])
    print("%-10s nan-mismatch %d  max|d| %.3e mean|d| %.3e  n(d>1e-3)=%d" % (name, (np.isnan(a)!=np.isnan(b)).sum(), d.max(), d.mean(), (d>1e-3).sum()))
for l in range(3):
    cmp("Wkf%d"%l, kfr["W"][l], kf["W"][l]); cmp("Ikf%d"%l, kfr["I"][l], kf["I"][l]); cmp("Wc%d"%l, curr["W"][l], cur["W"][l]); cmp("Ic%d"%l, curr["I"][l], cur["I"][l])
W1 = orc.warp_invdepth(cur["W"][L], kf["W"][L], Rp, tp); I1 = orc.warp_intensity(cur["I"][L], W1, Rp, tp)
W1r = refk.warp_invdepth(curr["W"][L], kfr["W"][L], Rp, tp); I1r = refk.warp_intensity(curr["I"][L], W1r, Rp, tp)
cmp("W1", W1r, W1); cmp("I1", I1r, I1)
# same inputs to the ref intensity warp
I1r2 = refk.warp_intensity(cuda(cur["I"][L]), cuda(W1), Rp, tp)
cmp("I1(same in)", I1r2, I1)
d = np.abs(I1r2.cpu().numpy() - I1); d[np.isnan(d)] = 0
ys, xs = np.where(d > 1e-3)
print("bad px sample", list(zip(ys[:10], xs[:10])), "vals", [(float(I1r2[y,x]), float(I1[y,x])) for y,x in list(zip(ys[:5], xs[:5]))])
eI = orc.compute_error(I1, kf["I"][L], 10000); eIr = refk.compute_error(I1r, kfr["I"][L], 10000)
cmp("eI", eIr, eI)
print("cpu", orc.sigma_nu_student(eI, 0, 5.0), "ref-on-ref", refk.sigma_nu_student(eIr, 0, 5.0))
