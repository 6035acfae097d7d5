"""Golden vectors for SURVEY section 8 (f3, f4) from the reference's OWN kernels (oracle/_ref/libref_oracle.so: undistortion.cu,
warping_registration.cu, image_generator.cu compiled verbatim for sm_100a).  Run on a B200:
    gpurun -- 'python tests/golden/make_golden_f34.py gpurun_out/ref_golden_f34.npz'
and commit the file as tests/golden/ref_golden_f34.npz.  Inputs are regenerated from the seed by the tests."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc  # noqa: E402
from oracle import ref as refk  # noqa: E402
import test_calib_ops as T  # noqa: E402  (calibration constants, input builders)


def inputs(rows=120, cols=160):
    """Everything the golden file was computed from (seeded)."""
    W, I, p = T.frame(rows, cols, seed=17)
    s = cols / 640.0
    rgb_i = {k: (v * s if k in ("fx", "fy", "cx", "cy") else v) for k, v in T.RGB_INTR.items()}
    dep_i = {k: (v * s if k in ("fx", "fy", "cx", "cy") else v) for k, v in T.DEPTH_INTR.items()}
    rng = np.random.default_rng(23)
    rgb = p["rgb_a"].numpy()
    dw = (W + rng.normal(0, 0.002, W.shape)).astype(np.float32)
    dw[rng.random(W.shape) < 0.05] = np.nan
    chans = [rgb[:, :, c].astype(np.float32) + rng.normal(0, 2, W.shape).astype(np.float32) for c in range(3)]
    chans[1][rng.random(W.shape) < 0.02] = np.nan
    ww = rng.uniform(0.5, 2.0, W.shape).astype(np.float32)
    depth_dst = W.copy()
    depth_dst[rng.random(W.shape) < 0.1] = np.nan
    colors_dst = np.ascontiguousarray(rgb[:, ::-1, :])
    weight_dst = rng.uniform(1.0, 5.0, W.shape).astype(np.float32)
    fx, fy, cx, cy = 525.0 * s, 525.0 * s, 319.5 * s, 239.5 * s
    vm = orc.vmap(W, fx, fy, cx, cy)
    gx, gy = orc.gradient(W)
    nm = orc.nmap_gradients(W, gx, gy, fx, fy, cx, cy)
    return dict(W=W, I=I, rgb=rgb, rgb_i=rgb_i, dep_i=dep_i, dw=dw, chans=chans, ww=ww, depth_dst=depth_dst,
                colors_dst=colors_dst, weight_dst=weight_dst, vm=vm, nm=nm, light=np.array([0.2, -0.1, -0.5], np.float32))


def main():
    out = sys.argv[1]
    x = inputs()
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    npy = lambda t: t.cpu().numpy()
    dRc_proj, t_dc_proj, cRd_proj = T.projective(x["rgb_i"], x["dep_i"])
    G = {}
    G["undist_I"] = npy(refk.undistort_intensity(cu(x["I"]), x["rgb_i"]))
    und_W = refk.undistort_depthinv(cu(x["W"]), x["dep_i"], T.DEPTH_DIST)
    G["undist_W"] = npy(und_W)
    G["register_W"] = npy(refk.register_depthinv(und_W, dRc_proj, t_dc_proj, cRd_proj))
    d, c, w = cu(x["depth_dst"]), cu(x["colors_dst"]), cu(x["weight_dst"])
    refk.integrate_warped_rgb(cu(x["dw"]), cu(x["chans"][0]), cu(x["chans"][1]), cu(x["chans"][2]), cu(x["ww"]), d, c, w)
    G["fuse_depth"], G["fuse_colors"], G["fuse_weight"] = npy(d), npy(c), npy(w)
    G["image_grey"] = npy(refk.generate_image(cu(x["vm"]), cu(x["nm"]), x["light"]))
    G["image_rgb"] = npy(refk.generate_image(cu(x["vm"]), cu(x["nm"]), x["light"], cu(x["rgb"])))
    np.savez_compressed(out, **G)
    print("wrote", out, {k: v.shape for k, v in G.items()})


if __name__ == "__main__":
    main()
