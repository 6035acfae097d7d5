"""Golden vectors of SURVEY section 8 (f3, f4) written by the reference's own kernels on a B200
(tests/golden/make_golden_f34.py -> tests/golden/ref_golden_f34.npz): the numpy restatement (CPU, every round) and the
CUDA path through the C ABI (GPU) must both reproduce them."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_f34 import inputs  # noqa: E402
from oracle import calib as oc  # noqa: E402
import test_calib_ops as T  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden_f34.npz")


def check_all(G, got):
    # float maps: same validity, a few ulp (FMA contraction); gathers may pick the neighbouring texel on a boundary
    assert T.agree(got["undist_I"], G["undist_I"], rel=1e-6, abs_tol=2e-2, frac=0.999)
    assert T.agree(got["undist_W"], G["undist_W"], rel=2e-6, frac=0.999)
    assert T.agree(got["register_W"], G["register_W"], rel=2e-6, frac=0.995)
    assert T.agree(got["fuse_depth"], G["fuse_depth"], rel=1e-6, frac=0.9999)
    assert T.agree(got["fuse_weight"], G["fuse_weight"], rel=1e-6, frac=0.9999)
    for k in ("fuse_colors", "image_grey", "image_rgb"):  # 8-bit results: at most one grey level on a rounding tie
        d = np.abs(got[k].astype(int) - G[k].astype(int))
        assert d.max() <= 1 and np.mean(d == 0) > 0.98, k


def test_numpy_oracle_reproduces_reference_kernels():
    G = dict(np.load(GOLD))
    x = inputs()
    dRc_proj, t_dc_proj, cRd_proj = T.projective(x["rgb_i"], x["dep_i"])
    got = {}
    got["undist_I"] = oc.undistort_intensity(x["I"], x["rgb_i"])
    got["undist_W"] = oc.undistort_depthinv(x["W"], x["dep_i"], T.DEPTH_DIST)
    got["register_W"] = oc.register_depthinv(G["undist_W"], dRc_proj, t_dc_proj, cRd_proj)
    d, c, w = x["depth_dst"].copy(), x["colors_dst"].copy(), x["weight_dst"].copy()
    oc.integrate_warped_rgb(x["dw"], *x["chans"], x["ww"], d, c, w)
    got["fuse_depth"], got["fuse_colors"], got["fuse_weight"] = d, c, w
    got["image_grey"] = oc.generate_image(x["vm"], x["nm"], x["light"])
    got["image_rgb"] = oc.generate_image(x["vm"], x["nm"], x["light"], x["rgb"])
    check_all(G, got)


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_kernels(ctx):
    import torch
    G = dict(np.load(GOLD))
    x = inputs()
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    npy = lambda t: t.cpu().numpy()
    dRc_proj, t_dc_proj, cRd_proj = T.projective(x["rgb_i"], x["dep_i"])
    got = {}
    got["undist_I"] = npy(ctx.undistort_intensity(cu(x["I"]), x["rgb_i"]))
    got["undist_W"] = npy(ctx.undistort_depthinv(cu(x["W"]), x["dep_i"], T.DEPTH_DIST))
    got["register_W"] = npy(ctx.register_depthinv(cu(G["undist_W"]), dRc_proj, t_dc_proj, cRd_proj))
    d, c, w = cu(x["depth_dst"]), cu(x["colors_dst"]), cu(x["weight_dst"])
    ctx.integrate_warped_rgb(cu(x["dw"]), cu(x["chans"][0]), cu(x["chans"][1]), cu(x["chans"][2]), cu(x["ww"]), d, c, w)
    got["fuse_depth"], got["fuse_colors"], got["fuse_weight"] = npy(d), npy(c), npy(w)
    got["image_grey"] = npy(ctx.generate_image(cu(x["vm"]), cu(x["nm"]), x["light"]))
    got["image_rgb"] = npy(ctx.generate_image(cu(x["vm"]), cu(x["nm"]), x["light"], cu(x["rgb"])))
    check_all(G, got)
