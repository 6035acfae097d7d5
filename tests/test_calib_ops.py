"""SURVEY section 8 (f3, f4): custom-calibration ingest (undistortIntensity, undistortDepthInv, registerDepthinv), colour
fusion (integrateWarpedRGB) and shaded previews (generateImage / generateImageRGB) through the C ABI, against the numpy
restatement (oracle/calib.py) and against the reference's own kernels (oracle/_ref)."""
import numpy as np
import pytest
import torch

from oracle import calib as oc
from oracle import ref as refk
import oracle as orc
from rgbid_slam_b200 import synth

# config_data/calibration_custom.ini of the reference
RGB_INTR = dict(fx=538.60293, fy=534.30083, cx=313.02560, cy=230.23339, k1=-0.01630)
DEPTH_INTR = dict(fx=572.02794, fy=567.29006, cx=315.76083, cy=232.71632, k1=-0.02711)
DEPTH_DIST = dict(c1=0.98954, c0=-1.2618e-03,
                  q0=[7.0023e-03, 1.0844e-02, -6.0580e-01, 1.2602e+00, -2.3050e-03, 1.6084e-02, 2.1441e-02, -1.8073e-02, -3.6722e-02],
                  q1=[-6.7052e-03, -1.9692e-03, 5.5808e-01, -1.2327e+00, 1.2714e-02, -2.0804e-02, -7.5163e-03, 3.1985e-02, 4.8632e-02],
                  xshift=4, yshift=4)
DRC = np.array([[0.9999, 0.0143, 0.0060], [-0.0143, 0.9999, -0.0018], [-0.0060, 0.0017, 1.0000]], dtype=np.float32)
T_DC = np.array([0.0263595, -0.0000973, 0.0002853], dtype=np.float32)


def K(i):
    return np.array([[i["fx"], 0, i["cx"]], [0, i["fy"], i["cy"]], [0, 0, 1]], dtype=np.float32)


def projective(rgb_intr, depth_intr):
    """dRc_proj = Kd dRc Kc^-1, t_dc_proj = Kd t_dc, cRd_proj = dRc_proj^-1 (src/visodo.cpp:789-807, float)"""
    dRc_proj = (K(depth_intr) @ DRC @ np.linalg.inv(K(rgb_intr))).astype(np.float32)
    return dRc_proj, (K(depth_intr) @ T_DC).astype(np.float32), np.linalg.inv(dRc_proj).astype(np.float32)


def frame(rows=480, cols=640, seed=3):
    p = synth.make_pair(seed=seed, rows=rows, cols=cols, noise=True)
    W = orc.depth_to_invdepth(p["depth_a"].numpy().astype(np.uint16))
    I = orc.intensity(p["rgb_a"].numpy())
    return W, I, p


def agree(a, b, rel=1e-5, frac=0.999, abs_tol=0.0):
    """same validity almost everywhere; values within tolerance on the common support"""
    nan_same = np.mean(np.isnan(a) == np.isnan(b))
    m = ~(np.isnan(a) | np.isnan(b))
    ok = np.abs(a[m] - b[m]) <= rel * np.abs(b[m]) + abs_tol
    return nan_same >= frac and np.mean(ok) >= frac


def test_oracle_undistortion_is_identity_without_distortion():
    """k = 0 and a pass-through depth model: undistortion must reproduce its input (interior)"""
    W, I, _ = frame(120, 160)
    intr = dict(fx=131.0, fy=131.0, cx=79.5, cy=59.5)
    out = oc.undistort_intensity(I, intr)
    assert np.allclose(out[1:-1, 1:-1], I[1:-1, 1:-1], atol=2e-2)
    dp = dict(c1=1.0, c0=0.0, q0=[0.0] * 9, q1=[0.0] * 9, xshift=0, yshift=0)
    out = oc.undistort_depthinv(W, intr, dp)
    m = ~(np.isnan(out[2:-2, 2:-2]) | np.isnan(W[2:-2, 2:-2]))
    assert np.allclose(out[2:-2, 2:-2][m], W[2:-2, 2:-2][m], rtol=1e-6)
    H = np.eye(3, dtype=np.float32)
    reg = oc.register_depthinv(W, H, np.zeros(3, np.float32), H)
    m = ~(np.isnan(reg) | np.isnan(W))
    # identity registration still dilates every sample over its 2 x 2 footprint with a z-buffer (max inverse depth)
    from scipy.ndimage import maximum_filter
    wmax = maximum_filter(np.nan_to_num(W, nan=0.0), size=3)
    assert np.mean(m) > 0.8 and np.all(reg[m] >= W[m] * (1 - 1e-5)) and np.all(reg[m] <= wmax[m] * (1 + 1e-5))


@pytest.mark.gpu
def test_undistort_and_register_vs_oracle_and_reference(ctx):
    W, I, _ = frame()
    Wg, Ig = torch.from_numpy(W).cuda(), torch.from_numpy(I).cuda()
    got_I = ctx.undistort_intensity(Ig, RGB_INTR).cpu().numpy()
    assert agree(got_I, oc.undistort_intensity(I, RGB_INTR), rel=1e-5, abs_tol=2e-2)  # 1/256-px weight steps near edges
    got_W = ctx.undistort_depthinv(Wg, DEPTH_INTR, DEPTH_DIST)
    assert agree(got_W.cpu().numpy(), oc.undistort_depthinv(W, DEPTH_INTR, DEPTH_DIST), rel=2e-6)
    dRc_proj, t_dc_proj, cRd_proj = projective(RGB_INTR, DEPTH_INTR)
    got_R = ctx.register_depthinv(got_W, dRc_proj, t_dc_proj, cRd_proj).cpu().numpy()
    want_R = oc.register_depthinv(got_W.cpu().numpy(), dRc_proj, t_dc_proj, cRd_proj)
    assert np.mean(~np.isnan(got_R)) > 0.7 and agree(got_R, want_R, rel=2e-6, frac=0.998)
    if refk.available():
        assert agree(got_I, refk.undistort_intensity(Ig, RGB_INTR).cpu().numpy(), rel=1e-6, abs_tol=1e-3, frac=0.9995)
        ref_W = refk.undistort_depthinv(Wg, DEPTH_INTR, DEPTH_DIST)
        assert agree(got_W.cpu().numpy(), ref_W.cpu().numpy(), rel=1e-6, frac=0.9995)
        ref_R = refk.register_depthinv(ref_W, dRc_proj, t_dc_proj, cRd_proj).cpu().numpy()
        assert agree(got_R, ref_R, rel=1e-6, frac=0.999)


@pytest.mark.gpu
def test_colour_fusion_and_previews_vs_oracle_and_reference(ctx):
    W, I, p = frame(240, 320, seed=9)
    rng = np.random.default_rng(5)
    rows, cols = W.shape
    rgb = p["rgb_a"].numpy()
    dw = (W + rng.normal(0, 0.002, W.shape)).astype(np.float32)     # some pixels inside, some outside the 0.0075 gate
    dw[rng.random(W.shape) < 0.05] = np.nan
    chans = [rgb[:, :, c].astype(np.float32) + rng.normal(0, 2, W.shape).astype(np.float32) for c in range(3)]
    chans[1][rng.random(W.shape) < 0.02] = np.nan
    ww = rng.uniform(0.5, 2.0, W.shape).astype(np.float32)
    depth_dst = W.copy(); depth_dst[rng.random(W.shape) < 0.1] = np.nan
    colors_dst = np.ascontiguousarray(rgb[:, ::-1, :])               # something different from the source colours
    weight_dst = rng.uniform(1.0, 5.0, W.shape).astype(np.float32)

    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    g_d, g_c, g_w = cu(depth_dst), cu(colors_dst), cu(weight_dst)
    ctx.integrate_warped_rgb(cu(dw), cu(chans[0]), cu(chans[1]), cu(chans[2]), cu(ww), g_d, g_c, g_w)
    o_d, o_c, o_w = depth_dst.copy(), colors_dst.copy(), weight_dst.copy()
    oc.integrate_warped_rgb(dw, chans[0], chans[1], chans[2], ww, o_d, o_c, o_w)
    assert agree(g_d.cpu().numpy(), o_d, rel=1e-6, frac=0.9999) and agree(g_w.cpu().numpy(), o_w, rel=1e-6, frac=0.9999)
    assert np.mean(np.abs(g_c.cpu().numpy().astype(int) - o_c.astype(int)) <= 1) > 0.9999  # .5 rounding of approx division
    assert np.mean(g_c.cpu().numpy() == o_c) > 0.99

    vm = orc.vmap(W, 262.5, 262.5, 159.5, 119.5)
    gx, gy = orc.gradient(W)
    nm = orc.nmap_gradients(W, gx, gy, 262.5, 262.5, 159.5, 119.5)
    light = np.array([0.2, -0.1, -0.5], dtype=np.float32)
    for with_rgb in (False, True):
        got = ctx.generate_image(cu(vm), cu(nm), light, cu(rgb) if with_rgb else None).cpu().numpy()
        want = oc.generate_image(vm, nm, light, rgb if with_rgb else None)
        assert np.mean(np.abs(got.astype(int) - want.astype(int)) <= 1) > 0.9999 and np.mean(got == want) > 0.98
        if refk.available():
            ref = refk.generate_image(cu(vm), cu(nm), light, cu(rgb) if with_rgb else None).cpu().numpy()
            assert np.mean(got == ref) > 0.9999
    if refk.available():
        r_d, r_c, r_w = cu(depth_dst), cu(colors_dst), cu(weight_dst)
        refk.integrate_warped_rgb(cu(dw), cu(chans[0]), cu(chans[1]), cu(chans[2]), cu(ww), r_d, r_c, r_w)
        assert np.array_equal(g_d.cpu().numpy(), r_d.cpu().numpy(), equal_nan=True)
        assert np.array_equal(g_c.cpu().numpy(), r_c.cpu().numpy()) and np.array_equal(g_w.cpu().numpy(), r_w.cpu().numpy())


@pytest.mark.gpu
def test_tracker_with_custom_calibration(ctx):
    """prepareImagesCustomCalibration inside the tracker (rgbid_tracker_set_custom_calibration): the current-frame
    maps must be exactly what the three bridge ops produce one after the other, and tracking must run on them."""
    import ctypes as C
    from rgbid_slam_b200 import capi, host
    rows, cols, n = 240, 320, 4
    seq = synth.make_sequence(seed=21, n_frames=n, rows=rows, cols=cols, noise=True)
    s = 0.5  # the calibration file is for 640 x 480
    rgb_i = {k: (v * s if k in ("fx", "fy", "cx", "cy") else v) for k, v in RGB_INTR.items()}
    dep_i = {k: (v * s if k in ("fx", "fy", "cx", "cy") else v) for k, v in DEPTH_INTR.items()}
    acfg = host.make_align_config(rows, cols, 3, capi.MODE_TRACKER, batch=2, fx=rgb_i["fx"], fy=rgb_i["fy"], cx=rgb_i["cx"],
                                  cy=rgb_i["cy"])
    trk = host.Tracker(ctx, host.make_tracker_config(acfg))
    trk.set_custom_calibration(rgb_i, dep_i, DEPTH_DIST, DRC, T_DC)
    for k in range(n):
        d, c = seq["depth"][k].cuda(), seq["rgb"][k].cuda()
        res = trk.track(torch.stack([d, d]).contiguous(), torch.stack([c, c]).contiguous())
        assert res[0].status == 0 and res[1].status == 0
    # maps of the last frame, stream 1
    p, pitch = C.c_void_p(), C.c_size_t()
    got = {}
    for name, which in (("W", 6), ("I", 7)):  # MAP_W_CUR, MAP_I_CUR (csrc/aligner.hpp)
        capi.check(trk.lib.rgbid_aligner_map(trk.aligner_handle, which, 0, 1, C.byref(p), C.byref(pitch)), "aligner_map")
        got[name] = host._wrap_device(p.value, rows, cols, pitch.value, ctx.device)
    W = ctx.convert_depth_to_invdepth(seq["depth"][n - 1].cuda())
    I = ctx.compute_intensity(seq["rgb"][n - 1].cuda())
    K = lambda i: np.array([[i["fx"], 0, i["cx"]], [0, i["fy"], i["cy"]], [0, 0, 1]], dtype=np.float32)
    dRc_proj = (K(dep_i) @ DRC @ np.linalg.inv(K(rgb_i))).astype(np.float32)
    want_I = ctx.undistort_intensity(I, rgb_i)
    want_W = ctx.register_depthinv(ctx.undistort_depthinv(W, dep_i, DEPTH_DIST), dRc_proj, (K(dep_i) @ T_DC).astype(np.float32),
                                   np.linalg.inv(dRc_proj).astype(np.float32))
    assert torch.equal(torch.nan_to_num(got["I"], nan=-1.0), torch.nan_to_num(want_I, nan=-1.0))
    assert agree(got["W"].cpu().numpy(), want_W.cpu().numpy(), rel=1e-5, frac=0.999)  # K^-1 products differ in the last bit
    assert float(torch.isnan(got["W"]).float().mean()) < 0.3
    trk.set_custom_calibration(None, None, None, None, None)
    trk.close()
