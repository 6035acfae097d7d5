"""Shared comparison of an implementation ("backend") against tests/golden/ref_golden_r1.npz -- outputs of the
REFERENCE'S OWN CUDA kernels on a B200 (see tests/golden/make_golden.py).  Used with the CPU oracle as backend
(test_golden_cpu.py: this is what pins the oracle) and with the new CUDA path (test_golden_gpu.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden_r1.npz")
ROWS, COLS, LEVELS = 96, 128, 3


def load():
    return np.load(GOLDEN)


def close_map(a, b, rtol, atol=0.0, frac=1.0, name=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    nan_agree = np.mean(np.isnan(a) == np.isnan(b))
    assert nan_agree >= frac, "%s: validity agreement %.6f < %.6f" % (name, nan_agree, frac)
    m = ~(np.isnan(a) | np.isnan(b))
    ok = np.abs(a[m] - b[m]) <= atol + rtol * np.abs(b[m])
    assert np.mean(ok) >= frac, "%s: %.6f of pixels within tolerance (need %.6f), max diff %.3e" % (
        name, np.mean(ok), frac, np.max(np.abs(a[m] - b[m])))


def sums_close(a, b, tol, name="", ignore_b=False):
    from util import sums_rel_err
    e = sums_rel_err(a, b, ignore_b=ignore_b)
    assert e < tol, "%s: 27-sum relative error %.3e >= %.1e" % (name, e, tol)


def rot_angle(Ra, Rb):
    dR = np.asarray(Ra).reshape(3, 3) @ np.asarray(Rb).reshape(3, 3).T
    return float(np.arccos(np.clip((np.trace(dR) - 1.0) / 2.0, -1.0, 1.0)))


def check_image_ops(B, G):
    """B: backend with numpy-in / numpy-out functions named like the oracle's."""
    dA, cA = G["depth"][0], G["rgb"][0]
    WA, IA = B.depth_to_invdepth(dA, 1.0), B.intensity(cA)
    close_map(WA, G["WA"], 3e-7, name="invdepth")
    close_map(B.depth_to_invdepth(dA, 5.0), G["invdepth_factor5"], 3e-7, name="invdepth/5")
    close_map(IA, G["IA"], 0, atol=2e-5, name="intensity")
    WA, IA = G["WA"], G["IA"]  # continue from the reference's own maps so that errors do not compound
    close_map(B.pyr_down(WA), G["pyr1_W"], 2e-6, name="pyr1_W")
    close_map(B.pyr_down(IA), G["pyr1_I"], 2e-6, name="pyr1_I")
    close_map(B.pyr_down(G["pyr1_W"]), G["pyr2_W"], 2e-6, name="pyr2_W")
    gx, gy = B.gradient(IA)
    assert np.array_equal(gx, G["gradI_x"], equal_nan=True) and np.array_equal(gy, G["gradI_y"], equal_nan=True)
    gx, gy = B.gradient(WA)
    assert np.array_equal(gx, G["gradW_x"], equal_nan=True) and np.array_equal(gy, G["gradW_y"], equal_nan=True)
    close_map(B.bilateral(WA, 2 * 0.0025), G["bilateral_W"], 5e-6, name="bilateral_W")
    close_map(B.bilateral(IA, 3.0), G["bilateral_I"], 5e-6, name="bilateral_I")
    i = G["intr"]
    v = B.vmap(WA, *i)
    close_map(v[:ROWS], G["vmap"][:ROWS], 2e-6, name="vmap.x")
    m = ~np.isnan(G["vmap"][:ROWS])
    for k in (1, 2):
        assert np.allclose(v[k * ROWS:(k + 1) * ROWS][m], G["vmap"][k * ROWS:(k + 1) * ROWS][m], rtol=2e-6)
    n = B.nmap_gradients(WA, G["gradW_x"], G["gradW_y"], *i)
    close_map(n[:ROWS], G["nmap"][:ROWS], 1e-5, atol=1e-6, frac=0.9995, name="nmap.x")


def check_warps(B, G):
    WA, IA = G["WA"], G["IA"]
    WB, IB = B.depth_to_invdepth(G["depth"][2], 1.0), B.intensity(G["rgb"][2])
    Rp, tp, Rf, tf = G["Rp"], G["tp"], G["Rf"], G["tf"]
    close_map(B.warp_invdepth(WB, WA, Rp, tp), G["warp_W"], 1e-5, frac=0.999, name="warp_W")
    close_map(B.warp_intensity(IB, WA, Rp, tp), G["warp_I_kfgeom"], 0, atol=2e-3, frac=0.995, name="warp_I(kf geometry)")
    close_map(B.warp_intensity(IB, G["warp_W"], Rp, tp), G["warp_I"], 0, atol=2e-3, frac=0.995, name="warp_I(warped geometry)")
    w = np.full((ROWS, COLS), 0.5, dtype=np.float32)
    Ww = B.warp_invdepth_weighted(WB, WA, w, Rp, tp)
    close_map(Ww, G["warp_weighted_W"], 1e-5, frac=0.999, name="warp_weighted_W")
    close_map(w, G["warp_weighted_weight"], 1e-4, frac=0.999, name="warp_weight")
    kf, kfw = G["fusion_kf_in"].copy(), np.ones((ROWS, COLS), dtype=np.float32)
    kf, kfw = B.integrate_warped_frame(G["warp_weighted_W"], G["warp_weighted_weight"], kf, kfw)
    close_map(kf, G["fusion_kf_out"], 1e-6, name="fusion_kf")
    close_map(kfw, G["fusion_weight_out"], 1e-6, name="fusion_weight")
    r, mask = B.visibility_ratio(WB, WA, Rf, tf, True)
    assert abs(r - float(G["visibility_ratio"])) < 2e-4
    ref_mask = np.unpackbits(G["overlap_mask"])[:ROWS * COLS].reshape(ROWS, COLS)
    assert np.mean(mask == ref_mask) > 0.9995
    assert abs(B.visibility_ratio(WA, WB, Rp, tp, False) - float(G["visibility_ratio_inv"])) < 2e-4


def check_scale(B, G):
    eI, eW = G["err_I"], G["err_W"]
    assert np.array_equal(B.compute_error(G["warp_I"], G["IA"], 3000), eI, equal_nan=True)
    assert np.array_equal(B.compute_error(G["warp_W"], G["WA"], 3000), eW, equal_nan=True)
    for e, key, s0 in ((eI, "sigma_nu_I", 5.0), (eW, "sigma_nu_W", 0.0025)):
        b, s, nu = B.sigma_nu_student(e, 0.0, s0)[:3]
        gb, gs, gnu = G[key]
        assert abs(s - gs) / gs < 1e-4 and abs(b - gb) < 1e-4 * gs and nu == gnu, (key, b, s, nu, G[key])
    assert B.nu_student(eI, 0.0, 5.0) == float(G["nu_only_I"]) and B.nu_student(eW, 0.0, 0.0025) == float(G["nu_only_W"])
    b, s, nu = B.sigma_nu_student(G["heavy_err"], 0.0, 5.0)[:3]
    assert nu == G["heavy_sigma_nu"][2] and abs(s - G["heavy_sigma_nu"][1]) / s < 1e-4
    for name, m in (("lsq", 0), ("huber", 1), ("tukey", 2), ("student", 3)):
        bb, ss = B.sigma_pdf(eI, 0.0, 5.0, m)
        assert abs(ss - G["sigma_pdf_I_" + name][1]) / ss < 1e-4, name
        c, t, n = B.chi_square(eI, eW, 5.0, 0.0025, m)
        gc, gt, gn = G["chi_" + name]
        assert n == gn and abs(c - gc) / gc < 1e-4 and abs(t - gt) < 1e-4, name


def check_systems(B, G):
    i = G["intr"]
    for k, (student_nu, mest, weighting) in enumerate(G["system_cfgs"]):
        s = B.build_system(G["WA"], G["IA"], G["gradW_x"], G["gradW_y"], G["gradI_x"], G["gradI_y"], G["warp_W"],
                           G["warp_I"], dict(fx=i[0], fy=i[1], cx=i[2], cy=i[3], mestimator=int(mest),
                                             weighting=int(weighting), student_nu=int(student_nu), sigma_depthinv=0.0012,
                                             sigma_int=3.5, bias_depthinv=1e-5, bias_int=0.2, nu_depthinv=4.25, nu_int=6.5))
        sums_close(s, G["system_sums"][k], 1e-5, name="system cfg %d" % k)  # north-star bar


def check_align(B, G):
    i = G["intr"]
    for mode, name, its in ((1, "align", [5, 5, 3]), (0, "tracker", [10, 5, 3])):
        out = B.align(G["WA"], G["IA"], G["depth"][2], G["rgb"][2], mode, its, i, 3000)
        sums_close(out["sums27"][0], G[name + "_trace_sums27"][0], 1e-5, name=name + " first iteration")
        assert np.array_equal(out["scale"][0][4:], G[name + "_trace_scale"][0][4:]), (out["scale"][0], G[name + "_trace_scale"][0])
        dt = np.linalg.norm(out["t"] - G[name + "_t"])
        ang = rot_angle(out["R"], G[name + "_R"])
        assert dt < 1e-4 and ang < 1e-4, (name, dt, ang)   # north-star bar
        for k in range(len(its)):  # per-iteration poses stay together
            assert np.linalg.norm(out["trace_t"][k] - G[name + "_trace_t"][k]) < 1e-4
        cov_rel = np.abs(out["cov"] - G[name + "_cov"]).max() / np.abs(G[name + "_cov"]).max()
        assert cov_rel < 1e-3, (name, cov_rel)
        if mode == 0:
            sums_close(out["cov_sums27"], G["tracker_cov_sums27"], 1e-4, name="covariance pass", ignore_b=True)
            assert abs(out["chi"][0] - G["tracker_chi"][0]) / G["tracker_chi"][0] < 1e-3 and out["chi"][2] == G["tracker_chi"][2]


def check_sequence(B, G):
    poses, flags, fused = B.track_sequence(G["depth"], G["rgb"], G["intr"], 3000)
    assert np.array_equal(np.asarray(flags), G["seq_flags"])
    for k in range(len(poses)):
        assert np.linalg.norm(poses[k][9:] - G["seq_poses"][k][9:]) < 1e-4
        assert rot_angle(poses[k][:9], G["seq_poses"][k][:9]) < 1e-4
    close_map(fused, G["seq_fused_kf"], 1e-4, frac=0.999, name="fused keyframe")
