"""Checks of the CPU oracle itself: un-vendored third-party arithmetic (Eigen LLT / inverse / JacobiSVD polar
factor, boost::math::digamma) against numpy / scipy, algebraic properties, and the restated algorithm on small
hand-checkable cases.  The golden vectors produced by the reference's own kernels are checked in
test_golden_cpu.py."""
import numpy as np
import pytest
import scipy.special

import oracle as orc
from util import pair_maps, rot_angle


def test_digamma_matches_scipy():
    for x in np.concatenate([np.arange(1.0, 5.6, 0.125), [0.5, 10.0, 37.5]]):
        assert abs(orc.digamma(x) - scipy.special.digamma(x)) < 1e-13


def test_llt_and_inverse_match_numpy():
    rng = np.random.default_rng(0)
    for _ in range(20):
        J = rng.normal(size=(40, 6))
        A = J.T @ J + 1e-3 * np.eye(6)
        b = rng.normal(size=6)
        assert np.allclose(orc.llt_solve6(A, b), np.linalg.solve(A, b), rtol=1e-10, atol=1e-12)
        assert np.allclose(orc.inverse6(A), np.linalg.inv(A), rtol=1e-9, atol=1e-12)
    assert np.isnan(orc.llt_solve6(-np.eye(6), np.ones(6))).all()  # not SPD -> NaN, the reference's "lost" trigger


def test_force_orthogonal_is_svd_polar_factor():
    rng = np.random.default_rng(1)
    for _ in range(20):
        M = np.eye(3) + 0.2 * rng.normal(size=(3, 3))
        U, _, Vt = np.linalg.svd(M)  # forceOrthogonalisation: U V^T (src/util_funcs.cpp:150-155)
        assert np.allclose(orc.force_orthogonal(M), U @ Vt, atol=1e-12)


def test_exp_log_maps():
    rng = np.random.default_rng(2)
    for scale in (1e-7, 1e-3, 0.3, 2.0):
        w, v = rng.normal(size=3) * scale, rng.normal(size=3)
        R, t = orc.exp_map(w, v)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12
        assert np.allclose(orc.exp_map_rot(w), R, atol=1e-14)
        tw = orc.log_map(R, t)
        assert np.allclose(tw[3:], w, atol=1e-9) and np.allclose(tw[:3], v, atol=1e-8)
    assert np.allclose(orc.exp_map_rot(np.zeros(3)), np.eye(3))


def test_gn_update_premultiplies_inverse_increment():
    """x = [v; w]: R_inc = exp(w)^-1, t_inc = -R_inc v, T <- T_inc T (src/visodo.cpp:1252-1263)."""
    A, x = np.eye(6), np.array([0.01, -0.02, 0.03, 0.001, 0.002, -0.003])
    R0, t0 = orc.exp_map(np.array([0.1, 0.2, -0.1]), np.array([0.3, -0.1, 0.2]))
    R, t, xs, bad = orc.gn_update(A, x, R0, t0)
    Rinc = orc.exp_map_rot(x[3:]).T
    assert bad == 0 and np.allclose(xs, x)
    assert np.allclose(R, Rinc @ R0, atol=1e-12) and np.allclose(t, Rinc @ t0 - Rinc @ x[:3], atol=1e-12)


def test_pyr_down_rule_on_small_case():
    src = np.arange(64, dtype=np.float32).reshape(8, 8)
    out = orc.pyr_down(src)
    assert out.shape == (4, 4)
    # interior pixel: normalised Gaussian of a linear ramp reproduces the centre value
    assert abs(out[1, 1] - src[2, 2]) < 1e-4 and abs(out[2, 2] - src[4, 4]) < 1e-4
    # corner (0,0): window clipped to 3x3 = 9 taps <= 12 -> invalid (pyrdown.cu:124-127)
    assert np.isnan(out[0, 0]) and not np.isnan(out[1, 1])
    src2 = src.copy()
    src2[1:4, 1:4] = np.nan  # 9 of 25 invalid -> 16 valid > 12 -> still valid
    assert not np.isnan(orc.pyr_down(src2)[1, 1])
    src2[0, 0:4] = np.nan    # 13 invalid -> 12 valid -> invalid
    assert np.isnan(orc.pyr_down(src2)[1, 1])


def test_gradient_sobel_over_8():
    y, x = np.mgrid[0:6, 0:7].astype(np.float32)
    gx, gy = orc.gradient(3 * x + 2 * y)
    assert np.allclose(gx[1:-1, 1:-1], 3.0) and np.allclose(gy[1:-1, 1:-1], 2.0)
    assert np.allclose(gx[:, 0], 1.5)  # clamp-to-edge halves the border gradient
    img = (3 * x + 2 * y).copy()
    img[3, 3] = np.nan
    gx, _ = orc.gradient(img)
    assert np.isnan(gx[2:5, 2:5]).all() and not np.isnan(gx[0, 0])  # zero weights still propagate NaN


def test_depth_and_intensity_conversions():
    d = np.array([[0, 1, 500, 10000, 20000]], dtype=np.uint16)
    w = orc.depth_to_invdepth(d)
    assert np.isnan(w[0, 0]) and w[0, 1] == 1000.0 and w[0, 2] == 2.0 and w[0, 3] == w[0, 4] == np.float32(0.1)
    assert orc.depth_to_invdepth(d, 5.0)[0, 2] == np.float32(np.float32(1 / np.float32(5)) * 1000 / 500)
    rgb = np.array([[[255, 255, 255], [0, 0, 0], [10, 20, 30]]], dtype=np.uint8)
    i = orc.intensity(rgb)
    assert i[0, 0] <= 255.0 and i[0, 1] == 0 and abs(i[0, 2] - (2.126 + 14.304 + 2.166)) < 1e-4


def test_identity_warp_properties():
    P = pair_maps(seed=5, rows=120, cols=160)
    i = P["intr"]
    Rp, tp = orc.projective_pose(np.eye(3), np.zeros(3), i["fx"], i["fy"], i["cx"], i["cy"])
    W1 = orc.warp_invdepth(P["WA"], P["WA"], Rp, tp)
    m = ~np.isnan(P["WA"])
    assert np.array_equal(np.isnan(W1), ~m) and np.allclose(W1[m], P["WA"][m], rtol=1e-5)
    I1 = orc.warp_intensity(P["IA"], P["WA"], Rp, tp)
    assert np.allclose(I1[m], P["IA"][m], atol=2e-3)  # 1/256 weight quantisation at ~integer coordinates
    assert orc.visibility_ratio(P["WA"], P["WA"], Rp, tp) > 0.97


def test_texture_fraction_quantisation_modes():
    img = np.tile(np.arange(8, dtype=np.float32) * 256.0, (4, 1))
    W = np.full((4, 8), 0.5, dtype=np.float32)
    Rp, tp = np.eye(3, dtype=np.float32), np.array([0.3 / 0.5 / 1.0 * 0.5, 0, 0], dtype=np.float32)  # shift x by 0.3 px * w
    orc.set_tex_frac_mode(orc.TEX_FRAC_EXACT)
    a = orc.warp_intensity(img, W, Rp, tp)
    orc.set_tex_frac_mode(orc.TEX_FRAC_ROUND)
    b = orc.warp_intensity(img, W, Rp, tp)
    assert np.nanmax(np.abs(a - b)) <= 256.0 / 512 + 1e-3  # half a quantisation step of the weight
    assert np.all((b[~np.isnan(b)] * 1.0) % 1.0 == 0)     # weights are multiples of 1/256 -> integers here


def test_student_nu_estimate_tracks_tail_weight():
    rng = np.random.default_rng(3)
    nus = []
    for dof in (2.2, 4.0, 30.0):
        e = (rng.standard_t(dof, 19200) * 3.0).astype(np.float32)
        b, s, nu, iters = orc.sigma_nu_student(e, 0.0, 5.0)
        nus.append(nu)
        assert 2 <= iters <= 10 and 2.0 <= nu <= 10.0 and abs(b) < 0.2
        assert nu in (2.0, 10.0) or (nu * 4) % 1 == 0  # reachable values: 2, 10, x.25, x.75
    assert nus[0] < nus[1] <= nus[2] == 10.0


def test_sampled_error_layout():
    a = np.arange(480 * 640, dtype=np.float32).reshape(480, 640)
    e = orc.compute_error(a, np.zeros_like(a), 10000)
    assert e.size == 19200 and e[1] == a[0, 4] and e[160] == a[4, 0]  # index y*cols_kept + x, stride 4


@pytest.mark.parametrize("mode,levels,its", [(orc.MODE_ALIGN, 4, [5, 5, 3, 0]), (orc.MODE_TRACKER, 3, [10, 5, 3])])
def test_alignment_recovers_known_motion(mode, levels, its):
    P = pair_maps(seed=20261018, rows=240, cols=320)
    i = P["intr"]
    cfg = orc.make_config(240, 320, levels, mode, its, i["fx"], i["fy"], i["cx"], i["cy"])
    out = orc.align(cfg, orc.prepare_keyframe(P["WA"], P["IA"], levels, mode == orc.MODE_TRACKER),
                    orc.prepare_current(P["WB"], P["IB"], levels))
    assert out["status"] == 0 and len(out["trace"]) == sum(its)
    assert np.linalg.norm(out["t"] - P["t_ab"]) < 3e-4 and rot_angle(out["R"], P["R_ab"]) < 3e-4
    assert np.allclose(out["cov"], out["cov"].T, rtol=1e-6) and np.all(np.diag(out["cov"]) > 0)
    # 27-vector layout [A00..A05,b0,A11..] round-trips through the unpack (estimate_VO.cu:771-786)
    s = out["trace"][0]["sums27"]
    A, b = np.zeros(36), np.zeros(6)
    orc.lib().orc_unpack_system(s.ctypes.data_as(orc.C.c_void_p), A.ctypes.data_as(orc.C.c_void_p), b.ctypes.data_as(orc.C.c_void_p))
    A = A.reshape(6, 6)
    assert A[0, 5] == s[5] and b[0] == s[6] and A[1, 1] == s[7] and A[5, 5] == s[25] and b[5] == s[26] and np.array_equal(A, A.T)


def test_warp_first_order_of_the_tracker():
    """WARP_ORDER = warpFirst (src/visodo.cpp:1078-1105): warp at level 0, then the pyramid of the warped maps, every
    iteration.  Same as pyrFirst when only level 0 iterates; a different (but equally convergent) path otherwise."""
    P = pair_maps(seed=20261019, rows=240, cols=320, noise=True)
    i = P["intr"]
    kf, cur = orc.prepare_keyframe(P["WA"], P["IA"], 3, True), orc.prepare_current(P["WB"], P["IB"], 3)

    def run(its, warp_first):
        cfg = orc.make_config(240, 320, 3, orc.MODE_TRACKER, its, i["fx"], i["fy"], i["cx"], i["cy"], warp_first=warp_first)
        return orc.align(cfg, kf, cur)

    a, b = run([4, 0, 0], 0), run([4, 0, 0], 1)
    # (not bit-equal: the oracle's OpenMP reductions do not fix the summation order between runs)
    assert np.allclose(a["R"], b["R"], atol=1e-9) and np.allclose(a["t"], b["t"], atol=1e-9)
    assert np.allclose(a["cov"], b["cov"], rtol=1e-6)
    pf, wf = run([10, 5, 3], 0), run([10, 5, 3], 1)
    assert wf["status"] == 0 and len(wf["trace"]) == 18
    assert np.linalg.norm(wf["t"] - P["t_ab"]) < 5e-4 and rot_angle(wf["R"], P["R_ab"]) < 5e-4
    # the first iteration (level 2) already sees different warped maps: pyrDown(warp(.)) != warp(pyrDown(.))
    assert not np.array_equal(pf["trace"][0]["sums27"], wf["trace"][0]["sums27"])
    assert np.allclose(pf["trace"][0]["sums27"][[0, 7, 13]], wf["trace"][0]["sums27"][[0, 7, 13]], rtol=0.2)
    # KeyframeAlign has no such option (src/keyframe_align.cpp:178-350): the flag is ignored there
    ca = orc.make_config(240, 320, 3, orc.MODE_ALIGN, [3, 3, 3], i["fx"], i["fy"], i["cx"], i["cy"], warp_first=1)
    cb = orc.make_config(240, 320, 3, orc.MODE_ALIGN, [3, 3, 3], i["fx"], i["fy"], i["cx"], i["cy"], warp_first=0)
    kfa = orc.prepare_keyframe(P["WA"], P["IA"], 3, False)
    assert np.allclose(orc.align(ca, kfa, cur)["t"], orc.align(cb, kfa, cur)["t"], atol=1e-9)


def test_alignment_reports_lost_on_empty_input():
    nan = np.full((60, 80), np.nan, dtype=np.float32)
    cfg = orc.make_config(60, 80, 2, orc.MODE_TRACKER, [2, 2], 100, 100, 40, 30)
    out = orc.align(cfg, orc.prepare_keyframe(nan, nan, 2, True), orc.prepare_current(nan, nan, 2))
    assert out["status"] == 1 and np.array_equal(out["cov"], 100 * np.eye(6))
