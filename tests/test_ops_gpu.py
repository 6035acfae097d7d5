"""Parity of every drop-in entry point (one per reference bridge function) through the C ABI against
(a) the CPU oracle and (b) -- when oracle/_ref/libref_oracle.so is present -- the reference's own kernels
running on the same GPU.  Tolerances are stated per test."""
import numpy as np
import pytest
import torch

from util import pair_maps, cuda, same_nan, max_abs_diff, sums_rel_err
import oracle as orc
from oracle import ref as refk

pytestmark = pytest.mark.gpu

ROWS, COLS = 480, 640


@pytest.fixture(scope="module")
def P():
    return pair_maps(seed=20261018, rows=ROWS, cols=COLS, noise=True)


@pytest.fixture(scope="module")
def pose(P):
    i = P["intr"]
    Rp, tp = orc.projective_pose(P["R_ab"], P["t_ab"], i["fx"], i["fy"], i["cx"], i["cy"], inverse=True)
    Rf, tf = orc.projective_pose(P["R_ab"], P["t_ab"], i["fx"], i["fy"], i["cx"], i["cy"], inverse=False)
    return dict(Rp=Rp, tp=tp, Rf=Rf, tf=tf)


def have_ref():
    return refk.available()


def test_ingest_bit_exact_invdepth(ctx, P):
    # integer -> float with one division: oracle uses IEEE division, the kernel the approximate one (<= 2 ulp)
    out = ctx.convert_depth_to_invdepth(cuda(P["dA"])).cpu().numpy()
    want = orc.depth_to_invdepth(P["dA"])
    assert same_nan(out, want)
    m = ~np.isnan(want)
    assert np.max(np.abs(out[m] - want[m]) / want[m]) < 3e-7
    if have_ref():
        r = refk.convert_depth_to_invdepth(cuda(P["dA"])).cpu().numpy()
        assert same_nan(out, r) and np.max(np.abs(out[m] - r[m]) / r[m]) < 3e-7


def test_ingest_edge_values(ctx):
    d = np.zeros((8, 16), dtype=np.uint16)
    d[0, :6] = [0, 1, 9999, 10000, 10001, 65535]
    out = ctx.convert_depth_to_invdepth(cuda(d), 5.0).cpu().numpy()
    want = orc.depth_to_invdepth(d, 5.0)
    assert same_nan(out, want) and np.isnan(out[0, 0]) and np.isnan(out[1, 1])
    np.testing.assert_allclose(out[0, 1:6], want[0, 1:6], rtol=3e-7)
    assert out[0, 4] == out[0, 3] == out[0, 5]  # clamp at 10 m


def test_intensity(ctx, P):
    out = ctx.compute_intensity(cuda(P["cA"])).cpu().numpy()
    want = orc.intensity(P["cA"])
    assert max_abs_diff(out, want) <= 2e-5  # FMA contraction on the GPU vs separate mul/add on the CPU
    if have_ref():
        assert max_abs_diff(out, refk.compute_intensity(cuda(P["cA"])).cpu().numpy()) <= 2e-5  # FMA association


def test_intensity_ragged(ctx):
    rng = np.random.default_rng(3)
    c = rng.integers(0, 256, (7, 13, 3), dtype=np.uint8)  # cols not a multiple of 4 -> scalar path
    assert max_abs_diff(ctx.compute_intensity(cuda(c)).cpu().numpy(), orc.intensity(c)) <= 2e-5


def test_decompose_rgb(ctx, P):
    r, g, b = ctx.decompose_rgb(cuda(P["cA"]))
    assert np.array_equal(r.cpu().numpy(), P["cA"][..., 0].astype(np.float32))
    assert np.array_equal(b.cpu().numpy(), P["cA"][..., 2].astype(np.float32))


@pytest.mark.parametrize("which", ["W", "I"])
def test_pyr_down(ctx, P, which):
    src = P[which + "A"]
    lvl, want = cuda(src), src
    for _ in range(3):
        lvl, want = ctx.pyr_down(lvl), orc.pyr_down(want)
        out = lvl.cpu().numpy()
        assert same_nan(out, want)  # validity rule count > 12 is integer logic: exact
        m = ~np.isnan(want)
        assert np.max(np.abs(out[m] - want[m]) / np.maximum(np.abs(want[m]), 1e-6)) < 2e-6  # __expf vs expf
        lvl = cuda(want)  # keep levels aligned so errors do not compound
    if have_ref():
        a, b = ctx.pyr_down(cuda(src)).cpu().numpy(), refk.pyr_down(cuda(src)).cpu().numpy()
        assert same_nan(a, b) and max_abs_diff(a, b) <= 1e-6 * max(1.0, float(np.nanmax(np.abs(b))))


def test_pyr_down_odd(ctx):
    rng = np.random.default_rng(5)
    src = rng.normal(size=(10, 14)).astype(np.float32)
    src[2:4, 3:9] = np.nan
    out, want = ctx.pyr_down(cuda(src)).cpu().numpy(), orc.pyr_down(src)
    assert out.shape == (5, 7) and same_nan(out, want) and max_abs_diff(out, want) < 1e-5


@pytest.mark.parametrize("which", ["W", "I"])
def test_gradient_bit_exact(ctx, P, which):
    src = P[which + "A"]
    gx, gy = ctx.compute_gradient(cuda(src))
    wx, wy = orc.gradient(src)
    # integer Sobel weights: products are exact, so GPU FMA == CPU mul+add; same summation order
    assert np.array_equal(gx.cpu().numpy(), wx, equal_nan=True)
    assert np.array_equal(gy.cpu().numpy(), wy, equal_nan=True)
    if have_ref():
        rx, ry = refk.gradient(cuda(src))
        assert np.array_equal(gx.cpu().numpy(), rx.cpu().numpy(), equal_nan=True)
        assert np.array_equal(gy.cpu().numpy(), ry.cpu().numpy(), equal_nan=True)


@pytest.mark.parametrize("which,sigma", [("W", 2 * 0.0025), ("I", 3.0)])
def test_bilateral(ctx, P, which, sigma):
    src = P[which + "A"]
    out, want = ctx.bilateral_filter(cuda(src), sigma).cpu().numpy(), orc.bilateral(src, sigma)
    assert same_nan(out, want)
    m = ~np.isnan(want)
    assert np.max(np.abs(out[m] - want[m]) / np.maximum(np.abs(want[m]), 1e-6)) < 5e-6
    if have_ref():
        r = refk.bilateral(cuda(src), sigma).cpu().numpy()
        assert same_nan(out, r) and np.max(np.abs(out[m] - r[m]) / np.maximum(np.abs(r[m]), 1e-6)) < 5e-6


def test_warp_invdepth(ctx, P, pose):
    out = ctx.warp_invdepth(cuda(P["WB"]), cuda(P["WA"]), pose["Rp"], pose["tp"]).cpu().numpy()
    want = orc.warp_invdepth(P["WB"], P["WA"], pose["Rp"], pose["tp"])
    # nearest-neighbour gather: a pixel whose source coordinate sits within float rounding of a texel boundary
    # may pick the neighbouring texel; require >= 99.9 % identical validity and tight agreement elsewhere
    agree = np.mean(np.isnan(out) == np.isnan(want))
    assert agree > 0.999
    m = ~(np.isnan(out) | np.isnan(want))
    rel = np.abs(out[m] - want[m]) / want[m]
    assert np.mean(rel < 1e-5) > 0.999
    if have_ref():
        r = refk.warp_invdepth(cuda(P["WB"]), cuda(P["WA"]), pose["Rp"], pose["tp"]).cpu().numpy()
        assert np.mean(np.isnan(out) == np.isnan(r)) > 0.9999
        m = ~(np.isnan(out) | np.isnan(r))
        assert np.mean(np.abs(out[m] - r[m]) / r[m] < 1e-6) > 0.9999


def test_warp_intensity_matches_texture_unit(ctx, P, pose):
    """The software bilinear sampler (1/256 weight quantisation) against the oracle and the hardware texture
    unit used by the reference (src/cuda/warping_registration.cu:943)."""
    out = ctx.warp_intensity(cuda(P["IB"]), cuda(P["WA"]), pose["Rp"], pose["tp"]).cpu().numpy()
    want = orc.warp_intensity(P["IB"], P["WA"], pose["Rp"], pose["tp"])
    assert np.mean(np.isnan(out) == np.isnan(want)) > 0.9999
    m = ~(np.isnan(out) | np.isnan(want))
    d = np.abs(out[m] - want[m])
    # a coordinate within float rounding of a 1/256 quantisation step may round the other way: one weight
    # step is worth |gradient|/256 grey levels
    assert np.mean(d < 1e-3) > 0.99 and d.max() < 0.5
    if have_ref():
        r = refk.warp_intensity(cuda(P["IB"]), cuda(P["WA"]), pose["Rp"], pose["tp"]).cpu().numpy()
        assert np.mean(np.isnan(out) == np.isnan(r)) > 0.9999
        m = ~(np.isnan(out) | np.isnan(r))
        d = np.abs(out[m] - r[m])
        print("texture-unit parity: mean |d| = %.3e, max = %.3e, frac(<1e-3) = %.5f" % (d.mean(), d.max(), np.mean(d < 1e-3)))
        assert np.mean(d < 1e-3) > 0.99 and d.max() < 0.5


def test_warp_weighted_and_integrate(ctx, P, pose):
    rng = np.random.default_rng(11)
    w_init = rng.uniform(0.5, 2.0, (ROWS, COLS)).astype(np.float32)
    wg = cuda(w_init.copy())
    out = ctx.warp_invdepth_weighted(cuda(P["WB"]), cuda(P["WA"]), wg, pose["Rp"], pose["tp"]).cpu().numpy()
    w_cpu = w_init.copy()
    want = orc.warp_invdepth_weighted(P["WB"], P["WA"], w_cpu, pose["Rp"], pose["tp"])
    assert np.mean(np.isnan(out) == np.isnan(want)) > 0.999
    m = ~(np.isnan(out) | np.isnan(want))
    assert np.mean(np.abs(out[m] - want[m]) / want[m] < 1e-5) > 0.999
    wgc = wg.cpu().numpy()
    assert np.mean(np.abs(wgc - w_cpu) / w_cpu < 1e-4) > 0.999
    # fusion: same inputs on both sides -> same decisions
    kf, kfw = P["WA"].copy(), np.ones_like(P["WA"])
    kf[100:140, 200:260] = np.nan  # adopt branch
    kf_g, kfw_g = cuda(kf.copy()), cuda(kfw.copy())
    ctx.integrate_warped_frame(cuda(want), cuda(w_cpu), kf_g, kfw_g)
    orc.integrate_warped_frame(want, w_cpu, kf, kfw)
    a, b = kf_g.cpu().numpy(), kf
    assert same_nan(a, b)
    mm = ~np.isnan(b)
    assert np.max(np.abs(a[mm] - b[mm]) / b[mm]) < 1e-6
    assert np.max(np.abs(kfw_g.cpu().numpy() - kfw) / kfw) < 1e-6
    if have_ref():
        kf2, kfw2 = cuda(P["WA"].copy()), cuda(np.ones_like(P["WA"]))
        kf3, kfw3 = kf2.clone(), kfw2.clone()
        refk.integrate_warped_frame(cuda(want), cuda(w_cpu), kf2, kfw2)
        ctx.integrate_warped_frame(cuda(want), cuda(w_cpu), kf3, kfw3)
        assert np.allclose(kf2.cpu().numpy(), kf3.cpu().numpy(), rtol=1e-6, equal_nan=True)


def test_visibility_ratio(ctx, P, pose):
    r_gpu, mask = ctx.visibility_ratio(cuda(P["WB"]), cuda(P["WA"]), pose["Rf"], pose["tf"], with_mask=True)
    r_cpu, mask_cpu = orc.visibility_ratio(P["WB"], P["WA"], pose["Rf"], pose["tf"], with_mask=True)
    assert abs(r_gpu - r_cpu) < 1e-4
    assert np.mean(mask.cpu().numpy() == mask_cpu) > 0.9999
    assert 0.5 < r_gpu <= 1.0
    if have_ref():
        r_ref, mask_ref = refk.visibility_ratio(cuda(P["WB"]), cuda(P["WA"]), pose["Rf"], pose["tf"], with_mask=True)
        assert abs(r_gpu - r_ref) < 2e-5
        assert np.mean(mask.cpu().numpy() == mask_ref.cpu().numpy()) > 0.99999
    # empty source: ratio 0 (warping_registration.cu:863-864)
    allnan = torch.full((ROWS, COLS), float("nan")).cuda()
    assert ctx.visibility_ratio(allnan, cuda(P["WA"]), pose["Rf"], pose["tf"]) == 0.0


def _warped(P, pose):
    W1 = orc.warp_invdepth(P["WB"], P["WA"], pose["Rp"], pose["tp"])
    I1 = orc.warp_intensity(P["IB"], W1, pose["Rp"], pose["tp"])
    return W1, I1


@pytest.mark.parametrize("nsamples,level", [(10000, 0), (19200, 0), (10000, 3), (9999999, 1)])
def test_compute_error_bit_exact(ctx, P, pose, nsamples, level):
    W1, _ = _warped(P, pose)
    a, b = W1, P["WA"]
    for _ in range(level):
        a, b = orc.pyr_down(a), orc.pyr_down(b)
    out = ctx.compute_error(cuda(a), cuda(b), nsamples).cpu().numpy()
    want = orc.compute_error(a, b, nsamples)
    assert out.shape == want.shape and np.array_equal(out, want, equal_nan=True)


def test_scale_estimation(ctx, P, pose):
    W1, I1 = _warped(P, pose)
    for (im1, im0, b0, s0) in ((I1, P["IA"], 0.0, 5.0), (W1, P["WA"], 0.0, 0.0025)):
        err = orc.compute_error(im1, im0, 10000)
        eg = cuda(err)
        b, s, nu = ctx.sigma_nu_student(eg, b0, s0)
        bo, so, nuo, _ = orc.sigma_nu_student(err, b0, s0)
        assert abs(s - so) / so < 1e-4 and abs(b - bo) < 1e-4 * so and nu == nuo
        assert ctx.nu_student(eg, b0, s0) == orc.nu_student(err, b0, s0)
        for mest in (orc.LSQ, orc.HUBER, orc.TUKEY, orc.STUDENT):
            b2, s2 = ctx.sigma_pdf(eg, b0, s0, mest)
            b2o, s2o = orc.sigma_pdf(err, b0, s0, mest)
            assert abs(s2 - s2o) / s2o < 1e-4 and abs(b2 - b2o) < 1e-4 * s2o
        if have_ref():
            br, sr, nur = refk.sigma_nu_student(eg, b0, s0)
            assert abs(s - sr) / sr < 1e-4 and abs(b - br) < 1e-4 * sr and nu == nur
            assert ctx.nu_student(eg, b0, s0) == refk.nu_student(eg, b0, s0)
            b3, s3 = refk.sigma_pdf(eg, b0, s0, orc.HUBER)
            b4, s4 = ctx.sigma_pdf(eg, b0, s0, orc.HUBER)
            assert abs(s3 - s4) / s3 < 1e-4


def test_nu_bisection_on_student_samples(ctx):
    """Heavy-tailed samples drive the bisection through its evaluated midpoints."""
    rng = np.random.default_rng(7)
    for dof in (2.5, 4.0, 7.0):
        err = (rng.standard_t(dof, 19200) * 2.0 + 0.3).astype(np.float32)
        err[::97] = np.nan
        err[5] = np.inf
        b, s, nu = ctx.sigma_nu_student(cuda(err), 0.0, 5.0)
        bo, so, nuo, _ = orc.sigma_nu_student(err, 0.0, 5.0)
        assert nu == nuo and abs(s - so) / so < 1e-4 and 2.0 <= nu <= 10.0


def test_chi_square(ctx, P, pose):
    W1, I1 = _warped(P, pose)
    eI, eW = orc.compute_error(I1, P["IA"]), orc.compute_error(W1, P["WA"])
    for mest in (orc.LSQ, orc.HUBER, orc.TUKEY, orc.STUDENT):
        c, t, n = ctx.chi_square(cuda(eI), cuda(eW), 5.0, 0.0025, mest)
        co, to, no = orc.chi_square(eI, eW, 5.0, 0.0025, mest)
        assert n == no and abs(c - co) / co < 1e-5 and abs(t - to) < 1e-5
    if have_ref():
        cr, tr, nr = refk.chi_square(cuda(eI), cuda(eW), 5.0, 0.0025, orc.STUDENT)
        c, t, n = ctx.chi_square(cuda(eI), cuda(eW), 5.0, 0.0025, orc.STUDENT)
        assert n == nr and abs(c - cr) / cr < 1e-4


@pytest.mark.parametrize("student_nu,mest,weighting", [(1, orc.STUDENT, orc.INDEPENDENT), (0, orc.HUBER, orc.INDEPENDENT),
                                                       (0, orc.TUKEY, orc.MIN_WEIGHT), (0, orc.LSQ, orc.GEOM_ONLY),
                                                       (0, orc.STUDENT, orc.PHOT_ONLY)])
def test_build_system(ctx, P, pose, student_nu, mest, weighting):
    from rgbid_slam_b200 import capi
    W1, I1 = _warped(P, pose)
    gWx, gWy = orc.gradient(P["WA"])
    gIx, gIy = orc.gradient(P["IA"])
    i = P["intr"]
    kw = dict(mestimator=mest, weighting=weighting, student_nu=student_nu, sigma_depthinv=0.0012, sigma_int=3.5,
              bias_depthinv=1e-5, bias_int=0.2, nu_depthinv=4.25, nu_int=6.5)
    po = orc.system_params(i["fx"], i["fy"], i["cx"], i["cy"], **kw)
    Ao, bo, so = orc.build_system(P["WA"], P["IA"], gWx, gWy, gIx, gIy, W1, I1, po)
    pg = capi.SystemParams(i["fx"], i["fy"], i["cx"], i["cy"], mest, weighting, student_nu, 0.0012, 3.5, 1e-5, 0.2, 4.25, 6.5)
    maps = [cuda(m) for m in (P["WA"], P["IA"], gWx, gWy, gIx, gIy, W1, I1)]
    A, b = ctx.build_system(*maps, pg)
    assert np.allclose(A, A.T)
    sg = np.concatenate([np.concatenate([A[r, r:], [b[r]]]) for r in range(6)])
    assert sums_rel_err(sg, so) < 1e-5, sums_rel_err(sg, so)   # north-star bar: residual sums within 1e-5 relative
    if have_ref():
        Ar, br = refk.build_system(*maps, po)
        sr = np.concatenate([np.concatenate([Ar[r, r:], [br[r]]]) for r in range(6)])
        print("vs reference kernels: rel err", sums_rel_err(sg, sr), " reference vs CPU oracle:", sums_rel_err(sr, so))
        assert sums_rel_err(sg, sr) < 1e-5


def test_build_system_mixed_pitches(ctx, P, pose):
    """The warped maps come from cudaMallocPitch in the reference's drivers while the keyframe maps are dense: every
    PtrStep carries its own step (rgbid_build_system_pitched)."""
    from rgbid_slam_b200 import capi
    W1, I1 = _warped(P, pose)
    gWx, gWy = orc.gradient(P["WA"])
    gIx, gIy = orc.gradient(P["IA"])
    i = P["intr"]
    pg = capi.SystemParams(i["fx"], i["fy"], i["cx"], i["cy"], 3, 0, 1, 0.0012, 3.5, 1e-5, 0.2, 4.25, 6.5)
    dense = [cuda(m) for m in (P["WA"], P["IA"], gWx, gWy, gIx, gIy, W1, I1)]
    A0, b0 = ctx.build_system(*dense, pg)

    def padded(m, extra):
        buf = torch.full((m.shape[0], m.shape[1] + extra), float("nan"), device="cuda")
        buf[:, :m.shape[1]] = m
        return buf[:, :m.shape[1]]
    mixed = [padded(m, e) for m, e in zip(dense, (0, 0, 0, 0, 0, 0, 64, 64))]
    mixed[2] = padded(dense[2], 4)
    mixed[5] = padded(dense[5], 3)     # 12 bytes more per row: not 16-byte aligned rows -> the scalar kernel
    A1, b1 = ctx.build_system(*mixed[:6], mixed[6], mixed[7], pg)
    assert np.max(np.abs(A1 - A0)) <= 1e-5 * np.max(np.abs(A0)) and np.max(np.abs(b1 - b0)) <= 1e-5 * np.max(np.abs(b0))
    mixed[5] = padded(dense[5], 8)     # all rows 16-byte aligned: the float4 kernel, bit-identical sums
    A2, b2 = ctx.build_system(*mixed[:6], mixed[6], mixed[7], pg)
    assert np.array_equal(A2, A0) and np.array_equal(b2, b0)


def test_build_system_all_invalid(ctx):
    from rgbid_slam_b200 import capi
    z = torch.full((16, 32), float("nan")).cuda()
    pg = capi.SystemParams(525, 525, 15.5, 7.5, 3, 0, 1, 0.0025, 5.0, 0, 0, 5, 5)
    A, b = ctx.build_system(z, z, z, z, z, z, z, z, pg)
    assert not A.any() and not b.any()


def test_vmap_nmap(ctx, P):
    i = P["intr"]
    W = P["WA"]
    gx, gy = orc.gradient(W)
    v = ctx.create_vmap(cuda(W), i["fx"], i["fy"], i["cx"], i["cy"]).cpu().numpy()
    vo = orc.vmap(W, i["fx"], i["fy"], i["cx"], i["cy"])
    assert same_nan(v[:ROWS], vo[:ROWS])
    m = ~np.isnan(vo[:ROWS])
    for k in range(3):
        a, b = v[k * ROWS:(k + 1) * ROWS][m], vo[k * ROWS:(k + 1) * ROWS][m]
        assert np.max(np.abs(a - b)) < 1e-5
    n = ctx.create_nmap_gradients(cuda(W), cuda(gx), cuda(gy), i["fx"], i["fy"], i["cx"], i["cy"]).cpu().numpy()
    no = orc.nmap_gradients(W, gx, gy, i["fx"], i["fy"], i["cx"], i["cy"])
    assert np.mean(np.isnan(n[:ROWS]) == np.isnan(no[:ROWS])) > 0.9999
    m = ~(np.isnan(n[:ROWS]) | np.isnan(no[:ROWS]))
    for k in range(3):
        assert np.max(np.abs(n[k * ROWS:(k + 1) * ROWS][m] - no[k * ROWS:(k + 1) * ROWS][m])) < 1e-5
    if have_ref():
        nr = refk.nmap_gradients(cuda(W), cuda(gx), cuda(gy), i["fx"], i["fy"], i["cx"], i["cy"]).cpu().numpy()
        assert np.mean(np.isnan(n[:ROWS]) == np.isnan(nr[:ROWS])) > 0.99999
