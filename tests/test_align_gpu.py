"""Parity of the fused, device-resident coarse-to-fine aligner and of the tracker state machine (through the C
ABI) against the CPU oracle and the reference's own kernels (oracle/_ref, when present).

Bars (BASELINE.json north_star): recovered SE(3) within 1e-4 rad / 1e-4 m, residual sums within 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

from util import pair_maps, cuda, rot_angle, sums_rel_err
import oracle as orc
from oracle import ref as refk
from oracle.tracker import OracleTracker
from rgbid_slam_b200 import capi, host, synth

pytestmark = pytest.mark.gpu

POSE_TOL_M, POSE_TOL_RAD, SUMS_TOL = 1e-4, 1e-4, 1e-5


def _gpu_align(ctx, P, rows, cols, levels, mode, iterations=None, from_rgbd=True, batch=1, **kw):
    cfg = host.make_align_config(rows, cols, levels, mode, batch=batch, iterations=iterations, **P["intr"], **kw)
    al = host.Aligner(ctx, cfg)
    WA, IA = cuda(P["WA"]), cuda(P["IA"])
    for b in range(batch):
        al.set_keyframe(b, WA, IA)
        if from_rgbd:
            al.set_current_rgbd(b, cuda(P["dB"]), cuda(P["cB"]))
        else:
            al.set_current(b, cuda(P["WB"]), cuda(P["IB"]))
    return al


def _oracle_align(P, rows, cols, levels, mode, iterations, kind="cpu", **kw):
    i = P["intr"]
    cfg = orc.make_config(rows, cols, levels, mode, iterations, i["fx"], i["fy"], i["cx"], i["cy"], **kw)
    tracker = (mode == orc.MODE_TRACKER)
    if kind == "cpu":
        return orc.align(cfg, orc.prepare_keyframe(P["WA"], P["IA"], levels, tracker), orc.prepare_current(P["WB"], P["IB"], levels))
    return refk.align(cfg, refk.prepare_keyframe(cuda(P["WA"]), cuda(P["IA"]), levels, tracker),
                      refk.prepare_current(cuda(P["WB"]), cuda(P["IB"]), levels))


def _check(out, ref, b=0, label=""):
    dt = float(np.linalg.norm(out["t"][b] - ref["t"]))
    ang = rot_angle(out["R"][b], ref["R"])
    print("%s: |dt| = %.2e m, angle = %.2e rad" % (label, dt, ang))
    assert out["status"][b] == 0 and ref["status"] == 0
    assert dt < POSE_TOL_M and ang < POSE_TOL_RAD, (label, dt, ang)
    return dt, ang


@pytest.mark.parametrize("mode,levels,its", [(capi.MODE_ALIGN, 4, [5, 5, 3, 0]), (capi.MODE_TRACKER, 3, [10, 5, 3])])
@pytest.mark.parametrize("noise", [False, True])
def test_align_pair_vs_oracle_and_reference(ctx, mode, levels, its, noise):
    rows, cols = 480, 640
    P = pair_maps(seed=20261018 + int(noise), rows=rows, cols=cols, noise=noise)
    al = _gpu_align(ctx, P, rows, cols, levels, mode, its)
    out = al.run(want_trace=True)
    ref = _oracle_align(P, rows, cols, levels, mode, its)
    _check(out, ref, label="vs CPU oracle")
    # ground truth is recovered too (mm depth quantisation bounds the accuracy)
    assert np.linalg.norm(out["t"][0] - P["t_ab"]) < 5e-4 and rot_angle(out["R"][0], P["R_ab"]) < 5e-4
    tr = out["trace"][0]
    # first iteration: identical pose on both sides -> the 27 sums must agree to the north-star bar
    assert sums_rel_err(tr[0]["sums27"], ref["trace"][0]["sums27"]) < SUMS_TOL
    assert tr[0]["nu_depthinv"] == ref["trace"][0]["nu_depthinv"] and tr[0]["nu_int"] == ref["trace"][0]["nu_int"]
    if mode == capi.MODE_TRACKER:
        assert abs(tr[0]["sigma_int"] - ref["trace"][0]["sigma_int"]) / ref["trace"][0]["sigma_int"] < 1e-4
        assert tr[0]["irls_iters_int"] == ref["trace"][0]["irls_iters_int"]
        # covariance pass + end-of-frame chi^2
        # (J^T r is ~0 at convergence -- pure cancellation noise -- and unused by the covariance: compare A only)
        assert sums_rel_err(tr[-1]["sums27"], ref["cov_sums27"], ignore_b=True) < 1e-4
        assert abs(out["stats"][0][0] - ref["chi_square"]) / ref["chi_square"] < 1e-3
        assert abs(out["stats"][0][2] - ref["ndof"]) / ref["ndof"] < 1e-3
    cov_rel = np.abs(out["cov"][0] - ref["cov"]).max() / np.abs(ref["cov"]).max()
    assert cov_rel < 1e-3, cov_rel
    if refk.available():
        r2 = _oracle_align(P, rows, cols, levels, mode, its, kind="ref")
        _check(out, r2, label="vs reference CUDA kernels")
        e = sums_rel_err(tr[0]["sums27"], r2["trace"][0]["sums27"])
        print("first-iteration sums vs reference kernels: %.2e (reference vs CPU oracle: %.2e)"
              % (e, sums_rel_err(r2["trace"][0]["sums27"], ref["trace"][0]["sums27"])))
        assert e < SUMS_TOL
    al.close()


def test_align_float_maps_equals_rgbd_ingest(ctx):
    rows, cols = 240, 320
    P = pair_maps(seed=77, rows=rows, cols=cols)
    a = _gpu_align(ctx, P, rows, cols, 4, capi.MODE_ALIGN, from_rgbd=True).run()
    b = _gpu_align(ctx, P, rows, cols, 4, capi.MODE_ALIGN, from_rgbd=False).run()
    assert np.linalg.norm(a["t"][0] - b["t"][0]) < 1e-6 and rot_angle(a["R"][0], b["R"][0]) < 1e-6


def test_align_is_deterministic_and_batch_invariant(ctx):
    rows, cols = 240, 320
    P = pair_maps(seed=78, rows=rows, cols=cols, noise=True)
    one = _gpu_align(ctx, P, rows, cols, 3, capi.MODE_TRACKER, batch=1).run()
    al = _gpu_align(ctx, P, rows, cols, 3, capi.MODE_TRACKER, batch=5)
    many1, many2 = al.run(), al.run()
    for b in range(5):
        assert np.array_equal(many1["R"][b], many2["R"][b]) and np.array_equal(many1["t"][b], many2["t"][b])  # replay
        # a different batch size changes the grid (partial-sum order), not the result beyond rounding
        assert np.allclose(many1["t"][b], one["t"][0], atol=1e-6) and np.allclose(many1["R"][b], one["R"][0], atol=1e-6)


def test_texture_unit_and_software_sampler_agree(ctx, monkeypatch):
    """The fused kernels gather through texture objects by default; RGBID_SAMPLER=soft selects the software
    sampler that reproduces the texture unit's 8-bit weights.  Both must give the same alignment."""
    rows, cols = 480, 640
    P = pair_maps(seed=90, rows=rows, cols=cols, noise=True)
    tex = _gpu_align(ctx, P, rows, cols, 3, capi.MODE_TRACKER).run(want_trace=True)
    monkeypatch.setenv("RGBID_SAMPLER", "soft")
    soft = _gpu_align(ctx, P, rows, cols, 3, capi.MODE_TRACKER).run(want_trace=True)
    monkeypatch.delenv("RGBID_SAMPLER")
    assert sums_rel_err(tex["trace"][0][0]["sums27"], soft["trace"][0][0]["sums27"]) < 1e-6
    assert np.linalg.norm(tex["t"][0] - soft["t"][0]) < 1e-6 and rot_angle(tex["R"][0], soft["R"][0]) < 1e-6


def test_align_host_upload_path(ctx):
    rows, cols = 240, 320
    P = pair_maps(seed=79, rows=rows, cols=cols)
    cfg = host.make_align_config(rows, cols, 4, capi.MODE_ALIGN, **P["intr"])
    al = host.Aligner(ctx, cfg)
    al.set_keyframe(0, P["WA"], P["IA"])          # numpy -> host path of the ABI
    al.set_current_rgbd(0, P["dB"], P["cB"])
    out = al.run()
    ref = _oracle_align(P, rows, cols, 4, orc.MODE_ALIGN, [5, 5, 3, 0])
    _check(out, ref, label="host upload")


def test_align_lost_on_empty_keyframe(ctx):
    """All-NaN inputs -> singular system -> NaN pose -> status RGBID_ERR_NAN, pose restored, cov = 100 I
    (src/visodo.cpp:1265-1274)."""
    rows, cols = 120, 160
    cfg = host.make_align_config(rows, cols, 3, capi.MODE_TRACKER, **synth.intrinsics_for(rows, cols))
    al = host.Aligner(ctx, cfg)
    nan = torch.full((rows, cols), float("nan")).cuda()
    al.set_keyframe(0, nan, nan)
    al.set_current(0, nan, nan)
    out = al.run()
    assert out["status"][0] == capi.ERR_NAN
    assert np.array_equal(out["R"][0], np.eye(3)) and np.array_equal(out["cov"][0], 100 * np.eye(6))


def test_align_1280x960_5_levels(ctx):
    """Config 4 geometry (the reference hard-codes 640x480; the new path is size / level parametric)."""
    rows, cols = 960, 1280
    P = pair_maps(seed=80, rows=rows, cols=cols)
    al = _gpu_align(ctx, P, rows, cols, 5, capi.MODE_ALIGN, [5, 5, 3, 0, 0])
    out = al.run()
    ref = _oracle_align(P, rows, cols, 5, orc.MODE_ALIGN, [5, 5, 3, 0, 0])
    _check(out, ref, label="1280x960x5")


def test_align_ragged_size_uses_the_scalar_generic_kernels(ctx):
    """322 x 242, 2 levels: cols % 4 != 0 and pitch != 4 * cols at both levels, so every kernel of the schedule takes
    its scalar / pitched fallback (generic system kernel with VEC = 1, scalar pyramid / gradient kernels)."""
    rows, cols, levels, its = 242, 322, 2, [6, 4]
    # the synthetic scene is rendered in 8 x 8 blocks: render 248 x 328 and keep the top-left window (same pixel
    # coordinates, hence the same intrinsics)
    P = pair_maps(seed=20261021, rows=248, cols=328, noise=True)
    for k in ("dA", "dB", "cA", "cB", "WA", "WB", "IA", "IB"):
        P[k] = np.ascontiguousarray(P[k][:rows, :cols])
    out = _gpu_align(ctx, P, rows, cols, levels, capi.MODE_TRACKER, its).run(want_trace=True)
    ref = _oracle_align(P, rows, cols, levels, orc.MODE_TRACKER, its)
    _check(out, ref, label="ragged vs CPU oracle")
    assert sums_rel_err(out["trace"][0][0]["sums27"], ref["trace"][0]["sums27"]) < SUMS_TOL


def test_huber_system_through_api(ctx):
    """Config 3: buildSystemGridStride(HUBER) with sigma from computeSigmaPdf, on freiburg1 intrinsics."""
    P = pair_maps(seed=81, rows=480, cols=640, noise=True)
    i = P["intr"]
    Rp, tp = orc.projective_pose(np.eye(3), np.zeros(3), i["fx"], i["fy"], i["cx"], i["cy"], inverse=True)
    W1g = ctx.warp_invdepth(cuda(P["WB"]), cuda(P["WA"]), Rp, tp)
    I1g = ctx.warp_intensity(cuda(P["IB"]), W1g, Rp, tp)
    bI, sI = ctx.sigma_pdf(ctx.compute_error(I1g, cuda(P["IA"]), 10000), 0.0, 5.0, capi.HUBER)
    bW, sW = ctx.sigma_pdf(ctx.compute_error(W1g, cuda(P["WA"]), 10000), 0.0, 0.0025, capi.HUBER)
    gWx, gWy = ctx.compute_gradient(cuda(P["WA"]))
    gIx, gIy = ctx.compute_gradient(cuda(P["IA"]))
    pg = capi.SystemParams(i["fx"], i["fy"], i["cx"], i["cy"], capi.HUBER, capi.INDEPENDENT, 0, sW, sI, bW, bI, 5, 5)
    A, b = ctx.build_system(cuda(P["WA"]), cuda(P["IA"]), gWx, gWy, gIx, gIy, W1g, I1g, pg)
    W1 = orc.warp_invdepth(P["WB"], P["WA"], Rp, tp)
    I1 = orc.warp_intensity(P["IB"], W1, Rp, tp)
    bIo, sIo = orc.sigma_pdf(orc.compute_error(I1, P["IA"], 10000), 0.0, 5.0, orc.HUBER)
    bWo, sWo = orc.sigma_pdf(orc.compute_error(W1, P["WA"], 10000), 0.0, 0.0025, orc.HUBER)
    assert abs(sI - sIo) / sIo < 1e-3 and abs(sW - sWo) / sWo < 1e-3
    gx, gy = orc.gradient(P["WA"])
    hx, hy = orc.gradient(P["IA"])
    po = orc.system_params(i["fx"], i["fy"], i["cx"], i["cy"], mestimator=orc.HUBER, student_nu=0, sigma_depthinv=sW,
                           sigma_int=sI, bias_depthinv=bW, bias_int=bI)
    Ao, bo, so = orc.build_system(P["WA"], P["IA"], gx, gy, hx, hy, W1, I1, po)
    sg = np.concatenate([np.concatenate([A[r, r:], [b[r]]]) for r in range(6)])
    assert sums_rel_err(sg, so) < 1e-4
    x = np.linalg.solve(A, b)
    assert np.linalg.norm(x[:3] - np.linalg.solve(Ao, bo)[:3]) < 1e-5


@pytest.mark.parametrize("kind", ["cpu", "ref"])
def test_tracker_sequence(ctx, kind):
    """20-frame synthetic TUM-style sequence through rgbid_tracker_track (host buffers in, like the reference's
    upload + trackNewFrame) against the restated trackNewFrame on the CPU oracle / the reference's kernels."""
    if kind == "ref" and not refk.available():
        pytest.skip("oracle/_ref/libref_oracle.so not present")
    rows, cols, n = 240, 320, 20
    seq = synth.make_sequence(seed=4242, n_frames=n, rows=rows, cols=cols, noise=True)
    intr = seq["intr"]
    acfg = host.make_align_config(rows, cols, 3, capi.MODE_TRACKER, batch=2, **intr)
    trk = host.Tracker(ctx, host.make_tracker_config(acfg))
    ot = OracleTracker(rows, cols, intr, levels=3, iterations=(10, 5, 3), kind=kind)
    worst_t = worst_r = 0.0
    kf_events = 0
    for k in range(n):
        d, c = seq["depth"][k], seq["rgb"][k]
        dd = torch.stack([d, d]).contiguous()
        cc = torch.stack([c, c]).contiguous()
        res = trk.track(dd, cc)  # CPU tensors -> host path
        if kind == "cpu":
            o = ot.track(d.numpy().astype(np.uint16), c.numpy())
        else:
            o = ot.track(d.cuda(), c.cuda())
        for b in range(2):
            r = res[b]
            assert r.status == o["status"] == 0
            dt = np.linalg.norm(np.array(r.t[:]) - o["t"])
            ang = rot_angle(np.array(r.R[:]).reshape(3, 3), o["R"])
            worst_t, worst_r = max(worst_t, dt), max(worst_r, ang)
            assert r.new_odo_keyframe == o["new_odo_keyframe"] and r.new_integr_keyframe == o["new_integr_keyframe"], k
            if k > 0:
                assert abs(r.visibility_odo - o["visibility_odo"]) < 2e-4
                assert abs(r.visibility_integr - o["visibility_integr"]) < 2e-4
        # both streams carry the same data
        assert np.array_equal(np.array(res[0].t[:]), np.array(res[1].t[:]))
        kf_events += o["new_odo_keyframe"]
        # ground truth
        gt_R, gt_t = synth.relative_pose(seq["poses"][0], seq["poses"][k])
        assert np.linalg.norm(np.array(res[0].t[:]) - gt_t) < 5e-3
    print("tracker vs %s: worst |dt| = %.2e m, worst angle = %.2e rad, odometry keyframes = %d" % (kind, worst_t, worst_r, kf_events))
    assert worst_t < POSE_TOL_M and worst_r < POSE_TOL_RAD
    # fused integration keyframe agrees with the restatement
    fused = trk.keyframe_map(0, 0).cpu().numpy()
    want = ot.intW if kind == "cpu" else ot.intW.cpu().numpy()
    assert np.mean(np.isnan(fused) == np.isnan(want)) > 0.999
    m = ~(np.isnan(fused) | np.isnan(want))
    assert np.mean(np.abs(fused[m] - want[m]) / want[m] < 1e-4) > 0.999
    trk.close()


def test_tracker_prefetch_is_the_same_tracker(ctx):
    """rgbid_tracker_prefetch uploads frame k + 1 on a copy stream while frame k is tracked; the results must be
    bit-identical to the plain host path (same kernels on the same bytes)."""
    rows, cols, n = 240, 320, 6
    seq = synth.make_sequence(seed=77, n_frames=n, rows=rows, cols=cols, noise=True)
    acfg = host.make_align_config(rows, cols, 3, capi.MODE_TRACKER, batch=2, **seq["intr"])
    frames = [(torch.stack([seq["depth"][k]] * 2).contiguous().pin_memory(), torch.stack([seq["rgb"][k]] * 2).contiguous().pin_memory())
              for k in range(n)]
    poses = []
    for use_prefetch in (False, True):
        trk = host.Tracker(ctx, host.make_tracker_config(acfg))
        out = []
        for k in range(n):
            if use_prefetch and k + 1 < n:
                trk.prefetch(*frames[k + 1])
            res = trk.track(*frames[k])
            out.append((np.array(res[0].R[:]), np.array(res[0].t[:]), res[0].new_odo_keyframe))
        trk.close()
        poses.append(out)
    for (Ra, ta, ka), (Rb, tb, kb) in zip(*poses):
        assert np.array_equal(Ra, Rb) and np.array_equal(ta, tb) and ka == kb


def test_stale_prefetch_is_dropped(ctx):
    """A prefetched frame is good for the current or the next track call only: a host buffer that was prefetched,
    not tracked, refilled and tracked later must be read again."""
    rows, cols, n = 240, 320, 4
    seq = synth.make_sequence(seed=78, n_frames=n, rows=rows, cols=cols, noise=True)
    acfg = host.make_align_config(rows, cols, 3, capi.MODE_TRACKER, batch=1, **seq["intr"])
    frames = [(seq["depth"][k][None].contiguous().pin_memory(), seq["rgb"][k][None].contiguous().pin_memory()) for k in range(n)]
    trk = host.Tracker(ctx, host.make_tracker_config(acfg))
    want = [np.array(trk.track(*frames[k])[0].t[:]) for k in range(n)]
    trk.close()
    trk = host.Tracker(ctx, host.make_tracker_config(acfg))
    sd, sc = frames[1][0].clone().pin_memory(), frames[1][1].clone().pin_memory()
    trk.track(*frames[0])
    trk.prefetch(sd, sc)          # frame 1's bytes are uploaded ...
    trk.track(*frames[1])         # ... but another buffer is tracked (twice)
    trk.track(*frames[2])
    sd.copy_(frames[3][0]); sc.copy_(frames[3][1])
    got = np.array(trk.track(sd, sc)[0].t[:])  # the refilled buffer must be read again, not served from the old upload
    trk.close()
    assert np.array_equal(got, want[3])


def test_trace_is_only_recorded_on_request(ctx):
    """rgbid_aligner_run records the per-iteration trace exactly when trace_out != NULL (rgbid_aligner_set_trace): a run
    without it leaves the device-side trace of the previous traced run untouched and still returns the right pose."""
    import ctypes as C
    rows, cols, levels, its = 240, 320, 3, [4, 3, 2]
    PA = pair_maps(seed=31, rows=rows, cols=cols, noise=True)
    PB = pair_maps(seed=32, rows=rows, cols=cols, noise=True)
    al = _gpu_align(ctx, PA, rows, cols, levels, capi.MODE_TRACKER, its)
    outA = al.run(want_trace=True)
    sumsA = np.array([np.asarray(tr["sums27"]) for tr in outA["trace"][0]])
    assert np.abs(sumsA).max() > 0
    # pair B, no trace
    al.set_keyframe(0, cuda(PB["WA"]), cuda(PB["IA"]))
    al.set_current_rgbd(0, cuda(PB["dB"]), cuda(PB["cB"]))
    outB = al.run(want_trace=False)
    refB = _oracle_align(PB, rows, cols, levels, orc.MODE_TRACKER, its)
    _check(outB, refB, label="untraced run")
    n = sum(its) + 1
    raw = (capi.IterTrace * n)()
    R, t = np.zeros((1, 9)), np.zeros((1, 3))
    capi.check(al.lib.rgbid_aligner_fetch(al.h, R.ctypes.data_as(capi.c_double_p), t.ctypes.data_as(capi.c_double_p), None, None, raw),
               "aligner_fetch")
    still = np.array([np.array(raw[i].sums27[:]) for i in range(n)])
    assert np.array_equal(still[:len(sumsA)], sumsA), "the untraced run overwrote the trace"
    # pair B again, traced: now the records are B's
    outB2 = al.run(want_trace=True)
    sumsB = np.array([np.asarray(tr["sums27"]) for tr in outB2["trace"][0]])
    assert not np.array_equal(sumsB[0], sumsA[0])
    assert sums_rel_err(sumsB[0], refB["trace"][0]["sums27"]) < SUMS_TOL
    assert np.allclose(outB2["t"], outB["t"], atol=1e-12) and np.allclose(outB2["R"], outB["R"], atol=1e-12)
