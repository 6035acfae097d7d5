"""Config-level parity: the workloads of BASELINE.json `configs` as written, through the C ABI,
against the restated trackNewFrame on the CPU oracle and on the reference's own CUDA kernels (oracle/_ref).

  config 2  640x480 x 120 frames, 4 levels, "full GN convergence" (|x| < 1e-6 or 20 iterations per level)
  config 3  freiburg1 stream, photometric + geometric, buildSystemGridStride(HUBER) + computeSigmaPdf call sequence
  config 4  1280x960, 5 levels, tracker mode with keyframe depth fusion
  config 5 / bench.py  640x480, 4 levels {10,5,3,0}, >= 8 distinct streams in one batch
plus the tracker's lost -> re-acquire path, the reference's CHI_SQUARED termination and the copy / fill bridge ops.

Bars (BASELINE.json north_star): SE(3) within 1e-4 m / 1e-4 rad per frame, identical keyframe decisions."""
import numpy as np
import pytest
import torch

from util import cuda, rot_angle, pair_maps, sums_rel_err
import oracle as orc
from oracle import ref as refk
from oracle.tracker import OracleTracker
from rgbid_slam_b200 import capi, host, synth

pytestmark = pytest.mark.gpu

POSE_TOL_M, POSE_TOL_RAD = 1e-4, 1e-4


def _need(kind):
    if kind == "ref" and not refk.available():
        pytest.skip("oracle/_ref/libref_oracle.so not present")


def _oracle_frame(kind, d, c):
    """One frame in the form the oracle back end wants: numpy on the CPU, CUDA tensors for the reference kernels."""
    if kind == "cpu":
        return d.cpu().numpy().astype(np.uint16), c.cpu().numpy()
    return d.cuda(), c.cuda()


def _run_streams(ctx, kind, rows, cols, levels, its, n_frames, seeds, noise=True, check_fused=False, frames_hook=None,
                 expect_lost=(), **cfg_kw):
    """Tracks len(seeds) DISTINCT synthetic sequences as one batch and each of them with its own OracleTracker; returns
    the worst pose differences.  frames_hook(k, depth, rgb) may corrupt a frame (same bytes go to both sides)."""
    B = len(seeds)
    seqs = [synth.make_sequence(seed=s, n_frames=n_frames, rows=rows, cols=cols, noise=noise, device="cuda") for s in seeds]
    intr = seqs[0]["intr"]
    acfg = host.make_align_config(rows, cols, levels, capi.MODE_TRACKER, batch=B, iterations=its, **intr, **cfg_kw)
    trk = host.Tracker(ctx, host.make_tracker_config(acfg))
    okw = dict(termination=cfg_kw.get("termination", 0), conv_eps=cfg_kw.get("conv_eps", 0.0),
               warp_first=cfg_kw.get("warp_first", 0))
    ots = [OracleTracker(rows, cols, intr, levels=levels, iterations=tuple(its), kind=kind, **okw) for _ in seeds]
    worst_t = worst_r = 0.0
    events = dict(odo=0, integr=0, lost=0, again=0)
    for k in range(n_frames):
        dd = torch.stack([q["depth"][k] for q in seqs]).contiguous()
        cc = torch.stack([q["rgb"][k] for q in seqs]).contiguous()
        if frames_hook is not None:
            dd, cc = frames_hook(k, dd, cc)
        res = trk.track(dd.cpu(), cc.cpu())  # host buffers in, like the reference's upload + trackNewFrame
        for b in range(B):
            r = res[b]
            o = ots[b].track(*_oracle_frame(kind, dd[b], cc[b]))
            assert (r.status == 0) == (o["status"] == 0), (k, b, r.status, o["status"])
            assert r.lost_again == o["lost_again"], (k, b)
            assert r.new_odo_keyframe == o["new_odo_keyframe"] and r.new_integr_keyframe == o["new_integr_keyframe"], (k, b)
            assert r.frame_index == o["frame_index"], (k, b)
            events["odo"] += o["new_odo_keyframe"]; events["integr"] += o["new_integr_keyframe"]
            events["lost"] += int(o["status"] != 0); events["again"] += o["lost_again"]
            dt = float(np.linalg.norm(np.array(r.t[:]) - o["t"]))
            ang = rot_angle(np.array(r.R[:]).reshape(3, 3), o["R"])
            worst_t, worst_r = max(worst_t, dt), max(worst_r, ang)
            assert dt < POSE_TOL_M and ang < POSE_TOL_RAD, (k, b, dt, ang)
            if k > 0 and r.status == 0:
                assert abs(r.visibility_odo - o["visibility_odo"]) < 2e-4
                assert abs(r.visibility_integr - o["visibility_integr"]) < 2e-4
            if (k, b) in expect_lost:
                assert r.status != 0
    if check_fused:
        for b in range(B):
            fused = trk.keyframe_map(0, b).cpu().numpy()
            want = ots[b].intW if kind == "cpu" else ots[b].intW.cpu().numpy()
            assert np.mean(np.isnan(fused) == np.isnan(want)) > 0.999
            m = ~(np.isnan(fused) | np.isnan(want))
            assert np.mean(np.abs(fused[m] - want[m]) / want[m] < 1e-4) > 0.999
    trk.close()
    print("vs %s: %d streams x %d frames, worst |dt| = %.2e m, worst angle = %.2e rad, events %s"
          % (kind, B, n_frames, worst_t, worst_r, events))
    return worst_t, worst_r, events


# ---- the bench.py / config-5 workload: 640x480, 4 levels {10,5,3,0}, distinct streams in one batch ----------------
@pytest.mark.parametrize("kind,nstreams", [("ref", 8), ("cpu", 2)])
def test_bench_workload_distinct_streams(ctx, kind, nstreams):
    _need(kind)
    seeds = [9100 + 17 * i for i in range(nstreams)]
    _run_streams(ctx, kind, 480, 640, 4, [10, 5, 3, 0], 30, seeds, check_fused=True)


# ---- config 2: 120 frames, 4 levels, full Gauss-Newton convergence -------------------------------------------------
def test_config2_full_convergence_120_frames(ctx):
    _need("ref")
    wt, wr, ev = _run_streams(ctx, "ref", 480, 640, 4, [20, 20, 20, 20], 120, [20261019], check_fused=True,
                              termination=capi.TERM_CONVERGENCE, conv_eps=1e-6)
    assert ev["lost"] == 0


def test_config2_full_convergence_vs_cpu_oracle(ctx):
    _run_streams(ctx, "cpu", 480, 640, 4, [20, 20, 20, 20], 4, [20261019], termination=capi.TERM_CONVERGENCE, conv_eps=1e-6)


def test_convergence_schedule_stops_levels_early(ctx):
    """A loose threshold must end levels before their budget, identically on both sides (the decision is taken on the
    same |x| up to rounding; the pose bar holds either way)."""
    rows, cols = 480, 640
    P = pair_maps(seed=31, rows=rows, cols=cols, noise=True)
    its = [20, 20, 20, 0]
    cfg = host.make_align_config(rows, cols, 4, capi.MODE_TRACKER, iterations=its, termination=capi.TERM_CONVERGENCE,
                                 conv_eps=2e-5, **P["intr"])
    al = host.Aligner(ctx, cfg)
    al.set_keyframe(0, cuda(P["WA"]), cuda(P["IA"]))
    al.set_current(0, cuda(P["WB"]), cuda(P["IB"]))
    out = al.run(want_trace=True)
    i = P["intr"]
    ocfg = orc.make_config(rows, cols, 4, orc.MODE_TRACKER, its, i["fx"], i["fy"], i["cx"], i["cy"], termination=2, conv_eps=2e-5)
    ref = orc.align(ocfg, orc.prepare_keyframe(P["WA"], P["IA"], 4, True), orc.prepare_current(P["WB"], P["IB"], 4))
    done = out["iterations_done"][0]
    print("iterations per level (L0..L3):", list(done[:4]), "oracle trace entries:", len(ref["trace"]))
    assert int(done[:3].sum()) < 60 and int(done[:3].sum()) == len(ref["trace"])
    assert np.linalg.norm(out["t"][0] - ref["t"]) < POSE_TOL_M and rot_angle(out["R"][0], ref["R"]) < POSE_TOL_RAD
    assert len(out["trace"][0]) == len(ref["trace"]) + 1  # + covariance pass
    al.close()


# ---- the reference's CHI_SQUARED termination (src/visodo.cpp:1134-1164) -------------------------------------------
@pytest.mark.parametrize("warp_first", [0, 1])
@pytest.mark.parametrize("kind", ["cpu", "ref"])
def test_chi_squared_termination(ctx, kind, warp_first):
    _need(kind)
    rows, cols, levels, its = 480, 640, 3, [10, 5, 3]
    P = pair_maps(seed=32 + warp_first, rows=rows, cols=cols, noise=True)
    cfg = host.make_align_config(rows, cols, levels, capi.MODE_TRACKER, iterations=its, termination=capi.TERM_CHI_SQUARED,
                                 warp_first=warp_first, **P["intr"])
    al = host.Aligner(ctx, cfg)
    al.set_keyframe(0, cuda(P["WA"]), cuda(P["IA"]))
    al.set_current(0, cuda(P["WB"]), cuda(P["IB"]))
    out = al.run()
    i = P["intr"]
    ocfg = orc.make_config(rows, cols, levels, orc.MODE_TRACKER, its, i["fx"], i["fy"], i["cx"], i["cy"], termination=1,
                           warp_first=warp_first)
    if kind == "cpu":
        ref = orc.align(ocfg, orc.prepare_keyframe(P["WA"], P["IA"], levels, True), orc.prepare_current(P["WB"], P["IB"], levels))
    else:
        ref = refk.align(ocfg, refk.prepare_keyframe(cuda(P["WA"]), cuda(P["IA"]), levels, True),
                         refk.prepare_current(cuda(P["WB"]), cuda(P["IB"]), levels))
    done = [int(v) for v in out["iterations_done"][0][:levels]]
    want = [sum(1 for T in ref["trace"] if T["level"] == l) for l in range(levels)]
    print("CHI_SQUARED %s warp_first=%d: iterations per level %s, %s oracle %s" % (kind, warp_first, done, kind, want))
    # `RMSE > RMSE_prev` compares float sums over 300 k residuals that agree to ~1e-7 once a level has converged, so
    # WHEN it fires is rounding noise (measured on B200: [1, 5, 3] here against [2, 3, 3] in the CPU oracle for the
    # same pair); what has to hold is that it fires on both sides (levels end early, the last increment is undone) and
    # that the recovered pose is the same
    assert sum(done) < sum(its) and sum(want) < sum(its), "the test never fired"
    assert all(d >= 1 for d in done)
    assert np.linalg.norm(out["t"][0] - ref["t"]) < POSE_TOL_M and rot_angle(out["R"][0], ref["R"]) < POSE_TOL_RAD
    al.close()


# ---- config 4: 1280x960, 5 levels, tracker mode, keyframe depth fusion on -----------------------------------------
@pytest.mark.parametrize("kind,nstreams,nframes", [("ref", 2, 8), ("cpu", 1, 4)])
def test_config4_1280x960_tracker_with_fusion(ctx, kind, nstreams, nframes):
    _need(kind)
    _run_streams(ctx, kind, 960, 1280, 5, [10, 5, 3, 0, 0], nframes, [4400 + i for i in range(nstreams)], check_fused=True)


# ---- lost -> lost again -> re-acquire (src/visodo.cpp:2056-2117) --------------------------------------------------
@pytest.mark.parametrize("kind", ["cpu", "ref"])
def test_tracker_lost_and_reacquire(ctx, kind):
    """Stream 0 receives an all-zero depth frame at k = 5: the alignment fails (lost), the keyframes are re-saved from
    that empty frame, so frame 6 fails AGAIN (no constraint, no keyframe reset, global_time_ stands still) and re-saves
    them from a good frame; frame 7 is tracked again.  Stream 1 is undisturbed and must not notice."""
    _need(kind)

    def hook(k, dd, cc):
        if k == 5:
            dd = dd.clone()
            dd[0].zero_()
        return dd, cc

    wt, wr, ev = _run_streams(ctx, kind, 240, 320, 3, [10, 5, 3], 12, [501, 502], frames_hook=hook,
                              expect_lost=((5, 0), (6, 0)))
    assert ev["lost"] == 2 and ev["again"] == 1


# ---- config 3: freiburg1 stream through the un-fused bridge calls, Huber weights, sigma from computeSigmaPdf ---------
class _GpuOps:
    """The reference's call sequence on the new library (one rgbid_* entry per bridge function)."""

    def __init__(self, ctx):
        self.c = ctx
        self.up = cuda
        self.warp_invdepth, self.warp_intensity = ctx.warp_invdepth, ctx.warp_intensity
        self.compute_error, self.sigma_pdf = ctx.compute_error, ctx.sigma_pdf
        self.gradient, self.pyr_down = ctx.compute_gradient, ctx.pyr_down
        self.build_system = ctx.build_system
        self.params = lambda **k: capi.SystemParams(k["fx"], k["fy"], k["cx"], k["cy"], capi.HUBER, capi.INDEPENDENT, 0,
                                                    k["sW"], k["sI"], k["bW"], k["bI"], 5, 5)


class _CpuOps:
    def __init__(self):
        self.up = lambda a: a
        self.warp_invdepth, self.warp_intensity = orc.warp_invdepth, orc.warp_intensity
        self.compute_error, self.sigma_pdf = orc.compute_error, orc.sigma_pdf
        self.gradient, self.pyr_down = orc.gradient, orc.pyr_down
        self.params = lambda **k: orc.system_params(k["fx"], k["fy"], k["cx"], k["cy"], mestimator=orc.HUBER, student_nu=0,
                                                    sigma_depthinv=k["sW"], sigma_int=k["sI"], bias_depthinv=k["bW"], bias_int=k["bI"])

    def build_system(self, *a):
        A, b, _ = orc.build_system(*a)
        return A, b


class _RefOps(_GpuOps):
    def __init__(self):
        self.up = cuda
        self.warp_invdepth, self.warp_intensity = refk.warp_invdepth, refk.warp_intensity
        self.compute_error, self.sigma_pdf = refk.compute_error, refk.sigma_pdf
        self.gradient, self.pyr_down = refk.gradient, refk.pyr_down
        self.build_system = refk.build_system
        self.params = lambda **k: orc.system_params(k["fx"], k["fy"], k["cx"], k["cy"], mestimator=orc.HUBER, student_nu=0,
                                                    sigma_depthinv=k["sW"], sigma_int=k["sI"], bias_depthinv=k["bW"], bias_int=k["bI"])


def _huber_frame_to_frame(ops, W_kf, I_kf, W_cur, I_cur, intr, levels, its, nsamples=10000):
    """Coarse-to-fine alignment driven call by call like src/visodo.cpp:1041-1281, with the commented-out variant of the
    reference selected: computeSigmaPdf (:1189-1190) + buildSystemGridStride(HUBER) (:1205-1214)."""
    pk_W, pk_I, pc_W, pc_I = [ops.up(W_kf)], [ops.up(I_kf)], [ops.up(W_cur)], [ops.up(I_cur)]
    for l in range(1, levels):
        pk_W.append(ops.pyr_down(pk_W[-1])); pk_I.append(ops.pyr_down(pk_I[-1]))
        pc_W.append(ops.pyr_down(pc_W[-1])); pc_I.append(ops.pyr_down(pc_I[-1]))
    R, t = np.eye(3), np.zeros(3)
    first_sums = None
    for l in range(levels - 1, -1, -1):
        d = float(1 << l)
        fx, fy, cx, cy = intr["fx"] / d, intr["fy"] / d, intr["cx"] / d, intr["cy"] / d
        gWx, gWy = ops.gradient(pk_W[l])
        gIx, gIy = ops.gradient(pk_I[l])
        for _ in range(its[l]):
            Rp, tp = orc.projective_pose(R, t, fx, fy, cx, cy, inverse=True)
            W1 = ops.warp_invdepth(pc_W[l], pk_W[l], Rp, tp)
            I1 = ops.warp_intensity(pc_I[l], W1, Rp, tp)
            bI, sI = ops.sigma_pdf(ops.compute_error(I1, pk_I[l], nsamples), 0.0, 5.0, orc.HUBER)
            bW, sW = ops.sigma_pdf(ops.compute_error(W1, pk_W[l], nsamples), 0.0, 0.0025, orc.HUBER)
            A, b = ops.build_system(pk_W[l], pk_I[l], gWx, gWy, gIx, gIy, W1, I1,
                                    ops.params(fx=fx, fy=fy, cx=cx, cy=cy, sW=sW, sI=sI, bW=bW, bI=bI))
            A, b = np.asarray(A, dtype=np.float64).reshape(6, 6), np.asarray(b, dtype=np.float64).reshape(6)
            if first_sums is None:
                first_sums = np.concatenate([np.concatenate([A[r, r:], [b[r]]]) for r in range(6)])
            R, t, _, bad = orc.gn_update(A, b, R, t)
            assert not bad
    return R, t, first_sums


@pytest.mark.parametrize("kind,nframes", [("ref", 20), ("cpu", 3)])
def test_config3_huber_stream_through_bridge_calls(ctx, kind, nframes):
    _need(kind)
    rows, cols, levels, its = 480, 640, 3, [4, 3, 2]
    seq = synth.make_sequence(seed=333, n_frames=nframes + 1, rows=rows, cols=cols, noise=True, device="cuda")
    intr = dict(synth.FREIBURG1)
    gpu, other = _GpuOps(ctx), (_CpuOps() if kind == "cpu" else _RefOps())
    worst_t = worst_r = worst_s = 0.0
    W_prev = I_prev = None
    for k in range(nframes + 1):
        d, c = seq["depth"][k].cpu().numpy().astype(np.uint16), seq["rgb"][k].cpu().numpy()
        W, I = orc.depth_to_invdepth(d), orc.intensity(c)
        if k > 0:  # frame-to-frame stream: the previous frame is the keyframe
            Rg, tg, sg = _huber_frame_to_frame(gpu, W_prev, I_prev, W, I, intr, levels, its)
            Ro, to, so = _huber_frame_to_frame(other, W_prev, I_prev, W, I, intr, levels, its)
            dt, ang = float(np.linalg.norm(tg - to)), rot_angle(Rg, Ro)
            worst_t, worst_r, worst_s = max(worst_t, dt), max(worst_r, ang), max(worst_s, sums_rel_err(sg, so))
            assert dt < POSE_TOL_M and ang < POSE_TOL_RAD, (k, dt, ang)
            gt_R, gt_t = synth.relative_pose(seq["poses"][k - 1], seq["poses"][k])
            assert np.linalg.norm(tg - gt_t) < 2e-3
        W_prev, I_prev = W, I
    print("config 3 vs %s: %d frames, worst |dt| = %.2e m, angle = %.2e rad, first-iteration sums %.2e" % (kind, nframes, worst_t, worst_r, worst_s))
    assert worst_s < 1e-4


# ---- a8: copyImage / initialiseDeviceMemory2D bridge ops -----------------------------------------------------------
def test_copy_and_fill_image(ctx):
    rows, cols = 37, 53  # ragged on purpose: pitch != 4 * cols on the destination view
    src = torch.randn(rows, cols, device="cuda")
    src[3, 5] = float("nan")
    big = torch.full((rows, 64), -7.0, device="cuda")
    dst = big[:, :cols]  # row-pitched view
    ctx.copy_image(src, dst)
    ctx.sync()
    assert torch.equal(torch.nan_to_num(dst, nan=123.0), torch.nan_to_num(src, nan=123.0))
    assert bool((big[:, cols:] == -7.0).all())  # nothing written beyond the row
    out = ctx.copy_image(src)
    assert torch.equal(torch.nan_to_num(out, nan=123.0), torch.nan_to_num(src, nan=123.0))
    ctx.fill_image(dst, 1.0)  # initialiseWeightKeyframe
    ctx.sync()
    assert bool((dst == 1.0).all()) and bool((big[:, cols:] == -7.0).all())
    ctx.fill_image(dst, float("nan"))
    ctx.sync()
    assert bool(torch.isnan(dst).all())
