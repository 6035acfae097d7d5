"""WARP_ORDER = warpFirst (src/visodo.cpp:1078-1105; the code default of the reference, `src/internal.h:107`; the shipped
config_data/visodoRGBDconfig.ini selects pyrFirst): above level 0 every Gauss-Newton iteration warps the current frame at
level 0 with the current pose and rebuilds the pyramid of the warped maps down to the level that iterates.  Parity of
the device-resident schedule against the CPU oracle and against the reference's own kernels driven the same way."""
import numpy as np
import pytest
import torch

from util import pair_maps, rot_angle, sums_rel_err
import oracle as orc
from oracle import ref as refk
from oracle.tracker import OracleTracker
from rgbid_slam_b200 import capi, host, synth
from test_align_gpu import _gpu_align, _oracle_align, _check, POSE_TOL_M, POSE_TOL_RAD, SUMS_TOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("noise", [False, True])
def test_warp_first_alignment_vs_oracle_and_reference(ctx, noise):
    rows, cols, levels, its = 480, 640, 3, [10, 5, 3]
    P = pair_maps(seed=20261020 + int(noise), rows=rows, cols=cols, noise=noise)
    al = _gpu_align(ctx, P, rows, cols, levels, capi.MODE_TRACKER, its, warp_first=1)
    out = al.run(want_trace=True)
    ref = _oracle_align(P, rows, cols, levels, orc.MODE_TRACKER, its, warp_first=1)
    _check(out, ref, label="warpFirst vs CPU oracle")
    assert np.linalg.norm(out["t"][0] - P["t_ab"]) < 5e-4 and rot_angle(out["R"][0], P["R_ab"]) < 5e-4
    tr = out["trace"][0]
    # first iteration (level 2, identical pose on both sides): sums from pyrDown^2(warp_0(.)) to the north-star bar
    assert tr[0]["level"] == 2
    assert sums_rel_err(tr[0]["sums27"], ref["trace"][0]["sums27"]) < SUMS_TOL
    assert tr[0]["nu_depthinv"] == ref["trace"][0]["nu_depthinv"] and tr[0]["nu_int"] == ref["trace"][0]["nu_int"]
    assert abs(tr[0]["sigma_int"] - ref["trace"][0]["sigma_int"]) / ref["trace"][0]["sigma_int"] < 1e-4
    # ... and it is not the pyrFirst computation
    pf = _gpu_align(ctx, P, rows, cols, levels, capi.MODE_TRACKER, its)
    out_pf = pf.run(want_trace=True)
    assert sums_rel_err(tr[0]["sums27"], out_pf["trace"][0][0]["sums27"]) > 1e-4
    # the covariance pass warps at the finest level in both orders
    assert sums_rel_err(tr[-1]["sums27"], ref["cov_sums27"], ignore_b=True) < 1e-4
    if refk.available():
        r2 = _oracle_align(P, rows, cols, levels, orc.MODE_TRACKER, its, kind="ref", warp_first=1)
        _check(out, r2, label="warpFirst vs reference CUDA kernels")
        assert sums_rel_err(tr[0]["sums27"], r2["trace"][0]["sums27"]) < SUMS_TOL
    pf.close()
    al.close()


def test_warp_first_batch_and_software_sampler(ctx, monkeypatch):
    """Batched streams give the same result as one; the software sampler (no texture objects) agrees with the texture path."""
    rows, cols = 240, 320
    P = pair_maps(seed=81, rows=rows, cols=cols, noise=True)
    one = _gpu_align(ctx, P, rows, cols, 3, capi.MODE_TRACKER, warp_first=1).run()
    many = _gpu_align(ctx, P, rows, cols, 3, capi.MODE_TRACKER, batch=3, warp_first=1).run()
    for b in range(3):
        assert np.allclose(many["t"][b], one["t"][0], atol=1e-6) and np.allclose(many["R"][b], one["R"][0], atol=1e-6)
    monkeypatch.setenv("RGBID_SAMPLER", "soft")
    soft = _gpu_align(ctx, P, rows, cols, 3, capi.MODE_TRACKER, warp_first=1).run()
    monkeypatch.delenv("RGBID_SAMPLER")
    assert np.linalg.norm(one["t"][0] - soft["t"][0]) < 1e-6 and rot_angle(one["R"][0], soft["R"][0]) < 1e-6


def test_warp_first_tracker_sequence(ctx):
    """The tracker state machine with the warpFirst schedule against the restated trackNewFrame (CPU oracle)."""
    rows, cols, n = 240, 320, 8
    seq = synth.make_sequence(seed=4343, n_frames=n, rows=rows, cols=cols, noise=True)
    intr = seq["intr"]
    acfg = host.make_align_config(rows, cols, 3, capi.MODE_TRACKER, batch=1, warp_first=1, **intr)
    trk = host.Tracker(ctx, host.make_tracker_config(acfg))
    ot = OracleTracker(rows, cols, intr, levels=3, iterations=(10, 5, 3), kind="cpu", warp_first=1)
    for k in range(n):
        d, c = seq["depth"][k], seq["rgb"][k]
        res = trk.track(d[None].contiguous(), c[None].contiguous())
        o = ot.track(d.numpy().astype(np.uint16), c.numpy())
        r = res[0]
        assert r.status == o["status"] == 0
        assert np.linalg.norm(np.array(r.t[:]) - o["t"]) < POSE_TOL_M
        assert rot_angle(np.array(r.R[:]).reshape(3, 3), o["R"]) < POSE_TOL_RAD
        assert r.new_odo_keyframe == o["new_odo_keyframe"] and r.new_integr_keyframe == o["new_integr_keyframe"], k
    trk.close()
