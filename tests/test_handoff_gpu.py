"""keyframe hand-off to the back end (resetIntegrationKeyframe, src/visodo.cpp:1577-1672) and the
sequential odometry constraints (:2126-2156) -- the C ABI's sink / frame results against the numpy restatement in
oracle/tracker.py on the same frames."""
import numpy as np
import pytest
import torch

from util import rot_angle
from oracle.tracker import OracleTracker
from rgbid_slam_b200 import capi, host, synth

pytestmark = pytest.mark.gpu


def test_keyframe_handoff_and_constraints(ctx):
    rows, cols, n = 240, 320, 12
    seq = synth.make_sequence(seed=2026, n_frames=n, rows=rows, cols=cols, noise=True)
    intr = seq["intr"]
    acfg = host.make_align_config(rows, cols, 3, capi.MODE_TRACKER, batch=2, **intr)
    # integration keyframes switch more often than odometry keyframes: the SEQ_KF chain then crosses odometry keyframes
    tcfg = host.make_tracker_config(acfg)
    tcfg.visratio_integr = 0.93
    tcfg.visratio_odo = 0.88
    trk = host.Tracker(ctx, tcfg)
    got = []
    trk.set_keyframe_sink(got.append)
    ot = OracleTracker(rows, cols, intr, levels=3, iterations=(10, 5, 3), kind="cpu", visratio_odo=0.88, visratio_integr=0.93)
    want, n_odo, n_int = [], 0, 0
    masks = {}
    for k in range(n):
        d, c = seq["depth"][k], seq["rgb"][k]
        mask_before = trk.overlap_mask(0).cpu().numpy().copy()
        res = trk.track(torch.stack([d, d]).contiguous(), torch.stack([c, c]).contiguous())
        o = ot.track(d.numpy().astype(np.uint16), c.numpy())
        r = res[0]
        assert r.new_odo_keyframe == o["new_odo_keyframe"] and r.new_integr_keyframe == o["new_integr_keyframe"], k
        if k > 0:
            n_odo += o["new_odo_keyframe"]; n_int += o["new_integr_keyframe"]
            sR, st, scov = o["seq"]
            assert rot_angle(np.array(r.seq_R[:]).reshape(3, 3), sR) < 1e-6 and np.linalg.norm(np.array(r.seq_t[:]) - st) < 1e-6
            got_cov = np.array(r.seq_cov[:]).reshape(6, 6)
            assert np.allclose(got_cov, got_cov.T, atol=1e-18) and np.abs(got_cov - scov).max() < 1e-3 * np.abs(scov).max()
        if o.get("kf_handoff") is not None:
            want.append(o["kf_handoff"])
            masks[o["kf_handoff"]["frame_index"]] = mask_before
    assert n_int >= 3 and n_odo < n_int, (n_odo, n_int)  # the configuration exercises the chain
    per_stream = [[g for g in got if g["stream"] == b] for b in range(2)]
    assert len(per_stream[0]) == len(per_stream[1]) == len(want)
    for g, g1, w in zip(per_stream[0], per_stream[1], want):
        assert (g["kf_index"], g["frame_index"]) == (w["kf_index"], w["frame_index"])
        assert rot_angle(g["R"], w["R"]) < 1e-5 and np.linalg.norm(g["t"] - w["t"]) < 1e-5
        assert rot_angle(g["rel_R"], w["rel_R"]) < 1e-5 and np.linalg.norm(g["rel_t"] - w["rel_t"]) < 1e-5
        assert np.abs(g["rel_cov"] - w["rel_cov"]).max() < 1e-3 * np.abs(w["rel_cov"]).max()
        # keyframe pose composed with the constraint is the pose of the frame that replaces it
        # (fused inverse depth and normals of the OUTGOING keyframe, as the reference downloads them)
        m = ~(np.isnan(g["depthinv"]) | np.isnan(w["depthinv"]))
        assert np.mean(np.isnan(g["depthinv"]) == np.isnan(w["depthinv"])) > 0.999
        assert np.mean(np.abs(g["depthinv"][m] - w["depthinv"][m]) / w["depthinv"][m] < 1e-4) > 0.999
        mn = ~(np.isnan(g["normals"]) | np.isnan(w["normals"]))
        assert np.mean(np.abs(g["normals"][mn] - w["normals"][mn]) < 1e-3) > 0.995
        # colours of the frame at which the keyframe was created; overlap mask as it was before the switch
        assert np.array_equal(g["colors"], seq["rgb"][w["kf_index"]].numpy())
        assert np.array_equal(g["overlap_mask"], masks[w["frame_index"]])
        # both streams carry the same data
        assert np.array_equal(g["depthinv"], g1["depthinv"], equal_nan=True) and np.array_equal(g["rel_cov"], g1["rel_cov"])
    trk.set_keyframe_sink(None)
    trk.close()
