"""Runs the C++ drop-in program (tests/cpp/test_dropin.cpp): the reference-shaped C++ surface -- DeviceArray2D,
RGBID_SLAM::device::* with the signatures of src/internal.h, VisodoTracker, KeyframeAlign -- and checks what it
computed against the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from util import rot_angle, sums_rel_err
import oracle as orc
from oracle.tracker import OracleTracker
from rgbid_slam_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_dropin")


def test_cpp_dropin_surface(built, tmp_path):
    assert os.path.exists(BIN), "tests/cpp/test_dropin was not built (see __graft_entry__.build)"
    rows, cols, n = 240, 320, 6
    seq = synth.make_sequence(seed=31, n_frames=n, rows=rows, cols=cols, noise=True)
    intr = seq["intr"]
    depth = seq["depth"].numpy().astype(np.uint16)
    rgb = seq["rgb"].numpy()
    seq_path, ini_path, out_path = tmp_path / "seq.bin", tmp_path / "calib.ini", tmp_path / "out.txt"
    with open(seq_path, "wb") as f:
        f.write(struct.pack("3i4f", n, rows, cols, intr["fx"], intr["fy"], intr["cx"], intr["cy"]))
        for k in range(n):
            f.write(depth[k].tobytes())
            f.write(rgb[k].tobytes())
    # calibration + tracker settings in the reference's INI dialect (config_data/*.ini)
    ini_path.write_text("[CALIBRATION]\nfx=%r\nfy=%r\ncx=%r\ncy=%r\n\n[VISODO]\n; comment\nM_ESTIMATOR = Student\n"
                        "SIGMA_ESTIMATOR = sigmaML\nODOMETRY_VISRATIO_THRESHOLD = 0.9\nINTEGRATION_VISRATIO_THRESHOLD = 0.7\n"
                        "FINEST_PYR_LEVEL = 0\nWARP_ORDER = pyrFirst\nIMAGE_FILTERING = none\n"
                        % (intr["fx"], intr["fy"], intr["cx"], intr["cy"]))
    r = subprocess.run([BIN, str(seq_path), str(ini_path), str(out_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = {}
    poses = []
    for ln in out_path.read_text().splitlines():
        tok = ln.split()
        if tok[0] == "pose":
            poses.append([float(v) for v in tok[1:]])
        else:
            lines[tok[0]] = [float(v) for v in tok[1:]]

    # 1. one iteration through the bridge functions vs the oracle
    WA, WB = orc.depth_to_invdepth(depth[0]), orc.depth_to_invdepth(depth[1])
    IA, IB = orc.intensity(rgb[0]), orc.intensity(rgb[1])
    Rp, tp = np.eye(3, dtype=np.float32), np.zeros(3, dtype=np.float32)
    W1 = orc.warp_invdepth(WB, WA, Rp, tp)
    I1 = orc.warp_intensity(IB, W1, Rp, tp)
    bi, si, nui, _ = orc.sigma_nu_student(orc.compute_error(I1, IA, 10000), 0.0, 5.0)
    bw, sw, nuw, _ = orc.sigma_nu_student(orc.compute_error(W1, WA, 10000), 0.0, 0.0025)
    gx, gy = orc.gradient(WA)
    hx, hy = orc.gradient(IA)
    p = orc.system_params(intr["fx"], intr["fy"], intr["cx"], intr["cy"], sigma_depthinv=sw, sigma_int=si, bias_depthinv=bw,
                          bias_int=bi, nu_depthinv=nuw, nu_int=max(nui, nuw))
    A, b, sums = orc.build_system(WA, IA, gx, gy, hx, hy, W1, I1, p)
    got = lines["scale"]
    assert abs(got[0] - si) / si < 1e-4 and abs(got[1] - sw) / sw < 1e-4 and got[4] == max(nui, nuw) and got[5] == nuw
    Ac, bc = np.array(lines["A"]).reshape(6, 6), np.array(lines["b"])
    got_sums = np.concatenate([np.concatenate([Ac[r_, r_:], [bc[r_]]]) for r_ in range(6)])
    assert np.array_equal(Ac, Ac.T) and sums_rel_err(got_sums, sums) < 1e-4
    assert abs(lines["vis"][0] - orc.visibility_ratio(WB, WA, Rp, tp)) < 2e-4
    assert lines["elapsed_ms"][0] > 0

    # 2. VisodoTracker vs the restated trackNewFrame
    ot = OracleTracker(rows, cols, intr, levels=3, iterations=(10, 5, 3), kind="cpu")
    for k in range(n):
        o = ot.track(depth[k], rgb[k])
        pr = poses[k]
        assert int(pr[0]) == k and int(pr[1]) == (1 if k > 0 else 0)
        assert np.linalg.norm(np.array(pr[11:14]) - o["t"]) < 1e-4 and rot_angle(np.array(pr[2:11]).reshape(3, 3), o["R"]) < 1e-4
        assert int(pr[14]) == o["new_odo_keyframe"] and int(pr[15]) == o["new_integr_keyframe"]

    # 2b. hand-off containers: one SEQ_ODO constraint per tracked frame, one keyframe + SEQ_KF per integration switch
    ho = lines["handoff"]
    n_int = sum(int(poses[k][15]) for k in range(1, n))
    assert int(ho[0]) == n_int and int(ho[1]) == n - 1 and int(ho[2]) == n_int
    if n_int:
        assert int(ho[3]) == 0 and int(ho[4]) > 0.5 * rows * cols  # first outgoing keyframe was created at frame 0

    # 3. KeyframeAlign (grey image rounded to 8 bit, as the Keyframe container stores it)
    G = [np.floor(orc.intensity(rgb[j]) + 0.5).astype(np.uint8).astype(np.float32) for j in (0, 2)]
    W = [orc.depth_to_invdepth(depth[j]) for j in (0, 2)]
    cfg = orc.make_config(rows, cols, 4, orc.MODE_ALIGN, [5, 5, 3, 0], intr["fx"], intr["fy"], intr["cx"], intr["cy"])
    ref = orc.align(cfg, orc.prepare_keyframe(W[0], G[0], 4, tracker=False), orc.prepare_current(W[1], G[1], 4))
    al = lines["align"]
    assert np.linalg.norm(np.array(al[9:12]) - ref["t"]) < 1e-4 and rot_angle(np.array(al[:9]).reshape(3, 3), ref["R"]) < 1e-4
    assert abs(al[12] - ref["cov"][0, 0]) / ref["cov"][0, 0] < 1e-3
